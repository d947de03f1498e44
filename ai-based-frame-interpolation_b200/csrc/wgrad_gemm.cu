// Weight gradient of a 3x3 convolution on the tensor cores (training step, reference model/train.py:196-197 backward).
//
//   dW[tap][co][ci] = sum over pixels q of dz[q][co] * x[q + tap][ci]
//
// is a GEMM whose reduction dimension is the pixel index. In NHWC both operands have that dimension as the slow one, i.e.
// they are "MN-major" tcgen05 operands as they lie in HBM: a TMA box {64 channels, bw, bh, 1 image} with bw*bh = 64 lands
// in shared memory as 64 pixel rows of 128 swizzled bytes — the canonical MN-major SWIZZLE_128B atom (8 K-rows x 128 B,
// atoms 1024 B apart along K, 64-channel blocks LBO apart along M/N). A tap is a coordinate shift of the box and the zero
// fill of out-of-bounds TMA reads is the convolution's padding, so nothing is transposed, padded or copied.
//
// GEMM shape. The nine taps are stacked along N: a column block is (shift of x, 64 input channels), so one dz slab in
// shared memory feeds 3-4 column blocks (N = 192 / 256) whatever the layer's channel count. Rows are output channels; a
// 64-output-channel layer would fill only half of the 128 MMA rows, so there the second row block is dz itself shifted
// one pixel to the left (dW[a + c] = sum_q dz[q - a] x[q + c]): rows = {a.dx = 0, 1} x 64 co, columns = {c.dy = -1,0,1} x
// {c.dx = -1, 0} x ci, and the duplicate (a.dx, c.dx) = (1, -1) of tap.dx = 0 is dropped in the epilogue (6 of 8 useful).
// The reduction is split over CTAs; partial tiles are added to dW with fp32 vector atomics.
#include "conv_gemm.cuh"
#include "ptx.cuh"
#include "train_kernels.cuh"

#include <cstring>

namespace fi {

namespace {

constexpr int WG_THREADS = 192;
constexpr int WG_BLOCK_BYTES = 64 * 128;        // one 64-channel x 64-pixel block
constexpr int WG_A_BYTES = 2 * WG_BLOCK_BYTES;  // M = 128 rows

struct WgradParams {
    int cout, cin, c0;                          // c0 = channels of the first x source (concat layers have two)
    int co_blocks, ci_blocks;                   // 64-channel blocks
    int stacked;                                // 1: cout == 64, rows = two column shifts of dz
    int taps;                                   // 9 (conv3x3) or 1 (pointwise: dW[co][ci] = sum_q dz[q][co] x[q][ci])
    int m_tiles, n_tiles;
    int k_chunks, chunk_slabs, total_slabs;     // split of the pixel slabs over work items
    int tiles_w, tiles_h, bw, bh;               // a slab is a bw x bh pixel box of one image
    float* dW;
};

__host__ __device__ constexpr int wg_stages(int n_blocks) { return n_blocks == 4 ? 4 : (n_blocks == 3 ? 5 : 6); }
__host__ __device__ constexpr int wg_smem(int n_blocks) {
    return 1024 + wg_stages(n_blocks) * (WG_A_BYTES + n_blocks * WG_BLOCK_BYTES) + 256;
}

// MN-major SWIZZLE_128B operand: 8 K-rows of 128 B per atom, atoms 1024 B apart along K, 64-element blocks `lbo` apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// column block nb -> shift of x and its 64-channel block
__device__ __forceinline__ void decode_col_block(const WgradParams& p, int nb, int& dy, int& dx, int& cb) {
    const int sidx = nb / p.ci_blocks;
    cb = nb - sidx * p.ci_blocks;
    if (p.taps == 1) {
        dy = dx = 0;
    } else if (p.stacked) {
        dy = sidx / 2 - 1;
        dx = sidx % 2 - 1;   // -1, 0
    } else {
        dy = sidx / 3 - 1;
        dx = sidx % 3 - 1;
    }
}

template <int N_BLOCKS>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_x0,
             const __grid_constant__ CUtensorMap map_x1, const WgradParams p) {
    constexpr int STAGES = wg_stages(N_BLOCKS);
    constexpr int N_TILE = 64 * N_BLOCKS;
    constexpr int B_BYTES = N_BLOCKS * WG_BLOCK_BYTES;
    constexpr uint32_t IDESC = umma_idesc_bf16(128, N_TILE) | (1u << 15) | (1u << 16);  // A and B MN-major
    constexpr int TMEM_COLS = 512;   // two accumulators of 192 or 256 columns

    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = base, smem_b = base + STAGES * WG_A_BYTES;
    const uint32_t bar = smem_b + STAGES * B_BYTES;
    const uint32_t bar_full = bar, bar_empty = bar + 8 * STAGES, bar_tfull = bar + 16 * STAGES, bar_tempty = bar_tfull + 16;
    const uint32_t tmem_slot = bar_tempty + 16;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_dz);
        tma_prefetch_desc(&map_x0);
        tma_prefetch_desc(&map_x1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // work item = (k chunk, m tile, n tile), tile fastest: CTAs running together share the same pixel slabs in L2
    const int tiles = p.m_tiles * p.n_tiles;
    const int total = tiles * p.k_chunks;
    auto decode = [&](int t, int& mt, int& nt, int& s0, int& s1) {
        const int kc = t / tiles;
        const int r = t - kc * tiles;
        mt = r / p.n_tiles;
        nt = r - mt * p.n_tiles;
        s0 = kc * p.chunk_slabs;
        s1 = s0 + p.chunk_slabs < p.total_slabs ? s0 + p.chunk_slabs : p.total_slabs;
    };

    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            int mt, nt, s0, s1;
            decode(t, mt, nt, s0, s1);
            const int a_blocks = p.stacked ? 2 : (p.co_blocks - 2 * mt < 2 ? p.co_blocks - 2 * mt : 2);
            int bdy[N_BLOCKS], bdx[N_BLOCKS], bcb[N_BLOCKS];
#pragma unroll
            for (int j = 0; j < N_BLOCKS; ++j) decode_col_block(p, nt * N_BLOCKS + j, bdy[j], bdx[j], bcb[j]);
            for (int s = s0; s < s1; ++s) {
                const int tw = s % p.tiles_w;
                const int r = s / p.tiles_w;
                const int th = r % p.tiles_h, img = r / p.tiles_h;
                const int w0 = tw * p.bw, h0 = th * p.bh;
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                const uint32_t full = bar_full + 8 * stage;
                if (elect_one()) {
                    mbar_expect_tx(full, (a_blocks + N_BLOCKS) * WG_BLOCK_BYTES);
                    const uint32_t sa = smem_a + stage * WG_A_BYTES, sb = smem_b + stage * B_BYTES;
                    for (int b = 0; b < a_blocks; ++b) {   // rows past cout are never stored: their block is not loaded
                        if (p.stacked) tma_load_4d(sa + b * WG_BLOCK_BYTES, &map_dz, full, 0, w0 - b, h0, img);
                        else tma_load_4d(sa + b * WG_BLOCK_BYTES, &map_dz, full, (2 * mt + b) * 64, w0, h0, img);
                    }
#pragma unroll
                    for (int j = 0; j < N_BLOCKS; ++j) {
                        const int c = bcb[j] * 64;
                        if (c < p.c0) tma_load_4d(sb + j * WG_BLOCK_BYTES, &map_x0, full, c, w0 + bdx[j], h0 + bdy[j], img);
                        else tma_load_4d(sb + j * WG_BLOCK_BYTES, &map_x1, full, c - p.c0, w0 + bdx[j], h0 + bdy[j], img);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            int mt, nt, s0, s1;
            decode(t, mt, nt, s0, s1);
            const int acc = it & 1;
            mbar_wait(bar_tempty + 8 * acc, ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            for (int s = s0; s < s1; ++s) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                const uint64_t da = umma_desc_mn_sw128(smem_a + stage * WG_A_BYTES, WG_BLOCK_BYTES);
                const uint64_t db = umma_desc_mn_sw128(smem_b + stage * B_BYTES, WG_BLOCK_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)  // 16 pixel rows (2048 B) per MMA
                        umma_bf16_ss(d_tmem, da + 128 * k, db + 128 * k, IDESC, (s > s0) || (k > 0));
                    umma_commit(bar_empty + 8 * stage);
                    if (s == s1 - 1) umma_commit(bar_tfull + 8 * acc);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        const int q = warp & 3;
        int it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            int mt, nt, s0, s1;
            decode(t, mt, nt, s0, s1);
            const int acc = it & 1;
            mbar_wait(bar_tfull + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
            const int row = q * 32 + lane;
            const int adx = p.stacked ? row >> 6 : 0;                      // warp-uniform (q)
            const int co = p.stacked ? (row & 63) : mt * 128 + row;
#pragma unroll 1
            for (int c = 0; c < N_TILE / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(taddr + c * 32, v);
                tmem_ld_wait();
                int dy, dx, cb;
                decode_col_block(p, nt * N_BLOCKS + (c >> 1), dy, dx, cb);
                const bool duplicate = adx == 1 && dx == -1;               // tap.dx = 0 is produced by (0, 0)
                if (co < p.cout && !duplicate) {
                    const int tap = p.taps == 1 ? 0 : (dy + 1) * 3 + (dx + adx + 1);
                    float* out = p.dW + (static_cast<size_t>(tap) * p.cout + co) * p.cin + cb * 64 + (c & 1) * 32;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        atomicAdd(reinterpret_cast<float4*>(out + 4 * j),
                                  make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                              __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive(bar_tempty + 8 * acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int N_BLOCKS>
const char* launch_wgrad(const CUtensorMap* maps, const WgradParams& p, int grid, cudaStream_t st) {
    auto k = wgrad_kernel<N_BLOCKS>;
    static std::atomic<uint64_t> configured{0};
    if (!smem_opt_in(k, wg_smem(N_BLOCKS), configured)) return "wgrad: cudaFuncSetAttribute failed";
    k<<<grid, WG_THREADS, wg_smem(N_BLOCKS), st>>>(maps[0], maps[1], maps[2], p);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* encode_pixels(CUtensorMap* map, const void* base, int N, int H, int W, int C, int bw, int bh) {
    const uint64_t dims[4] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                              static_cast<uint64_t>(N)};
    const uint64_t strides[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(W) * C,
                                 static_cast<uint64_t>(H) * W * C};
    const uint32_t box[4] = {64, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), 1};
    return encode_bf16_map_public(map, base, 4, dims, strides, box);
}

}  // namespace

const char* wgrad_launch(const void* dz, const void* x0, int c0, const void* x1, int c1, int N, int H, int W, int cout,
                         float* dW, int num_sms, cudaStream_t st, int taps) {
    if (!dz || !x0 || !dW || (c1 > 0 && !x1)) return "wgrad: null operand";
    if (taps != 9 && taps != 1) return "wgrad: taps must be 9 or 1";
    if (N <= 0 || H <= 0 || W <= 0) return "wgrad: empty shape";
    if (c0 <= 0 || c0 % 64 || c1 < 0 || c1 % 64 || cout <= 0 || cout % 64)
        return "wgrad: channel counts must be multiples of 64";
    const int cin = c0 + c1;
    WgradParams p;
    memset(&p, 0, sizeof p);
    p.cout = cout;
    p.cin = cin;
    p.c0 = c0;
    p.dW = dW;
    p.co_blocks = cout / 64;
    p.ci_blocks = cin / 64;
    p.taps = taps;
    p.stacked = (cout == 64 && taps == 9) ? 1 : 0;
    p.m_tiles = p.stacked ? 1 : (p.co_blocks + 1) / 2;
    const int col_blocks = (taps == 1 ? 1 : (p.stacked ? 6 : 9)) * p.ci_blocks;   // 3x3: always a multiple of 3
    const int n_blocks = col_blocks % 4 == 0 ? 4 : (col_blocks % 3 == 0 ? 3 : (col_blocks % 2 == 0 ? 2 : 1));
    if (n_blocks == 1) return "wgrad: pointwise mode needs an even number of 64-channel input blocks";
    p.n_tiles = col_blocks / n_blocks;
    // pixel slab: the 64-pixel box shape that wastes the fewest out-of-bounds pixels
    long long best = -1;
    for (int bw = 16; bw >= 1; bw >>= 1) {
        const int bh = 64 / bw;
        const long long covered = static_cast<long long>((W + bw - 1) / bw) * ((H + bh - 1) / bh);
        if (best < 0 || covered < best) {
            best = covered;
            p.bw = bw;
            p.bh = bh;
        }
    }
    p.tiles_w = (W + p.bw - 1) / p.bw;
    p.tiles_h = (H + p.bh - 1) / p.bh;
    p.total_slabs = N * p.tiles_w * p.tiles_h;
    // split the reduction so that there are a few work items per SM
    const int base_items = p.m_tiles * p.n_tiles;
    int chunks = (4 * num_sms + base_items - 1) / base_items;
    if (chunks > p.total_slabs) chunks = p.total_slabs;
    if (chunks < 1) chunks = 1;
    p.chunk_slabs = (p.total_slabs + chunks - 1) / chunks;
    p.k_chunks = (p.total_slabs + p.chunk_slabs - 1) / p.chunk_slabs;
    alignas(64) CUtensorMap maps[3];
    const char* e;
    if ((e = encode_pixels(&maps[0], dz, N, H, W, cout, p.bw, p.bh))) return e;
    if ((e = encode_pixels(&maps[1], x0, N, H, W, c0, p.bw, p.bh))) return e;
    if ((e = encode_pixels(&maps[2], c1 > 0 ? x1 : x0, N, H, W, c1 > 0 ? c1 : c0, p.bw, p.bh))) return e;
    const long long total = static_cast<long long>(base_items) * p.k_chunks;
    const int grid = static_cast<int>(total < num_sms ? total : num_sms);
    return n_blocks == 4 ? launch_wgrad<4>(maps, p, grid, st)
                         : (n_blocks == 3 ? launch_wgrad<3>(maps, p, grid, st) : launch_wgrad<2>(maps, p, grid, st));
}

}  // namespace fi
