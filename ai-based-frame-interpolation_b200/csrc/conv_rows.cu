// Row-stacked variant of the tcgen05 implicit-GEMM conv3x3 for the 64-output-channel layers at full resolution
// (up4.conv.0 and up4.conv.3 + head of reference model/unet.py:82, 35-63).
//
// Why: with M = 128 pixels and N = 64 output channels every tcgen05.mma reads 4 KB of A and 2 KB of B from shared
// memory for 32 cycles of tensor work, so the 50.8-cycle shared-memory floor (tools/probe/mma_probe.cu) caps these
// layers at 63 % of the tensor peak whatever the rest of the kernel does. Here the three taps of one filter ROW are
// stacked along N: for a fixed dy one MMA multiplies the pixel rows by [W(dy,-1) | W(dy,0) | W(dy,+1)] (N = 192), i.e.
//     D[q][j*64 + co] = sum_dy sum_ci A[q + dy*row][ci] * W(dy, j-1)[co][ci]
// 12 instead of 36 MMAs per 64-channel slab, each reading 4 + 6 KB for 96 tensor cycles: tensor-bound. The dx shift
// moves to the epilogue: out[p] = D[p-1][0:64] + D[p][64:128] + D[p+1][128:192]. The accumulator rows of one warp are
// 32 consecutive pixels of ONE image row (TMEM lane = pixel), so p-1 / p+1 are the neighbouring lanes: two warp shuffles
// per output channel; lanes 0 and 31 only feed their neighbours (tile = 4 rows x 30 output columns, 94 % of the MMA rows).
//
// Shared memory: A stages of one TMA box {64 ch, 32 px, 6 rows} = 24 KB (rows = y*32 + x: the view of filter row dy is
// the plain 128-row SW128 tile at +dy*4096 B — no shifted descriptors), all 9 x SLABS [64 x 64] weight slabs resident
// (the three slabs of a filter row are consecutive: one N = 192 B operand). TMEM: 2 accumulators x 192 columns.
// Warp roles (608 threads): 0 = A producer, 1 = TMEM owner + MMA issuer, 2 = weight loads, 3..18 = two epilogue sets
// of eight warps that alternate tiles (two warps per TMEM lane quadrant, 32 output channels each: the epilogue reads
// three times the accumulator columns of the halo kernels and spends most of its time waiting for shuffles, so it needs
// the extra warps to stay ahead of the MMAs). EPI_STORE only (bf16 NHWC, 16-byte global stores: a pixel's 64 channels
// are one 128 B line). Summation order differs from conv_halo.cu (fp32), results agree to bf16 noise.
//
// OPT-IN (FI_ROWS=1: layers with two K slabs per tap, i.e. up4.conv.0; FI_ROWS=2: one-slab layers as well). Measured
// on B200 at four 1080p pairs (profiles/r02_rows.md) it is parity-green but SLOWER than the halo kernels: up4.conv.0
// 1.05 ms against 0.96 ms (CTA-pair halo kernel), and with eight epilogue warps and the head epilogue up4.conv.3 took
// 0.69 ms against 0.60 ms. The MMAs do run tensor-bound (ncu: 66 % tensor-pipe activity against 55 % for the halo
// kernels while they run), but the epilogue has to move three times the accumulator columns out of TMEM (96 KB per
// 120 output pixels, ~43 B/cycle achieved) and pay two shuffles per channel: 42 % of its samples wait on the shuffle
// scoreboard, the issuer waits 28 % of the time for a free accumulator, and doubling the epilogue warps (8 -> 16)
// only moved 1.14 -> 1.05 ms. The shared-memory port floor of N = 64 MMAs is traded for a TMEM-read floor of about the
// same height; the default therefore stays with the halo kernels.
#include "conv_gemm.cuh"
#include "ptx.cuh"

#include <cstdlib>
#include <cstring>

namespace fi {

namespace {

constexpr int RT_W = 30, RT_H = 4;                   // output tile
constexpr int RB_W = 32, RB_H = 6;                   // halo box: x0-1 .. x0+30, y0-1 .. y0+4
constexpr int R_A_BYTES = RB_W * RB_H * 128;         // 24576
constexpr int R_ROW_BYTES = RB_W * 128;              // 4096: one image row of the box = 32 GEMM rows
constexpr int R_W_BYTES = 64 * 128;                  // one [64 x 64] weight slab
constexpr int R_THREADS = 32 * 19;
constexpr int R_ACC_COLS = 192;
constexpr int R_TMEM_COLS = 512;

__host__ __device__ constexpr int rows_a_stages(int slabs) { return slabs == 1 ? 4 : 3; }
__host__ __device__ constexpr int rows_smem_bytes(int slabs) {
    return 1024 + rows_a_stages(slabs) * R_A_BYTES + slabs * 9 * R_W_BYTES + 512;
}

struct RTile {
    int img, y0, x0;
};
__device__ __forceinline__ RTile decode_rtile(int t, const ConvKernelParams& p) {
    const int per_img = p.tiles_y * p.tiles_x;
    RTile c;
    c.img = t / per_img;
    const int m = t - c.img * per_img;
    const int ty = m / p.tiles_x;
    c.y0 = ty * RT_H;
    c.x0 = (m - ty * p.tiles_x) * RT_W;
    return c;
}

template <int SLABS>
__global__ void __launch_bounds__(R_THREADS, 1)
conv_rows_kernel(const __grid_constant__ ConvMaps maps, const ConvKernelParams p) {
    constexpr int A_STAGES = rows_a_stages(SLABS);
    constexpr uint32_t IDESC = umma_idesc_bf16(128, R_ACC_COLS);

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_a + A_STAGES * R_A_BYTES;
    const uint32_t smem_bar = smem_b + SLABS * 9 * R_W_BYTES;
    const uint32_t bar_afull = smem_bar;                     // A_STAGES
    const uint32_t bar_aempty = bar_afull + 8 * A_STAGES;    // A_STAGES
    const uint32_t bar_tfull = bar_aempty + 8 * A_STAGES;    // 2
    const uint32_t bar_tempty = bar_tfull + 16;              // 2
    const uint32_t bar_bres = bar_tempty + 16;               // 1: all weights landed
    const uint32_t tmem_slot = bar_bres + 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a[0]);
        tma_prefetch_desc(&maps.a[2]);
        tma_prefetch_desc(&maps.b);
        for (int s = 0; s < A_STAGES; ++s) {
            mbar_init(bar_afull + 8 * s, 1);
            mbar_init(bar_aempty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 8);   // the eight warps of the epilogue set that owns accumulator a
        }
        mbar_init(bar_bres, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, R_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_launch_dependents();
    pdl_wait();

    const int total_tiles = p.n_img * p.tiles_y * p.tiles_x;

    if (warp == 0) {
        // ------------------------------------------------------------ A producer: one halo box per (tile, slab)
        int stage = 0;
        uint32_t phase = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const RTile tc = decode_rtile(t, p);
            const int tp = t + p.prefetch_dist * static_cast<int>(gridDim.x);
            if (p.prefetch_dist > 0 && tp < total_tiles && lane == 0) {
                const RTile pc = decode_rtile(tp, p);
                for (int g = 0; g < p.nseg; ++g) {
                    const int mid = p.seg_map[g];
                    const int ox = mid >= 2 ? p.off_x : 0, oy = mid >= 2 ? p.off_y : 0;
                    for (int c = 0; c < p.seg_slabs[g]; ++c)
                        tma_prefetch_l2_4d(&maps.a[mid], c * BLOCK_K, pc.x0 - 1 - ox, pc.y0 - 1 - oy, pc.img);
                }
            }
            int seg = 0, left = p.seg_slabs[0];
            for (int s = 0; s < SLABS; ++s) {
                while (left == 0) left = p.seg_slabs[++seg];
                const int local = p.seg_slabs[seg] - left;
                --left;
                const int mid = p.seg_map[seg];
                mbar_wait(bar_aempty + 8 * stage, phase ^ 1);
                const uint32_t full = bar_afull + 8 * stage;
                if (elect_one()) {
                    mbar_expect_tx(full, R_A_BYTES);
                    const int ox = mid >= 2 ? p.off_x : 0, oy = mid >= 2 ? p.off_y : 0;  // F.pad of src1
                    tma_load_4d(smem_a + stage * R_A_BYTES, &maps.a[mid], full, local * BLOCK_K, tc.x0 - 1 - ox,
                                tc.y0 - 1 - oy, tc.img);
                }
                __syncwarp();
                if (++stage == A_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 2) {
        // ------------------------------------------------------------ weights, once: slab (s, tap) at (s*9 + tap) * 8 KB
        if (lane == 0) {
            mbar_expect_tx(bar_bres, SLABS * 9 * R_W_BYTES);
            for (int s = 0; s < SLABS; ++s)
                for (int tap = 0; tap < 9; ++tap)
                    tma_load_2d(smem_b + (s * 9 + tap) * R_W_BYTES, &maps.b, bar_bres, (tap * SLABS + s) * BLOCK_K, 0);
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        mbar_wait(bar_bres, 0);
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(bar_tempty + 8 * acc, ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * R_ACC_COLS;
            for (int s = 0; s < SLABS; ++s) {
                mbar_wait(bar_afull + 8 * stage, phase);
                tc_fence_after();
                const uint32_t a_base = smem_a + stage * R_A_BYTES;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const uint64_t da = umma_desc_sw128(a_base + dy * R_ROW_BYTES);
                    const uint64_t db = umma_desc_sw128(smem_b + (s * 9 + dy * 3) * R_W_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / 16; ++k)
                            umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, IDESC, (s | dy | k) != 0);
                        if (dy == 2) {
                            umma_commit(bar_aempty + 8 * stage);
                            if (s == SLABS - 1) umma_commit(bar_tfull + 8 * acc);
                        }
                    }
                    __syncwarp();
                }
                if (++stage == A_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps 3..18: set s owns the tiles with
        // (iteration & 1) == s and accumulator s; warp q = warp & 3 reads TMEM lanes [32q, 32q+32) = image row y0 + q,
        // output channels [32*half, 32*half + 32).
        const int q = warp & 3;
        const int set = (warp - 3) >> 3;
        const int half = ((warp - 3) >> 2) & 1;
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            if ((it & 1) != set) continue;
            const RTile tc = decode_rtile(t, p);
            mbar_wait(bar_tfull + 8 * set, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + set * R_ACC_COLS;
            const int y = tc.y0 + q;
            const int x = tc.x0 - 1 + lane;
            const bool inside = lane >= 1 && lane <= RT_W && y < p.H && x < p.W;
            uint8_t* dst_px = static_cast<uint8_t*>(p.dst) +
                              ((static_cast<size_t>(tc.img) * p.H + (inside ? y : 0)) * p.W + (inside ? x : 0)) * 128;
#pragma unroll 1
            for (int cc = 2 * half; cc < 2 * half + 2; ++cc) {   // 16 output channels at a time
                uint32_t b0[16], b1[16], b2[16];
                tmem_ld_32x32b_x16(taddr + cc * 16, b0);
                tmem_ld_32x32b_x16(taddr + 64 + cc * 16, b1);
                tmem_ld_32x32b_x16(taddr + 128 + cc * 16, b2);
                tmem_ld_wait();
                float f[16];
                const float4* bias4 = reinterpret_cast<const float4*>(p.bias + cc * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 b = __ldg(bias4 + j);
                    const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int i = 4 * j + e;
                        const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(b0[i]), 1);
                        const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(b2[i]), 1);
                        float v = (left + __uint_as_float(b1[i])) + right + bb[e];
                        if (p.relu) v = fmaxf(v, 0.0f);
                        f[i] = v;
                    }
                }
                if (inside) {
                    {
                        uint4 lo, hi;
                        lo.x = pack_bf16x2(f[0], f[1]);
                        lo.y = pack_bf16x2(f[2], f[3]);
                        lo.z = pack_bf16x2(f[4], f[5]);
                        lo.w = pack_bf16x2(f[6], f[7]);
                        hi.x = pack_bf16x2(f[8], f[9]);
                        hi.y = pack_bf16x2(f[10], f[11]);
                        hi.z = pack_bf16x2(f[12], f[13]);
                        hi.w = pack_bf16x2(f[14], f[15]);
                        uint4* d4 = reinterpret_cast<uint4*>(dst_px + cc * 32);
                        d4[0] = lo;
                        d4[1] = hi;
                    }
                }
            }
            // every TMEM read of this accumulator is complete: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive(bar_tempty + 8 * set);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, R_TMEM_COLS);
    }
}

template <int SLABS>
const char* launch_rows_inst(const ConvLaunch& l, cudaStream_t stream) {
    auto kfn = conv_rows_kernel<SLABS>;
    static std::atomic<uint64_t> configured{0};
    constexpr int smem = rows_smem_bytes(SLABS);
    static_assert(smem <= 232448, "row-stacked kernel exceeds the 227 KB shared memory limit");
    if (!smem_opt_in(kfn, smem, configured)) return "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
    const cudaError_t e = launch_kernel(kfn, dim3(l.grid), dim3(R_THREADS), smem, stream, l.maps, l.p);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace

// Opt-in through FI_ROWS (see the header). Cout = 64, plain bf16 store (no pooled output: the 2x2 window would span two
// epilogue warps; no head: with one K slab the epilogue is already the longer side), bf16 mode.
bool conv_rows_eligible(const ConvDesc& d) {
    const char* env = getenv("FI_ROWS");
    if (!env || (env[0] != '1' && env[0] != '2')) return false;
    if (d.taps != 9 || d.n_total != 64 || d.precise || d.mode != EPI_STORE) return false;
    if (d.c0 % BLOCK_K || d.c1 % BLOCK_K) return false;
    const int slabs = (d.c0 + d.c1) / BLOCK_K;
    return slabs == 2 || (slabs == 1 && env[0] == '2');
}

void conv_rows_geometry(int* tile_w, int* tile_h, int* box_w, int* box_h) {
    *tile_w = RT_W;
    *tile_h = RT_H;
    *box_w = RB_W;
    *box_h = RB_H;
}

const char* conv_rows_launch(const ConvLaunch& l, cudaStream_t stream) {
    if (l.mode != EPI_STORE) return "conv(rows): no kernel instantiation for this mode";
    if (l.p.slabs == 1) return launch_rows_inst<1>(l, stream);
    if (l.p.slabs == 2) return launch_rows_inst<2>(l, stream);
    return "conv(rows): no kernel instantiation for this K";
}

}  // namespace fi
