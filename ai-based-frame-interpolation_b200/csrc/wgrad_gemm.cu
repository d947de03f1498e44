// Weight gradient of a 3x3 convolution on the tensor cores (training step, reference model/train.py:196-197 backward).
//
//   dW[tap][co][ci] = sum over pixels q of dz[q][co] * x[q + tap][ci]
//
// is a GEMM whose reduction dimension is the pixel index. Both operands are first transposed to channel-major,
// zero-padded layouts (transpose_pad_kernel): dzT [Cout][Kp], xT [Cin][Kp] with Kp = N*(H+2)*(W+2) rounded up. In that
// layout the row part of a tap is a plain offset (dy-1)*Wp8 along K (a multiple of 16 bytes, as TMA box starts must be)
// and the column part selects one of three pre-shifted copies of xT; the zero border of dzT kills the products that
// would wrap around an image row — so the mainloop is an ordinary K-major tcgen05 GEMM (the descriptors validated in
// conv_gemm.cu): A = dzT rows [128 co] x 64 K, B = xT rows [N_TILE ci] x 64 K shifted by the tap, D = [128 x N_TILE]
// fp32 in TMEM. K is split over CTAs; partial tiles are added to dW with fp32 vector atomics.
#include "conv_gemm.cuh"
#include "ptx.cuh"
#include "train_kernels.cuh"

#include <cstring>

namespace fi {

namespace {

constexpr int WG_THREADS = 192;
constexpr int WG_A_BYTES = 128 * 128;

struct WgradParams {
    int cout, cin, m_tiles, n_tiles, k_chunks, chunk_slabs, total_slabs, wp;
    float* dW;
};

__host__ __device__ constexpr int wg_stages(int n_tile) { return n_tile == 256 ? 4 : 6; }
__host__ __device__ constexpr int wg_smem(int n_tile) {
    return 1024 + wg_stages(n_tile) * (WG_A_BYTES + n_tile * 128) + 256;
}

template <int N_TILE>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const WgradParams p) {
    constexpr int STAGES = wg_stages(N_TILE);
    constexpr int B_BYTES = N_TILE * 128;
    constexpr uint32_t IDESC = umma_idesc_bf16(128, N_TILE);
    constexpr int TMEM_COLS = 2 * N_TILE < 32 ? 32 : 2 * N_TILE;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = base, smem_b = base + STAGES * WG_A_BYTES;
    const uint32_t bar = smem_b + STAGES * B_BYTES;
    const uint32_t bar_full = bar, bar_empty = bar + 8 * STAGES, bar_tfull = bar + 16 * STAGES, bar_tempty = bar_tfull + 16;
    const uint32_t tmem_slot = bar_tempty + 16;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
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

    // work item = (tap, m tile, n tile, k chunk); k chunk fastest so that neighbouring CTAs share operand rows in L2
    const int total = 9 * p.m_tiles * p.n_tiles * p.k_chunks;
    auto decode = [&](int t, int& tap, int& mt, int& nt, int& s0, int& s1) {
        const int kc = t % p.k_chunks;
        int r = t / p.k_chunks;
        nt = r % p.n_tiles;
        r /= p.n_tiles;
        mt = r % p.m_tiles;
        tap = r / p.m_tiles;
        s0 = kc * p.chunk_slabs;
        s1 = s0 + p.chunk_slabs < p.total_slabs ? s0 + p.chunk_slabs : p.total_slabs;
    };

    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            int tap, mt, nt, s0, s1;
            decode(t, tap, mt, nt, s0, s1);
            const int off = (tap / 3 - 1) * p.wp;   // row shift: a multiple of 8 elements (16 B), as TMA requires
            const int brow = (tap % 3) * p.cin;      // column shift: the pre-shifted copy of xT
            for (int s = s0; s < s1; ++s) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                const uint32_t full = bar_full + 8 * stage;
                if (elect_one()) {
                    mbar_expect_tx(full, WG_A_BYTES + B_BYTES);
                    tma_load_2d(smem_a + stage * WG_A_BYTES, &map_a, full, s * 64, mt * 128);
                    tma_load_2d(smem_b + stage * B_BYTES, &map_b, full, s * 64 + off, brow + nt * N_TILE);
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
            int tap, mt, nt, s0, s1;
            decode(t, tap, mt, nt, s0, s1);
            const int acc = it & 1;
            mbar_wait(bar_tempty + 8 * acc, ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * N_TILE;
            for (int s = s0; s < s1; ++s) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                const uint64_t da = umma_desc_sw128(smem_a + stage * WG_A_BYTES);
                const uint64_t db = umma_desc_sw128(smem_b + stage * B_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, IDESC, (s > s0) || (k > 0));
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
            int tap, mt, nt, s0, s1;
            decode(t, tap, mt, nt, s0, s1);
            const int acc = it & 1;
            mbar_wait(bar_tfull + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * N_TILE;
            const int co = mt * 128 + q * 32 + lane;
            float* row = p.dW + (static_cast<size_t>(tap) * p.cout + co) * p.cin + nt * N_TILE;
#pragma unroll 1
            for (int c = 0; c < N_TILE / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(taddr + c * 32, v);
                tmem_ld_wait();
                if (co < p.cout) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        atomicAdd(reinterpret_cast<float4*>(row + c * 32 + 4 * j),
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

template <int N_TILE>
const char* launch_wgrad(const CUtensorMap& ma, const CUtensorMap& mb, const WgradParams& p, int grid, cudaStream_t st) {
    auto k = wgrad_kernel<N_TILE>;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, wg_smem(N_TILE)) != cudaSuccess)
            return "wgrad: cudaFuncSetAttribute failed";
        configured = true;
    }
    k<<<grid, WG_THREADS, wg_smem(N_TILE), st>>>(ma, mb, p);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace

const char* wgrad_launch(const void* dzT, const void* xT, int cout, int cin, long long Kp, int Wp, float* dW,
                         int num_sms, cudaStream_t st) {
    if (!dzT || !xT || !dW) return "wgrad: null operand";
    if (cin % 64 || cout % 8 || Kp % 64 || Kp <= 0 || Wp % 8) return "wgrad: cin % 64, Kp % 64 and row pitch % 8 must be 0";
    const int n_tile = cin % 256 == 0 ? 256 : (cin % 128 == 0 ? 128 : 64);
    WgradParams p;
    memset(&p, 0, sizeof p);
    p.cout = cout;
    p.cin = cin;
    p.m_tiles = (cout + 127) / 128;
    p.n_tiles = cin / n_tile;
    p.total_slabs = static_cast<int>(Kp / 64);
    p.wp = Wp;
    p.dW = dW;
    // split K so that there are a few work items per SM
    const int base_items = 9 * p.m_tiles * p.n_tiles;
    int chunks = (4 * num_sms + base_items - 1) / base_items;
    if (chunks > p.total_slabs) chunks = p.total_slabs;
    if (chunks < 1) chunks = 1;
    p.chunk_slabs = (p.total_slabs + chunks - 1) / chunks;
    p.k_chunks = (p.total_slabs + p.chunk_slabs - 1) / p.chunk_slabs;
    alignas(64) CUtensorMap ma, mb;
    const char* e;
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(Kp), static_cast<uint64_t>(cout)};
        const uint64_t strides[1] = {static_cast<uint64_t>(Kp)};
        const uint32_t box[2] = {64, 128};
        if ((e = encode_bf16_map_public(&ma, dzT, 2, dims, strides, box))) return e;
    }
    {
        const uint64_t dims[2] = {static_cast<uint64_t>(Kp), static_cast<uint64_t>(3 * cin)};  // three shifted copies
        const uint64_t strides[1] = {static_cast<uint64_t>(Kp)};
        const uint32_t box[2] = {64, static_cast<uint32_t>(n_tile)};
        if ((e = encode_bf16_map_public(&mb, xT, 2, dims, strides, box))) return e;
    }
    const long long total = 9LL * p.m_tiles * p.n_tiles * p.k_chunks;
    const int grid = static_cast<int>(total < num_sms ? total : num_sms);
    switch (n_tile) {
        case 64: return launch_wgrad<64>(ma, mb, p, grid, st);
        case 128: return launch_wgrad<128>(ma, mb, p, grid, st);
        default: return launch_wgrad<256>(ma, mb, p, grid, st);
    }
}

}  // namespace fi
