// CTA-pair (cta_group::2) variant of the per-tap implicit-GEMM conv for the wide layers (Cout multiple of 256).
//
// Why: measured with tools/probe/mma_probe.cu, a tcgen05.mma costs max(N/2, operand bytes per SM / ~120) cycles, and
// the TMA fills compete for the same ~128 B/cycle of shared-memory bandwidth. A single CTA doing M128 x N256 reads
// 12 KB of operands and receives 12 KB of new tiles per MMA-time: 192 B/cycle of demand. In a CTA pair each SM keeps
// its own 128 pixel rows of A but only HALF of the weight slab (128 of the 256 rows); the pair's MMA is M256 x N256,
// the hardware reads both halves, and each SM's traffic drops to 8 KB + 8 KB per 128 cycles = 128 B/cycle.
//
// Protocol (cluster of 2, same code in both CTAs, rank 0 = leader):
//   producers (warp 0 of both)   wait on their LOCAL empty barrier, load their own A tile and their half of B with
//                                cp.async.bulk.tensor...cta_group::2, completion bytes go to the LEADER's full barrier
//   MMA issuer (warp 1, leader)  waits on its full barrier (2 x 32 KB), issues tcgen05.mma.cta_group::2, and commits
//                                with a multicast to the empty / accumulator-full barriers of BOTH CTAs
//   epilogue (warps 2..5, both)  read their own TMEM lanes; release the accumulator by arriving (locally or through
//                                mapa) on the leader's accumulator-empty barrier (8 arrivals per phase)
#include "conv_epilogue.cuh"
#include "conv_gemm.cuh"
#include "ptx.cuh"

#include <cstdio>

namespace fi {

namespace {

constexpr int C2_THREADS = 192;
constexpr int C2_BLOCK_N = 256;
constexpr int C2_A_BYTES = BLOCK_M * 128;            // 16 KB: this CTA's 128 pixel rows
constexpr int C2_B_BYTES = (C2_BLOCK_N / 2) * 128;   // 16 KB: this CTA's half of the weight slab
constexpr int C2_STAGE_BYTES = C2_A_BYTES + C2_B_BYTES;
constexpr int C2_STAGES = 5;
constexpr int C2_EPI_BYTES = 4 * (2 * 4096 + 2 * 1024);
constexpr int C2_SMEM = 1024 + C2_STAGES * C2_STAGE_BYTES + C2_EPI_BYTES + 256;
static_assert(C2_SMEM <= 232448, "pair kernel exceeds shared memory");

struct PairTile {
    EpiTile t;
    bool valid;
};
__device__ __forceinline__ PairTile decode_pair(int qi, uint32_t rank, const ConvKernelParams& p) {
    const int per_img = p.tiles_y * p.tiles_x;
    const int m_tiles = p.n_img * per_img;
    const int pairs = (m_tiles + 1) >> 1;
    PairTile r;
    r.t.nb = qi / pairs;
    int m = 2 * (qi - r.t.nb * pairs) + static_cast<int>(rank);
    r.valid = m < m_tiles;          // odd tile count: the last pair's second CTA recomputes the last tile, stores off
    if (!r.valid) m = m_tiles - 1;
    r.t.img = m / per_img;
    m -= r.t.img * per_img;
    const int ty = m / p.tiles_x;
    r.t.y0 = ty * TILE_H;
    r.t.x0 = (m - ty * p.tiles_x) * TILE_W;
    return r;
}

template <int MODE, bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(C2_THREADS, 1)
conv_gemm2_kernel(const __grid_constant__ ConvMaps maps, const ConvKernelParams p) {
    constexpr uint32_t IDESC = umma_idesc_bf16(256, C2_BLOCK_N);
    constexpr int TMEM_COLS = 512;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_a + C2_STAGES * C2_A_BYTES;
    const uint32_t smem_stage = smem_b + C2_STAGES * C2_B_BYTES;
    const uint32_t smem_pool = smem_stage + 4 * 2 * 4096;
    const uint32_t smem_bar = smem_pool + 4 * 2 * 1024;
    const uint32_t bar_full = smem_bar;
    const uint32_t bar_empty = smem_bar + 8 * C2_STAGES;
    const uint32_t bar_tfull = smem_bar + 16 * C2_STAGES;
    const uint32_t bar_tempty = bar_tfull + 16;
    const uint32_t tmem_slot = bar_tempty + 16;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int n_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&maps.a[0]);
        tma_prefetch_desc(&maps.a[2]);
        tma_prefetch_desc(&maps.b);
        tma_prefetch_desc(&maps.out[0]);
        for (int s = 0; s < C2_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);   // leader: its own arrive.expect_tx; bytes arrive from both CTAs
            mbar_init(bar_empty + 8 * s, 1);  // one multicast commit per use
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 8);  // leader: 4 epilogue warps of each CTA
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers exist before any remote signal can arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_launch_dependents();
    pdl_wait();

    const int per_img = p.tiles_y * p.tiles_x;
    const int pairs_per_nb = (p.n_img * per_img + 1) >> 1;
    const int total = p.n_blocks * pairs_per_nb;
    const int slabs = p.slabs;
    const int k_iters = p.taps * slabs;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        int stage = 0;
        uint32_t phase = 0;
        for (int qi = cluster_id; qi < total; qi += n_clusters) {
            const PairTile pt = decode_pair(qi, rank, p);
            for (int tap = 0; tap < p.taps; ++tap) {
                const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
                const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
                int seg = 0, left = p.seg_slabs[0];
                for (int s = 0; s < slabs; ++s) {
                    while (left == 0) left = p.seg_slabs[++seg];
                    const int local = p.seg_slabs[seg] - left;
                    --left;
                    const int mid = p.seg_map[seg];
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t full = bar_full + 8 * stage;
                    if (elect_one()) {
                        if (leader) mbar_expect_tx(full, 2 * C2_STAGE_BYTES);
                        const int ox = mid >= 2 ? p.off_x : 0, oy = mid >= 2 ? p.off_y : 0;
                        tma_load_4d_2sm(smem_a + stage * C2_A_BYTES, &maps.a[mid], full, local * BLOCK_K,
                                        pt.t.x0 + dx - ox, pt.t.y0 + dy - oy, pt.t.img);
                        tma_load_2d_2sm(smem_b + stage * C2_B_BYTES, &maps.b, full, (tap * slabs + s) * BLOCK_K,
                                        pt.t.nb * C2_BLOCK_N + static_cast<int>(rank) * (C2_BLOCK_N / 2));
                    }
                    __syncwarp();
                    if (++stage == C2_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int qi = cluster_id; qi < total; qi += n_clusters, ++it) {
                const int acc = it & 1;
                mbar_wait(bar_tempty + 8 * acc, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * C2_BLOCK_N;
                for (int kb = 0; kb < k_iters; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(smem_a + stage * C2_A_BYTES);
                    const uint64_t db = umma_desc_sw128(smem_b + stage * C2_B_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / 16; ++k)
                            umma_bf16_ss_2sm(d_tmem, da + 2 * k, db + 2 * k, IDESC, (kb | k) != 0);
                        umma_commit_2sm(bar_empty + 8 * stage);
                        if (kb == k_iters - 1) umma_commit_2sm(bar_tfull + 8 * acc);
                    }
                    __syncwarp();
                    if (++stage == C2_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps 2..5 (both CTAs)
        const int q = warp & 3;
        const uint32_t my_stage = smem_stage + q * (2 * 4096);
        const uint32_t my_pool = smem_pool + q * (2 * 1024);
        int buf = 0;
        int it = 0;
        for (int qi = cluster_id; qi < total; qi += n_clusters, ++it) {
            const PairTile pt = decode_pair(qi, rank, p);
            const int acc = it & 1;
            mbar_wait(bar_tfull + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * C2_BLOCK_N;
#pragma unroll 1
            for (int c = 0; c < C2_BLOCK_N / 64; ++c) {
                epilogue_chunk_8x16<C2_BLOCK_N, MODE, SPLIT>(maps, p, pt.t, taddr, c, q, lane, my_stage, my_pool, buf,
                                                             pt.valid);
            }
            tc_fence_before();
            __syncwarp();
            if (elect_one()) {
                if (leader) mbar_arrive(bar_tempty + 8 * acc);
                else mbar_arrive_cluster(bar_tempty + 8 * acc, 0);
            }
        }
        __syncwarp();
        if (elect_one()) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer may still be reading TMEM / receiving commits
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    }
}

template <int MODE, bool SPLIT = false>
const char* launch_pair_inst(const ConvLaunch& l, cudaStream_t stream) {
    if constexpr (!SPLIT) {
        if (l.split) return launch_pair_inst<MODE, true>(l, stream);
    }
    auto kfn = conv_gemm2_kernel<MODE, SPLIT>;
    static std::atomic<uint64_t> configured{0};
    if (!smem_opt_in(kfn, C2_SMEM, configured)) return "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
    const cudaError_t e = launch_kernel(kfn, dim3(l.grid), dim3(C2_THREADS), C2_SMEM, stream, l.maps, l.p);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace

const char* conv_pair_launch(const ConvLaunch& l, cudaStream_t stream) {
    switch (l.mode) {
        case EPI_STORE: return launch_pair_inst<EPI_STORE>(l, stream);
        case EPI_STORE_POOL: return launch_pair_inst<EPI_STORE_POOL>(l, stream);
        case EPI_CONVT: return launch_pair_inst<EPI_CONVT>(l, stream);
        default: return "conv(pair): no kernel instantiation for this mode";
    }
}

}  // namespace fi
