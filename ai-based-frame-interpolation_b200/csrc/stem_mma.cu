// Stem convolution (inc.double_conv.0, reference model/unet.py:12-14 with C_in = n_channels <= 8) on the tensor cores.
//
// K = 9*C_in is tiny (18 for the grey frame pair), so the layer is bound by writing 128 B/pixel of bf16 NHWC output;
// the fp32 CUDA-core version spent 1152 FMA lane-ops per pixel and ran at ~20 % of that bound. Here producer threads
// build the im2col row of their pixel directly in a 128B-swizzled smem tile and one tcgen05.mma slab does the math.
// To keep fp32-grade accuracy on the un-rounded input (the whole point of an exact stem, SURVEY.md "hard parts"), both
// operands are split into bf16 hi + lo parts and three of the four cross products are accumulated:
//     A row  = [ x_hi (KT) | x_hi (KT) | x_lo (KT) | 0 ... ]        KT = 9*C_in, padded to a multiple of 64
//     B row  = [ w_hi (KT) | w_lo (KT) | w_hi (KT) | 0 ... ]        -> sum = x_hi*w_hi + x_hi*w_lo + x_lo*w_hi
// (the dropped x_lo*w_lo term is O(2^-18) relative). Input normalisation u8/255*2-1 (reference inference.py:32-35),
// the frame-pair torch.cat (unet.py:109) and the conv zero padding are all done by the gather.
// C_in > 4 (the colour pair, C_in = 6): the row is laid out tap-major with the channels padded to 8, so that one tap of
// one pixel is exactly one 16-byte chunk: K = 3 segments x 9 taps x 8 = 216 (4 slabs). The input tile is parked in
// shared memory as one hi chunk and one lo chunk per pixel, and a pixel's row is 27 chunk copies (LDS.128 -> STS.128)
// instead of ~2 000 instructions of 16-bit element shuffling; two A stages so that two CTAs fit per SM (measured by
// knocking phases out: gather, row build and the MMA/epilogue skeleton were strictly serial in a lone CTA, 1.45 ms for
// four 1080p colour pairs; two co-resident CTAs overlap them: 0.90 ms).
// Warp roles (288 threads): 0..3 = im2col producers (thread = pixel of the 8x16 tile), 4 = MMA issuer + TMEM owner,
// 5..8 = epilogue (bias + ReLU -> bf16 -> swizzled staging -> TMA store). Two CTAs fit per SM.
#include "aux_kernels.cuh"
#include "conv_gemm.cuh"
#include "ptx.cuh"

#include <cstring>

namespace fi {

namespace {

constexpr int SM_THREADS = 288;
__host__ __device__ constexpr int stem_a_stages(int cin) { return cin > 4 ? 2 : 3; }   // chunked rows: 2 CTAs per SM
constexpr int SM_A_BYTES = 128 * 128;  // 128 pixels x 64 bf16

struct StemParams {
    PlaneSrc src[2];
    int N, H, W, tiles_x, tiles_y;
    int split;  // precise mode: also store lo = bf16(value - hi)
    int linear; // no ReLU
    const float* bias;
};

__host__ __device__ constexpr bool stem_chunked(int cin) { return cin > 4; }   // tap-major rows, channels padded to 8
__host__ __device__ constexpr int stem_slabs(int cin) { return stem_chunked(cin) ? 4 : (27 * cin + 63) / 64; }
constexpr int SM_IN_ROWS = TILE_H + 2, SM_IN_COLS = TILE_W + 2, SM_IN_PITCH = 20;  // input tile + halo, padded pitch
__host__ __device__ constexpr int stem_in_bytes(int cin) {
    return stem_chunked(cin) ? 2 * 2 * SM_IN_ROWS * SM_IN_COLS * 16   // two buffers x (hi, lo) x 180 pixels x 16 B
                             : 2 * cin * SM_IN_ROWS * SM_IN_PITCH * 4;
}
__host__ __device__ constexpr int stem_smem_bytes(int cin) {
    return 1024 + stem_a_stages(cin) * SM_A_BYTES + stem_slabs(cin) * 8192 + 4 * 2 * 4096 + 256 + 1040 + stem_in_bytes(cin);
}

__device__ __forceinline__ float norm_u8_stem(uint8_t u) {
    return __fsub_rn(__fmul_rn(2.0f, __fdiv_rn(static_cast<float>(u), 255.0f)), 1.0f);
}
__device__ __forceinline__ uint32_t bf16_bits(float v) {  // round-to-nearest-even bf16, as the upper 16 bits
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v), "f"(0.0f));
    return r & 0xffff0000u;
}

template <int CIN, bool U8>
__global__ void __launch_bounds__(SM_THREADS, (CIN <= 3 || CIN > 4) ? 2 : 1)
stem_mma_kernel(const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out,
                const __grid_constant__ CUtensorMap map_out_lo, const StemParams p) {
    constexpr int KT = 9 * CIN;
    constexpr int SLABS = stem_slabs(CIN);
    constexpr int SM_A_STAGES = stem_a_stages(CIN);
    constexpr uint32_t IDESC = umma_idesc_bf16(128, 64);
    constexpr int TMEM_COLS = 128;  // 2 accumulators x 64 columns

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_a + SM_A_STAGES * SM_A_BYTES;
    const uint32_t smem_stage = smem_b + SLABS * 8192;
    const uint32_t smem_bar = smem_stage + 4 * 2 * 4096;
    const uint32_t bar_full = smem_bar;                       // SM_A_STAGES
    const uint32_t bar_empty = bar_full + 8 * SM_A_STAGES;    // SM_A_STAGES
    const uint32_t bar_tfull = bar_empty + 8 * SM_A_STAGES;   // 2
    const uint32_t bar_tempty = bar_tfull + 16;               // 2
    const uint32_t bar_bres = bar_tempty + 16;                // 1
    const uint32_t tmem_slot = bar_bres + 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    // u8 -> (bf16 hi | bf16 lo << 16) of the normalised value, and the double-buffered split input tile
    uint32_t* lut = reinterpret_cast<uint32_t*>(smem_raw + (smem_bar + 256 - smem_u32(smem_raw)));
    uint32_t* in_tile = lut + 260;  // 257 LUT slots (slot 256 = zero padding), padded

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (U8 && threadIdx.x < 256) {
        const float v = norm_u8_stem(static_cast<uint8_t>(threadIdx.x));
        const uint32_t h = bf16_bits(v);
        lut[threadIdx.x] = (h >> 16) | bf16_bits(v - __uint_as_float(h));
        if (threadIdx.x == 0) lut[256] = 0u;
    }
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
        for (int s = 0; s < SM_A_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 4);   // one arrive per producer warp
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 4);
        }
        mbar_init(bar_bres, 1);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_launch_dependents();
    pdl_wait();

    const int per_img = p.tiles_y * p.tiles_x;
    const int total_tiles = p.N * per_img;

    if (warp < 4 && stem_chunked(CIN)) {
        // ------------------------------------------------------------ chunked producers (C_in > 4): thread = pixel
        if constexpr (stem_chunked(CIN)) {
            const int m = threadIdx.x;  // 0..127, tile row = m / 16, tile column = m % 16
            constexpr int NPX = SM_IN_ROWS * SM_IN_COLS;   // 180 tile pixels incl. halo
            constexpr int PER = (NPX + 127) / 128;         // tile pixels converted by one thread
            uint4* chunk_tile = reinterpret_cast<uint4*>(in_tile);   // [buffer][hi | lo][NPX] 16-byte chunks
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            uint32_t raw[PER][CIN];
            auto fetch = [&](int tile_idx) {
                const int img = tile_idx / per_img;
                const int r = tile_idx - img * per_img;
                const int y0 = (r / p.tiles_x) * TILE_H, x0 = (r % p.tiles_x) * TILE_W;
#pragma unroll
                for (int k = 0; k < PER; ++k) {
                    const int i = m + 128 * k;
                    const int rr = i / SM_IN_COLS, col = i - rr * SM_IN_COLS;
                    const int yy = y0 + rr - 1, xx = x0 + col - 1;
                    const bool in = i < NPX && yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
#pragma unroll
                    for (int c = 0; c < CIN; ++c) {
                        uint32_t v = U8 ? 256u : 0u;  // out of bounds -> zero padding (LUT slot 256 / +0.0f)
                        if (in) {
                            const bool first = c < p.src[0].channels;
                            const PlaneSrc& sp = first ? p.src[0] : p.src[1];
                            const int cc = first ? c : c - p.src[0].channels;
                            const long long off = img * sp.batch_stride + cc * sp.chan_stride + yy * sp.row_stride +
                                                  xx * sp.px_stride;
                            if (U8) v = __ldg(static_cast<const uint8_t*>(sp.ptr) + off);
                            else v = __float_as_uint(__ldg(static_cast<const float*>(sp.ptr) + off));
                        }
                        raw[k][c] = v;
                    }
                }
            };
            if (static_cast<int>(blockIdx.x) < total_tiles) fetch(blockIdx.x);
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                uint4* hi_tile = chunk_tile + (it & 1) * 2 * NPX;
                uint4* lo_tile = hi_tile + NPX;
#pragma unroll
                for (int k = 0; k < PER; ++k) {
                    const int i = m + 128 * k;
                    if (i < NPX) {
                        uint32_t pk[8];   // per channel: bf16 hi in the low half, bf16 lo in the high half
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            if (c >= CIN) pk[c] = 0u;
                            else if (U8) pk[c] = lut[raw[k][c]];
                            else {
                                const float v = __uint_as_float(raw[k][c]);
                                const uint32_t h = bf16_bits(v);
                                pk[c] = (h >> 16) | bf16_bits(v - __uint_as_float(h));
                            }
                        }
                        hi_tile[i] = make_uint4(__byte_perm(pk[0], pk[1], 0x5410), __byte_perm(pk[2], pk[3], 0x5410),
                                                __byte_perm(pk[4], pk[5], 0x5410), __byte_perm(pk[6], pk[7], 0x5410));
                        lo_tile[i] = make_uint4(__byte_perm(pk[0], pk[1], 0x7632), __byte_perm(pk[2], pk[3], 0x7632),
                                                __byte_perm(pk[4], pk[5], 0x7632), __byte_perm(pk[6], pk[7], 0x7632));
                    }
                }
                if (t + static_cast<int>(gridDim.x) < total_tiles) fetch(t + gridDim.x);
                asm volatile("bar.sync 1, 128;" ::: "memory");   // two buffers: one barrier per tile (see below)
                const int base_px = (m >> 4) * SM_IN_COLS + (m & 15);
#pragma unroll
                for (int s = 0; s < SLABS; ++s) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t row = smem_a + stage * SM_A_BYTES + m * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int q = s * 8 + j;          // chunk index in the row: [hi x 9 | hi x 9 | lo x 9 | 0 x 5]
                        uint4 v = make_uint4(0u, 0u, 0u, 0u);
                        if (q < 27) {
                            const int tap = q % 9;
                            const int n = base_px + (tap / 3) * SM_IN_COLS + (tap % 3);
                            v = q < 18 ? hi_tile[n] : lo_tile[n];
                        }
                        st_shared_v4(row + ((j ^ (m & 7)) << 4), v.x, v.y, v.z, v.w);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + 8 * stage);
                    if (++stage == SM_A_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp < 4) {
        if constexpr (!stem_chunked(CIN)) {
        // ------------------------------------------------------------ im2col producers: thread = pixel
        // Phase A (cooperative): the (8+2)x(16+2) input tile of every channel is loaded once, normalised and split into
        // bf16 hi/lo, and parked in smem. Phase B: each thread assembles the im2col row of its pixel from 9*CIN LDS.
        const int m = threadIdx.x;  // 0..127, tile row = m / 16, tile column = m % 16
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        constexpr int PLANE = SM_IN_ROWS * SM_IN_PITCH;
        constexpr int IN_ELEMS = CIN * SM_IN_ROWS * SM_IN_COLS;
        constexpr int NLOAD = (IN_ELEMS + 127) / 128;
        // raw input values of the NEXT tile travel in registers while the current tile is assembled: the global-load
        // latency (the stall that dominated the first version of this kernel) overlaps phase B.
        uint32_t raw[NLOAD];
        auto fetch = [&](int tile_idx) {
            const int img = tile_idx / per_img;
            const int r = tile_idx - img * per_img;
            const int y0 = (r / p.tiles_x) * TILE_H, x0 = (r % p.tiles_x) * TILE_W;
#pragma unroll
            for (int k = 0; k < NLOAD; ++k) {
                const int i = m + 128 * k;
                uint32_t v = U8 ? 256u : 0u;  // out of bounds -> zero padding (LUT slot 256 / +0.0f)
                if (i < IN_ELEMS) {
                    const int c = i / (SM_IN_ROWS * SM_IN_COLS);
                    const int rr = (i - c * SM_IN_ROWS * SM_IN_COLS) / SM_IN_COLS;
                    const int col = i - c * SM_IN_ROWS * SM_IN_COLS - rr * SM_IN_COLS;
                    const int yy = y0 + rr - 1, xx = x0 + col - 1;
                    if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
                        const bool first = c < p.src[0].channels;
                        const PlaneSrc& sp = first ? p.src[0] : p.src[1];
                        const int cc = first ? c : c - p.src[0].channels;
                        const long long off = img * sp.batch_stride + cc * sp.chan_stride + yy * sp.row_stride +
                                              xx * sp.px_stride;
                        if (U8) v = __ldg(static_cast<const uint8_t*>(sp.ptr) + off);
                        else v = __float_as_uint(__ldg(static_cast<const float*>(sp.ptr) + off));
                    }
                }
                raw[k] = v;
            }
        };
        if (static_cast<int>(blockIdx.x) < total_tiles) fetch(blockIdx.x);
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            uint32_t* tile = in_tile + (it & 1) * CIN * PLANE;
#pragma unroll
            for (int k = 0; k < NLOAD; ++k) {
                const int i = m + 128 * k;
                if (i < IN_ELEMS) {
                    const int c = i / (SM_IN_ROWS * SM_IN_COLS);
                    const int rr = (i - c * SM_IN_ROWS * SM_IN_COLS) / SM_IN_COLS;
                    const int col = i - c * SM_IN_ROWS * SM_IN_COLS - rr * SM_IN_COLS;
                    uint32_t packed;
                    if (U8) {
                        packed = lut[raw[k]];
                    } else {
                        const float v = __uint_as_float(raw[k]);
                        const uint32_t h = bf16_bits(v);
                        packed = (h >> 16) | bf16_bits(v - __uint_as_float(h));
                    }
                    tile[c * PLANE + rr * SM_IN_PITCH + col] = packed;
                }
            }
            if (t + static_cast<int>(gridDim.x) < total_tiles) fetch(t + gridDim.x);
            // one barrier per tile is enough with two buffers: a thread re-writes buffer b only after every producer
            // passed the barrier of the tile in between, i.e. after they all finished reading b
            asm volatile("bar.sync 1, 128;" ::: "memory");
            uint32_t hl[KT];  // low half = bf16 hi part, high half = bf16 lo part
            const uint32_t* px = tile + (m >> 4) * SM_IN_PITCH + (m & 15);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                for (int c = 0; c < CIN; ++c)
                    hl[tap * CIN + c] = px[c * PLANE + (tap / 3) * SM_IN_PITCH + (tap % 3)];
            auto elem = [&](int e) -> uint32_t {  // 16-bit K element e of the row [x_hi | x_hi | x_lo | 0]
                return e < KT ? (hl[e] & 0xffffu)
                              : (e < 2 * KT ? (hl[e - KT] & 0xffffu) : (e < 3 * KT ? (hl[e - 2 * KT] >> 16) : 0u));
            };
#pragma unroll
            for (int s = 0; s < SLABS; ++s) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                const uint32_t row = smem_a + stage * SM_A_BYTES + m * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t wv[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int e = s * 64 + j * 8 + q * 2;
                        wv[q] = elem(e) | (elem(e + 1) << 16);
                    }
                    st_shared_v4(row + ((j ^ (m & 7)) << 4), wv[0], wv[1], wv[2], wv[3]);
                }
                fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full + 8 * stage);
                if (++stage == SM_A_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        }
    } else if (warp == 4) {
        // ------------------------------------------------------------ weights (once) + MMA issue
        if (elect_one()) {
            mbar_expect_tx(bar_bres, SLABS * 8192);
            for (int s = 0; s < SLABS; ++s) tma_load_2d(smem_b + s * 8192, &map_b, bar_bres, s * 64, 0);
        }
        __syncwarp();
        mbar_wait(bar_bres, 0);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(bar_tempty + 8 * acc, ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 64;
            for (int s = 0; s < SLABS; ++s) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                const uint64_t da = umma_desc_sw128(smem_a + stage * SM_A_BYTES);
                const uint64_t db = umma_desc_sw128(smem_b + s * 8192);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, IDESC, (s | k) != 0);
                    umma_commit(bar_empty + 8 * stage);
                    if (s == SLABS - 1) umma_commit(bar_tfull + 8 * acc);
                }
                __syncwarp();
                if (++stage == SM_A_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue warps 5..8
        const int q = warp & 3;
        const uint32_t my_stage = smem_stage + q * 8192;
        int buf = 0;
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int img = t / per_img;
            const int r = t - img * per_img;
            const int y0 = (r / p.tiles_x) * TILE_H, x0 = (r % p.tiles_x) * TILE_W;
            const int acc = it & 1;
            mbar_wait(bar_tfull + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 64;
            uint32_t v0[32], v1[32];
            tmem_ld_32x32b_x32(taddr, v0);
            tmem_ld_32x32b_x32(taddr + 32, v1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive(bar_tempty + 8 * acc);  // accumulator is in registers: release it early
            const float4* bias4 = reinterpret_cast<const float4*>(p.bias);
            const float lo_clamp = p.linear ? -3.0e38f : 0.f;  // ReLU, or (training forward) none
            // split: buffer 0 = hi tile, buffer 1 = lo tile (both must be free); else the two buffers alternate
            if (elect_one()) {
                if (p.split) tma_store_wait_read<0>();
                else tma_store_wait_read<1>();
            }
            __syncwarp();
            const uint32_t sbuf = p.split ? my_stage : my_stage + buf * 4096;
            const uint32_t slo = my_stage + 4096;
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // 16-byte chunk j = channels 8j .. 8j+7
                const float4 b0 = __ldg(bias4 + 2 * j), b1 = __ldg(bias4 + 2 * j + 1);
                const uint32_t* v = j < 4 ? v0 : v1;
                const int o = (j & 3) * 8;
                float f[8];
                f[0] = fmaxf(__uint_as_float(v[o + 0]) + b0.x, lo_clamp);
                f[1] = fmaxf(__uint_as_float(v[o + 1]) + b0.y, lo_clamp);
                f[2] = fmaxf(__uint_as_float(v[o + 2]) + b0.z, lo_clamp);
                f[3] = fmaxf(__uint_as_float(v[o + 3]) + b0.w, lo_clamp);
                f[4] = fmaxf(__uint_as_float(v[o + 4]) + b1.x, lo_clamp);
                f[5] = fmaxf(__uint_as_float(v[o + 5]) + b1.y, lo_clamp);
                f[6] = fmaxf(__uint_as_float(v[o + 6]) + b1.z, lo_clamp);
                f[7] = fmaxf(__uint_as_float(v[o + 7]) + b1.w, lo_clamp);
                uint32_t hw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) hw[k] = pack_bf16x2(f[2 * k], f[2 * k + 1]);
                const uint32_t off = lane * 128 + ((j ^ (lane & 7)) << 4);
                st_shared_v4(sbuf + off, hw[0], hw[1], hw[2], hw[3]);
                if (p.split) {
                    uint32_t lw[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        lw[k] = pack_bf16x2(f[2 * k] - bf16lo_f(hw[k]), f[2 * k + 1] - bf16hi_f(hw[k]));
                    st_shared_v4(slo + off, lw[0], lw[1], lw[2], lw[3]);
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one()) {
                tma_store_4d(&map_out, sbuf, 0, x0, y0 + 2 * q, img);  // box {64, 16, 2, 1}
                if (p.split) tma_store_4d(&map_out_lo, slo, 0, x0, y0 + 2 * q, img);
                tma_store_commit();
            }
            buf ^= 1;
        }
        __syncwarp();
        if (elect_one()) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int CIN>
const char* launch_stem(const StemDesc& d, const CUtensorMap& map_b, const CUtensorMap& map_out,
                        const CUtensorMap& map_out_lo, const StemParams& p, int grid, cudaStream_t stream) {
    constexpr int smem = stem_smem_bytes(CIN);
    static std::atomic<uint64_t> configured[2];
    if (d.is_u8) {
        auto k = stem_mma_kernel<CIN, true>;
        if (!smem_opt_in(k, smem, configured[1])) return "stem: cudaFuncSetAttribute failed";
        launch_kernel(k, dim3(grid), dim3(SM_THREADS), smem, stream, map_b, map_out, map_out_lo, p);
    } else {
        auto k = stem_mma_kernel<CIN, false>;
        if (!smem_opt_in(k, smem, configured[0])) return "stem: cudaFuncSetAttribute failed";
        launch_kernel(k, dim3(grid), dim3(SM_THREADS), smem, stream, map_b, map_out, map_out_lo, p);
    }
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

uint16_t host_bf16(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7fffu + ((u >> 16) & 1u);
    return static_cast<uint16_t>(u >> 16);
}
float host_bf16_to_f32(uint16_t h) {
    const uint32_t u = static_cast<uint32_t>(h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

}  // namespace

int stem_packed_k(int cin) { return stem_slabs(cin) * 64; }

void stem_pack_weights(const float* w, int cin, uint16_t* out) {
    // w: fp32 [64][cin][3][3] (BN already folded); out: bf16 [64][stem_packed_k(cin)] = [w_hi | w_lo | w_hi | 0]
    // cin > 4: tap-major with the channels padded to 8 (segment length 72), matching the chunked producer
    const bool chunked = stem_chunked(cin);
    const int kt = chunked ? 72 : 9 * cin, kp = stem_packed_k(cin);
    for (int co = 0; co < 64; ++co) {
        uint16_t* row = out + static_cast<size_t>(co) * kp;
        for (int k = 0; k < kp; ++k) row[k] = 0;
        for (int tap = 0; tap < 9; ++tap)
            for (int c = 0; c < cin; ++c) {
                const float v = w[(static_cast<size_t>(co) * cin + c) * 9 + tap];
                const uint16_t h = host_bf16(v);
                const uint16_t l = host_bf16(v - host_bf16_to_f32(h));
                const int k = chunked ? tap * 8 + c : tap * cin + c;
                row[k] = h;
                row[kt + k] = l;
                row[2 * kt + k] = h;
            }
    }
}

const char* stem_conv_launch(const StemDesc& d, cudaStream_t stream) {
    if (d.cin < 1 || d.cin > 8) return "stem: 1..8 input channels supported";
    if (d.src[0].channels + d.src[1].channels != d.cin) return "stem: plane sources do not add up to cin";
    if (d.N <= 0 || d.H <= 0 || d.W <= 0) return "stem: empty shape";
    if (!d.wpack || !d.bias || !d.dst || !d.src[0].ptr) return "stem: null operand";
    StemParams p;
    memset(&p, 0, sizeof p);
    p.src[0] = d.src[0];
    p.src[1] = d.src[1];
    p.N = d.N;
    p.H = d.H;
    p.W = d.W;
    p.tiles_x = (d.W + TILE_W - 1) / TILE_W;
    p.tiles_y = (d.H + TILE_H - 1) / TILE_H;
    p.bias = d.bias;
    p.split = d.dst_lo != nullptr;
    p.linear = d.linear;
    const long long tiles = static_cast<long long>(d.N) * p.tiles_x * p.tiles_y;
    if (tiles > 0x7fffffffLL) return "stem: too many tiles";
    alignas(64) CUtensorMap map_b, map_out, map_out_lo;
    const char* e;
    {
        const uint64_t kp = stem_packed_k(d.cin);
        const uint64_t dims[2] = {kp, 64};
        const uint64_t strides[1] = {kp};
        const uint32_t box[2] = {64, 64};
        if ((e = encode_bf16_map_public(&map_b, d.wpack, 2, dims, strides, box))) return e;
    }
    {
        const uint64_t dims[4] = {64, static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H),
                                  static_cast<uint64_t>(d.N)};
        const uint64_t strides[3] = {64, static_cast<uint64_t>(d.W) * 64, static_cast<uint64_t>(d.H) * d.W * 64};
        const uint32_t box[4] = {64, TILE_W, 2, 1};
        if ((e = encode_bf16_map_public(&map_out, d.dst, 4, dims, strides, box))) return e;
        map_out_lo = map_out;
        if (d.dst_lo && (e = encode_bf16_map_public(&map_out_lo, d.dst_lo, 4, dims, strides, box))) return e;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = static_cast<int>(tiles < 2LL * sms ? tiles : 2LL * sms);
    switch (d.cin) {
        case 1: return launch_stem<1>(d, map_b, map_out, map_out_lo, p, grid, stream);
        case 2: return launch_stem<2>(d, map_b, map_out, map_out_lo, p, grid, stream);
        case 3: return launch_stem<3>(d, map_b, map_out, map_out_lo, p, grid, stream);
        case 4: return launch_stem<4>(d, map_b, map_out, map_out_lo, p, grid, stream);
        case 5: return launch_stem<5>(d, map_b, map_out, map_out_lo, p, grid, stream);
        case 6: return launch_stem<6>(d, map_b, map_out, map_out_lo, p, grid, stream);
        case 7: return launch_stem<7>(d, map_b, map_out, map_out_lo, p, grid, stream);
        default: return launch_stem<8>(d, map_b, map_out, map_out_lo, p, grid, stream);
    }
}

}  // namespace fi
