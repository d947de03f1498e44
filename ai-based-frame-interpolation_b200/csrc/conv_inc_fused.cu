// The whole `inc` DoubleConv of the grey network (reference model/unet.py:12-17 twice, + the MaxPool2d of down1, :28)
// in ONE kernel: inc.double_conv.0 (C_in <= 2 -> 64, the stem) is computed on the fly for the halo of every output tile
// of inc.double_conv.3 (64 -> 64), so the 128 B/pixel intermediate tensor `inc.mid` is neither written to nor read from
// HBM (265 MB per 1080p pair each way) and one launch disappears. Same arithmetic as the two separate kernels —
// stem_mma.cu's bf16 hi/lo-split K = 64 MMA on the un-rounded normalised input, bias + ReLU, ONE bf16 rounding, then
// conv_halo.cu's nine tap-shifted-view MMAs per K step with resident weights — so the results are bit-identical.
//
// Tile = 8 (x) x 16 (y) output pixels = 128 GEMM rows (row = h*8 + w, one "column half" of conv_halo.cu's super tile).
// Its halo is 10 x 18 pixels of inc.mid = 180 rows of the stem GEMM = two M = 128 tiles (70 % filled). Shared memory:
//   halo[2]    2 x 36 KB   10 x 18 pixels at pitch 16 (slot = hy*16 + hx, 128 swizzled bytes per pixel), double buffered
//   conv W     72 KB       nine [64 x 64] tap slabs, resident        stem W   8 KB    one [64 x 64] hi/lo-split slab
//   stem A     2 x 16 KB   im2col rows of the two stem M tiles       input    4 KB    (8+4) x (16+4) raw pixels x C_in
// Warp roles (736 threads): 0..5 im2col producers (thread = one halo pixel: 4 warps for stem M tile 0, 2 for the 64 backed
// rows of M tile 1), 6 TMEM owner + MMA issuer + weight loads, 7..14 mid epilogue, one set per stem M tile (stem
// accumulator -> bias, ReLU, bf16 -> halo buffer; zero outside the image = the conv padding), 15..22 final epilogue in
// two sets that alternate tiles (conv accumulator -> bias, ReLU, bf16, 2x2 max pool -> TMA stores). TMEM: 2 x 64 conv
// columns + 2 sets x 2 M tiles x 64 stem columns. Both biases travel as kernel parameters (constant-bank operands).
// Measured (B200, four 1080p pairs): 0.89 ms serialised against 0.22 + 0.74 ms for the two launches at the same clock,
// +3 % frames/s for the whole forward at the sustained, power-capped clock (409 vs 397 frames/s, alternating runs of 60
// steps): 1 GB less HBM traffic per step. Six iterations, each read off the ncu source page (profiles/README.md):
//   1.19 ms  8 producer warps, one epilogue set: both epilogues busy ~3 000 cycles per tile, everything else waiting
//   1.14     two final-epilogue sets, biases in shared memory
//   1.06     two stem accumulator sets: the issuer no longer waits for the mid epilogue
//   0.98     one mid-epilogue set per stem M tile, 32-column epilogue passes (80 registers, no spills)
//   0.886    six producer warps, one halo pixel per thread: the issuer was waiting for the stem A tiles
//   0.889    biases as constant-bank operands: -36 % shared-memory wavefronts, same time — the limit is bytes, not
//            wavefronts: the issuer's barrier waits all succeed at once and every other role waits for it
// The tile period (~3 300 cycles) is now 88 % of the shared-memory-port floor: 264 KB of MMA operand reads + ~100 KB of
// row / halo / staging traffic per tile at 128 B/cycle = 2 900 cycles.
// The issuer runs one tile ahead with the stem: stem(t+1) is issued before conv(t), so the halo of tile t+1 is built
// (mid epilogue) while the tensor pipe works through the 36 MMAs of tile t.
#include "aux_kernels.cuh"
#include "conv_epilogue.cuh"
#include "conv_gemm.cuh"
#include "ptx.cuh"

#include <cstring>

namespace fi {

namespace {

constexpr int FT_W = 8, FT_H = 16;                       // output tile
constexpr int FH_W = FT_W + 2, FH_H = FT_H + 2;          // halo of inc.mid: 10 x 18
constexpr int FH_PITCH = 16;                             // halo slots per row
constexpr int FH_PIXELS = FH_W * FH_H;                   // 180 stem rows
constexpr int FH_BYTES = FH_PITCH * FH_H * 128;          // 36864
constexpr int FIN_W = FT_W + 4, FIN_H = FT_H + 4;        // raw input tile: 12 x 20
constexpr int FIN_PITCH = 13;
constexpr int F_THREADS = 23 * 32;
constexpr int F_PRODUCERS = 192;                         // 4 warps: stem M tile 0, 2 warps: the 64 backed rows of M tile 1
constexpr int F_SA1_ROWS = 64;                           // stem M tile 1 holds 52 valid rows: only 64 are backed by memory
constexpr int F_TMEM_COLS = 512;
constexpr int F_STEM_COL = 128;                          // TMEM columns [128, 384): two sets of the two stem M tiles

__host__ __device__ constexpr int fused_in_bytes(int cin) { return 2 * cin * FIN_H * FIN_PITCH * 4; }
__host__ __device__ constexpr int fused_smem_bytes(int cin) {
    return 1024 + 2 * FH_BYTES + 9 * 8192 + 8192 + (16384 + F_SA1_ROWS * 128) + 8 * (4096 + 1024) + 256 + 1040 + 512 +
           fused_in_bytes(cin);
}

struct FusedParams {
    PlaneSrc src[2];
    int N, H, W, tiles_x, tiles_y;
};
// Folded BatchNorm shifts by value: the epilogues read them as constant-bank operands of the FADDs (no shared-memory
// loads: the per-pixel bias reads were 36 % of the kernel's shared-memory wavefronts and slowed the MMA operand fetch).
struct FusedBias {
    float stem[64];
    float conv[64];
};

__device__ __forceinline__ float norm_u8_fused(uint8_t u) {
    return __fsub_rn(__fmul_rn(2.0f, __fdiv_rn(static_cast<float>(u), 255.0f)), 1.0f);
}
__device__ __forceinline__ uint32_t bf16_bits_fused(float v) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v), "f"(0.0f));
    return r & 0xffff0000u;
}
// K-major SWIZZLE_128B view of the halo buffer: 8-row groups are 8 consecutive pixels of one halo row, consecutive
// groups one halo row (FH_PITCH slots) apart.
__device__ __forceinline__ uint64_t fused_halo_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>((FH_PITCH * 128) >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

template <int CIN, bool U8>
__global__ void __launch_bounds__(F_THREADS, 1)
inc_fused_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ CUtensorMap map_stem_w,
                 const ConvKernelParams p, const FusedParams fp, const FusedBias fb) {
    constexpr int KT = 9 * CIN;
    static_assert(3 * KT <= 64, "fused inc kernel: the hi/lo-split stem row must fit one 64-element K slab");
    constexpr uint32_t IDESC = umma_idesc_bf16(128, 64);

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_halo = smem_base;
    const uint32_t smem_cw = smem_halo + 2 * FH_BYTES;
    const uint32_t smem_sw = smem_cw + 9 * 8192;
    const uint32_t smem_sa = smem_sw + 8192;
    // stem A tile 1 is backed by 64 rows only; the MMA's rows 64..127 read whatever follows (staging) into accumulator
    // lanes that nobody reads (halo pixels 180..255 do not exist)
    const uint32_t smem_stage = smem_sa + 16384 + F_SA1_ROWS * 128;
    const uint32_t smem_pool = smem_stage + 8 * 4096;
    const uint32_t smem_bar = smem_pool + 8 * 1024;
    const uint32_t bar_sa_full = smem_bar;            // 2: stem A tile mt built (4 producer warps each)
    const uint32_t bar_sa_empty = smem_bar + 16;      // 2: stem MMAs of tile mt retired
    const uint32_t bar_st_full = smem_bar + 32;       // 2: stem accumulator set complete
    const uint32_t bar_st_empty = smem_bar + 48;      // 2: mid epilogue has read it (4 warps)
    const uint32_t bar_h_full = smem_bar + 64;        // 2: halo buffer written (4 mid-epilogue warps)
    const uint32_t bar_h_empty = smem_bar + 80;       // 2: conv MMAs reading it retired
    const uint32_t bar_t_full = smem_bar + 96;        // 2: conv accumulator complete
    const uint32_t bar_t_empty = smem_bar + 112;      // 2: final epilogue has read it (4 warps)
    const uint32_t bar_w = smem_bar + 128;            // 1: weights resident
    const uint32_t tmem_slot = smem_bar + 136;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    uint32_t* lut = reinterpret_cast<uint32_t*>(smem_raw + (smem_bar + 256 - smem_u32(smem_raw)));
    uint32_t* in_tile = lut + 260 + 128;   // 128 words after the table are unused (they held the biases, now parameters)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (U8 && threadIdx.x < 256) {
        const float v = norm_u8_fused(static_cast<uint8_t>(threadIdx.x));
        const uint32_t h = bf16_bits_fused(v);
        lut[threadIdx.x] = (h >> 16) | bf16_bits_fused(v - __uint_as_float(h));
        if (threadIdx.x == 0) lut[256] = 0u;
    }
    if (threadIdx.x == 192) {
        tma_prefetch_desc(&maps.b);
        tma_prefetch_desc(&map_stem_w);
        tma_prefetch_desc(&maps.out[0]);
        tma_prefetch_desc(&maps.pool[0]);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_sa_full + 8 * i, i == 0 ? 4 : 2);
            mbar_init(bar_sa_empty + 8 * i, 1);
            mbar_init(bar_h_full + 8 * i, 8);
            mbar_init(bar_h_empty + 8 * i, 1);
            mbar_init(bar_t_full + 8 * i, 1);
            mbar_init(bar_t_empty + 8 * i, 4);
            mbar_init(bar_st_full + 8 * i, 1);
            mbar_init(bar_st_empty + 8 * i, 8);
        }
        mbar_init(bar_w, 1);
        fence_mbar_init();
    }
    if (warp == 6) tmem_alloc(tmem_slot, F_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_launch_dependents();
    pdl_wait();

    const int per_img = fp.tiles_y * fp.tiles_x;
    const int total_tiles = fp.N * per_img;
    auto tile_origin = [&](int t, int& img, int& y0, int& x0) {
        img = t / per_img;
        const int r = t - img * per_img;
        y0 = (r / fp.tiles_x) * FT_H;
        x0 = (r % fp.tiles_x) * FT_W;
    };

    if (warp < 6) {
        // ------------------------------------------------------------ im2col producers: thread = one halo pixel (stem row).
        // Warps 0..3 build stem M tile 0, warps 4..5 the 64 backed rows of M tile 1 (52 of them are halo pixels).
        const int pth = threadIdx.x;           // 0..191
        const int mt = pth >> 7, m = pth & 127;
        const int hpix = mt * 128 + m;         // halo pixel of this row, valid below FH_PIXELS
        constexpr int PLANE = FIN_H * FIN_PITCH;
        constexpr int IN_ELEMS = CIN * FIN_H * FIN_W;
        constexpr int NLOAD = (IN_ELEMS + F_PRODUCERS - 1) / F_PRODUCERS;
        uint32_t raw[NLOAD];
        auto fetch = [&](int t) {
            int img, y0, x0;
            tile_origin(t, img, y0, x0);
#pragma unroll
            for (int k = 0; k < NLOAD; ++k) {
                const int i = pth + F_PRODUCERS * k;
                uint32_t v = U8 ? 256u : 0u;   // out of bounds -> zero padding of the stem conv
                if (i < IN_ELEMS) {
                    const int c = i / (FIN_H * FIN_W);
                    const int rr = (i - c * FIN_H * FIN_W) / FIN_W;
                    const int col = i - c * FIN_H * FIN_W - rr * FIN_W;
                    const int yy = y0 - 2 + rr, xx = x0 - 2 + col;
                    if (yy >= 0 && yy < fp.H && xx >= 0 && xx < fp.W) {
                        const bool first = c < fp.src[0].channels;
                        const PlaneSrc& sp = first ? fp.src[0] : fp.src[1];
                        const int cc = first ? c : c - fp.src[0].channels;
                        const long long off = img * sp.batch_stride + cc * sp.chan_stride + yy * sp.row_stride +
                                              xx * sp.px_stride;
                        if (U8) v = __ldg(static_cast<const uint8_t*>(sp.ptr) + off);
                        else v = __float_as_uint(__ldg(static_cast<const float*>(sp.ptr) + off));
                    }
                }
                raw[k] = v;
            }
        };
        int it = 0;
        if (static_cast<int>(blockIdx.x) < total_tiles) fetch(blockIdx.x);
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            uint32_t* tile = in_tile + (it & 1) * CIN * PLANE;
#pragma unroll
            for (int k = 0; k < NLOAD; ++k) {
                const int i = pth + F_PRODUCERS * k;
                if (i < IN_ELEMS) {
                    const int c = i / (FIN_H * FIN_W);
                    const int rr = (i - c * FIN_H * FIN_W) / FIN_W;
                    const int col = i - c * FIN_H * FIN_W - rr * FIN_W;
                    uint32_t packed;
                    if (U8) {
                        packed = lut[raw[k]];
                    } else {
                        const float v = __uint_as_float(raw[k]);
                        const uint32_t h = bf16_bits_fused(v);
                        packed = (h >> 16) | bf16_bits_fused(v - __uint_as_float(h));
                    }
                    tile[c * PLANE + rr * FIN_PITCH + col] = packed;
                }
            }
            if (t + static_cast<int>(gridDim.x) < total_tiles) fetch(t + gridDim.x);
            // two input buffers: one barrier per tile (a thread re-writes buffer b only after every producer passed the
            // barrier of the tile in between, i.e. after all of them finished reading b)
            asm volatile("bar.sync 1, 192;" ::: "memory");
            uint32_t hl[KT];   // low half = bf16 hi part, high half = bf16 lo part of the normalised input
#pragma unroll
            for (int e = 0; e < KT; ++e) hl[e] = 0u;
            if (hpix < FH_PIXELS) {
                const int hy = hpix / FH_W, hx = hpix - hy * FH_W;
                const uint32_t* px = tile + hy * FIN_PITCH + hx;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                    for (int c = 0; c < CIN; ++c) hl[tap * CIN + c] = px[c * PLANE + (tap / 3) * FIN_PITCH + (tap % 3)];
            }
            auto elem = [&](int e) -> uint32_t {   // 16-bit K element e of the row [x_hi | x_hi | x_lo | 0]
                return e < KT ? (hl[e] & 0xffffu)
                              : (e < 2 * KT ? (hl[e - KT] & 0xffffu) : (e < 3 * KT ? (hl[e - 2 * KT] >> 16) : 0u));
            };
            mbar_wait(bar_sa_empty + 8 * mt, (it & 1) ^ 1);
            const uint32_t row = smem_sa + mt * 16384 + m * 128;   // M tile 1: m < 64 (its producers are warps 4, 5)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t wv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int e = j * 8 + q * 2;
                    wv[q] = elem(e) | (elem(e + 1) << 16);
                }
                st_shared_v4(row + ((j ^ (m & 7)) << 4), wv[0], wv[1], wv[2], wv[3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_sa_full + 8 * mt);
        }
    } else if (warp == 6) {
        // ------------------------------------------------------------ weights (once) + MMA issue
        if (elect_one()) {
            mbar_expect_tx(bar_w, 10 * 8192);
            for (int tap = 0; tap < 9; ++tap) tma_load_2d(smem_cw + tap * 8192, &maps.b, bar_w, tap * BLOCK_K, 0);
            tma_load_2d(smem_sw, &map_stem_w, bar_w, 0, 0);
        }
        __syncwarp();
        mbar_wait(bar_w, 0);
        int stems = 0;
        auto issue_stem = [&]() {
            const uint32_t par = stems & 1;     // stem A tiles: one use per tile
            const int sset = stems & 1;         // stem accumulators: two sets, so the issuer never waits for the mid
            mbar_wait(bar_st_empty + 8 * sset, ((stems >> 1) & 1) ^ 1);   // epilogue of the tile just before this one
            tc_fence_after();
            const uint64_t db = umma_desc_sw128(smem_sw);
            for (int smt = 0; smt < 2; ++smt) {
                mbar_wait(bar_sa_full + 8 * smt, par);
                tc_fence_after();
                const uint64_t da = umma_desc_sw128(smem_sa + smt * 16384);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_base + F_STEM_COL + sset * 128 + smt * 64, da + 2 * k, db + 2 * k, IDESC, k != 0);
                    umma_commit(bar_sa_empty + 8 * smt);
                    if (smt == 1) umma_commit(bar_st_full + 8 * sset);
                }
                __syncwarp();
            }
            ++stems;
        };
        int it = 0;
        if (static_cast<int>(blockIdx.x) < total_tiles) issue_stem();
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            if (t + static_cast<int>(gridDim.x) < total_tiles) issue_stem();   // one tile ahead of the conv
            const int acc = it & 1;
            const uint32_t par2 = (it >> 1) & 1;
            mbar_wait(bar_t_empty + 8 * acc, par2 ^ 1);
            mbar_wait(bar_h_full + 8 * acc, par2);
            tc_fence_after();
            const uint32_t a_base = smem_halo + acc * FH_BYTES;
            const uint32_t d_tmem = tmem_base + acc * 64;
            for (int tap = 0; tap < 9; ++tap) {
                const int dy = tap / 3, dx = tap - 3 * dy;
                const uint64_t da = fused_halo_desc(a_base + (dy * FH_PITCH + dx) * 128);
                const uint64_t db = umma_desc_sw128(smem_cw + tap * 8192);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, IDESC, (tap | k) != 0);
                    if (tap == 8) {
                        umma_commit(bar_h_empty + 8 * acc);
                        umma_commit(bar_t_full + 8 * acc);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp < 15) {
        // ------------------------------------------------------------ mid epilogue, warps 7..14: set (warp-7)/4 owns stem
        // M tile `smt`; 32 accumulator columns at a time (register budget: 80 per thread with 21 warps)
        const int q = warp & 3;
        const int smt = (warp - 7) >> 2;
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            int img, y0, x0;
            tile_origin(t, img, y0, x0);
            const int hs = it & 1;
            mbar_wait(bar_st_full + 8 * hs, (it >> 1) & 1);
            mbar_wait(bar_h_empty + 8 * hs, ((it >> 1) & 1) ^ 1);   // conv MMAs of the tile before last retired
            tc_fence_after();
            const uint32_t hbuf = smem_halo + hs * FH_BYTES;
            const int hpix = smt * 128 + q * 32 + lane;
            if (smt * 128 + q * 32 < FH_PIXELS) {                   // warp-uniform: this lane quarter holds halo pixels
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + F_STEM_COL + hs * 128 + smt * 64;
                const int hy = hpix / FH_W, hx = hpix - hy * FH_W;
                const int yy = y0 - 1 + hy, xx = x0 - 1 + hx;
                const bool valid = hpix < FH_PIXELS;
                const bool inside = yy >= 0 && yy < fp.H && xx >= 0 && xx < fp.W;   // outside: the conv's zero padding
                const int slot = hy * FH_PITCH + hx;
                const uint32_t rowaddr = hbuf + slot * 128;
#pragma unroll
                for (int half = 0; half < 2; ++half) {   // unrolled: the bias indices below are compile-time constants
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + 32 * half, v);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {   // 16-byte chunk j = channels 8j .. 8j+7
                            const int j = 4 * half + jj;
                            const int o = jj * 8;
                            uint32_t hw[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                hw[e] = pack_bf16x2(fmaxf(__uint_as_float(v[o + 2 * e]) + fb.stem[8 * j + 2 * e], 0.f),
                                                    fmaxf(__uint_as_float(v[o + 2 * e + 1]) + fb.stem[8 * j + 2 * e + 1], 0.f));
                            if (!inside) hw[0] = hw[1] = hw[2] = hw[3] = 0u;
                            st_shared_v4(rowaddr + ((j ^ (slot & 7)) << 4), hw[0], hw[1], hw[2], hw[3]);
                        }
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_st_empty + 8 * hs);
                mbar_arrive(bar_h_full + 8 * hs);
            }
        }
    } else {
        // ------------------------------------------------------------ final epilogue warps 15..22: two sets, set s owns
        // the tiles with (iteration & 1) == s and therefore conv accumulator s. Same arithmetic as epilogue_chunk_halo
        // (bias, ReLU, one bf16 rounding, 2x2 max of the rounded values), 32 columns at a time.
        const int q = warp & 3;
        const int ew = warp - 15;
        const int set = ew >> 2;
        const uint32_t sbuf = smem_stage + ew * 4096;
        const uint32_t pbuf = smem_pool + ew * 1024;
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            if ((it & 1) != set) continue;
            int img, y0, x0;
            tile_origin(t, img, y0, x0);
            const int acc = set;
            mbar_wait(bar_t_full + 8 * acc, (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 64;
            if (elect_one()) tma_store_wait_read<0>();   // the previous tile's stores have read the staging tiles
            __syncwarp();
            const uint32_t row = sbuf + lane * 128;      // lane = (row in 0..3) * 8 + column
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(taddr + 32 * half, v);
                tmem_ld_wait();
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = 4 * half + jj;
                    const int o = jj * 8;
                    uint32_t hw[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        hw[e] = pack_bf16x2(fmaxf(__uint_as_float(v[o + 2 * e]) + fb.conv[8 * j + 2 * e], 0.f),
                                            fmaxf(__uint_as_float(v[o + 2 * e + 1]) + fb.conv[8 * j + 2 * e + 1], 0.f));
                    st_shared_v4(row + ((j ^ (lane & 7)) << 4), hw[0], hw[1], hw[2], hw[3]);
                }
            }
            // every TMEM read of this accumulator is complete: hand it back to the MMA warp before the stores
            tc_fence_before();
            __syncwarp();
            if (elect_one()) mbar_arrive(bar_t_empty + 8 * acc);
            fence_proxy_async_smem();
            __syncwarp();
            const int yq = y0 + 4 * q;
            if (elect_one()) tma_store_4d(&maps.out[0], sbuf, 0, x0, yq, img);   // box {64, 8, 4, 1}
            // pooled 2 rows x 4 columns: max over lanes {2ph*8 + 2pw, +1, +8, +9} (bf16 max commutes with the rounding)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int pp = lane >> 2;  // pooled pixel 0..7 = ph*4 + pw
                const int j = (lane & 3) * 2 + i;
                const int r0 = (pp >> 2) * 16 + (pp & 3) * 2;
                const int r1 = r0 + 1, r2 = r0 + 8, r3 = r0 + 9;
                const uint4 m0 = ld_shared_v4(sbuf + r0 * 128 + ((j ^ (r0 & 7)) << 4));
                const uint4 m1 = ld_shared_v4(sbuf + r1 * 128 + ((j ^ (r1 & 7)) << 4));
                const uint4 m2 = ld_shared_v4(sbuf + r2 * 128 + ((j ^ (r2 & 7)) << 4));
                const uint4 m3 = ld_shared_v4(sbuf + r3 * 128 + ((j ^ (r3 & 7)) << 4));
                uint4 m;
                m.x = bf16x2_max(bf16x2_max(m0.x, m1.x), bf16x2_max(m2.x, m3.x));
                m.y = bf16x2_max(bf16x2_max(m0.y, m1.y), bf16x2_max(m2.y, m3.y));
                m.z = bf16x2_max(bf16x2_max(m0.z, m1.z), bf16x2_max(m2.z, m3.z));
                m.w = bf16x2_max(bf16x2_max(m0.w, m1.w), bf16x2_max(m2.w, m3.w));
                st_shared_v4(pbuf + pp * 128 + ((j ^ (pp & 7)) << 4), m.x, m.y, m.z, m.w);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one()) {
                tma_store_4d(&maps.pool[0], pbuf, 0, x0 >> 1, yq >> 1, img);   // box {64, 4, 2, 1}
                tma_store_commit();
            }
        }
        __syncwarp();
        if (elect_one()) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 6) {
        tc_fence_after();
        tmem_dealloc(tmem_base, F_TMEM_COLS);
    }
}

template <int CIN, bool U8>
const char* launch_fused_inst(const ConvLaunch& conv, const CUtensorMap& map_stem_w, const FusedParams& fp,
                              const FusedBias& fb, int grid, cudaStream_t stream) {
    auto k = inc_fused_kernel<CIN, U8>;
    constexpr int smem = fused_smem_bytes(CIN);
    static_assert(smem <= 232448, "fused inc kernel exceeds the 227 KB shared memory limit");
    static std::atomic<uint64_t> configured{0};
    if (!smem_opt_in(k, smem, configured)) return "inc(fused): cudaFuncSetAttribute failed";
    const cudaError_t e = launch_kernel(k, dim3(grid), dim3(F_THREADS), smem, stream, conv.maps, map_stem_w, conv.p, fp, fb);
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace

bool inc_fused_eligible(int cin, const ConvLaunch& conv) {
    return cin >= 1 && cin <= 2 && conv.halo && !conv.pair && !conv.split && conv.mode == EPI_STORE_POOL &&
           conv.block_n == 64 && conv.p.slabs == 1;
}

// d: the stem description (dst unused); conv: the prepared launch of inc.double_conv.3 (its A tensor map is not used).
// host_bias: the folded shifts of both layers on the HOST, stem[64] then conv[64] (they travel as kernel parameters).
const char* inc_fused_launch(const StemDesc& d, const ConvLaunch& conv, const float* host_bias, int n_img, int num_sms,
                             cudaStream_t stream) {
    if (!inc_fused_eligible(d.cin, conv)) return "inc(fused): layer is not eligible";
    if (d.src[0].channels + d.src[1].channels != d.cin) return "inc(fused): plane sources do not add up to cin";
    if (!d.wpack || !host_bias || !d.src[0].ptr) return "inc(fused): null operand";
    FusedBias fb;
    memcpy(fb.stem, host_bias, sizeof fb.stem);
    memcpy(fb.conv, host_bias + 64, sizeof fb.conv);
    FusedParams fp;
    memset(&fp, 0, sizeof fp);
    fp.src[0] = d.src[0];
    fp.src[1] = d.src[1];
    fp.N = n_img;
    fp.H = d.H;
    fp.W = d.W;
    fp.tiles_x = (d.W + FT_W - 1) / FT_W;
    fp.tiles_y = (d.H + FT_H - 1) / FT_H;
    const long long tiles = static_cast<long long>(n_img) * fp.tiles_x * fp.tiles_y;
    if (tiles > 0x7fffffffLL) return "inc(fused): too many tiles";
    alignas(64) CUtensorMap map_w;
    {
        const uint64_t kp = stem_packed_k(d.cin);
        const uint64_t dims[2] = {kp, 64};
        const uint64_t strides[1] = {kp};
        const uint32_t box[2] = {64, 64};
        const char* e = encode_bf16_map_public(&map_w, d.wpack, 2, dims, strides, box);
        if (e) return e;
    }
    const int grid = static_cast<int>(tiles < num_sms ? tiles : num_sms);
    if (d.cin == 1) return d.is_u8 ? launch_fused_inst<1, true>(conv, map_w, fp, fb, grid, stream)
                                   : launch_fused_inst<1, false>(conv, map_w, fp, fb, grid, stream);
    return d.is_u8 ? launch_fused_inst<2, true>(conv, map_w, fp, fb, grid, stream)
                   : launch_fused_inst<2, false>(conv, map_w, fp, fb, grid, stream);
}

}  // namespace fi
