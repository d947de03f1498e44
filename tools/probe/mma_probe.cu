// Micro-probe: pure tcgen05.mma issue rate (SS mode, bf16, M=128 per CTA) for N = 64/128/256 with cta_group::1 and
// cta_group::2, operands resident in smem (garbage), no TMA and no waits inside the loop. Tells whether the narrow
// layers' ceiling is a per-instruction cost of the 1-CTA path.   nvcc -arch=sm_100a -o mma_probe mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../ai-based-frame-interpolation_b200/csrc/ptx.cuh"
using namespace fi;

__device__ __forceinline__ void umma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc),
                 "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo) {   // MN-major SWIZZLE_128B (wgrad_gemm.cu)
    return static_cast<uint64_t>((addr >> 4) & 0x3FFF) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16) |
           (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
}

// MAJOR: 0 = both operands K-major, 1 = both MN-major, 2 = A MN-major / B K-major
template <int N, int CG, int MAJOR = 0>
__global__ void __launch_bounds__(128, 1) probe(int iters, int n_acc, unsigned long long* cycles) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t sa = base, sb = base + 4 * 16384, bar = sb + 4 * 32768, slot = bar + 16;
    uint32_t* slot_p = reinterpret_cast<uint32_t*>(raw + (slot - smem_u32(raw)));
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = CG == 2 ? cluster_rank() : 0;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) {
        if (CG == 1) tmem_alloc(slot, 512);
        else if (rank == 0 || true) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync();
    tc_fence_after();
    const uint32_t tmem = *slot_p;
    constexpr uint32_t IDESC = umma_idesc_bf16(CG == 2 ? 256 : 128, N) | (MAJOR >= 1 ? (1u << 15) : 0u) |
                               (MAJOR == 1 ? (1u << 16) : 0u);
    long long t0 = clock64();
    if (warp == 1 && rank == 0) {
        if (elect_one()) {
            for (int i = 0; i < iters; ++i) {
                const int st = i & 3;
                const uint64_t da = MAJOR >= 1 ? desc_mn(sa + st * 16384, 8192) : umma_desc_sw128(sa + st * 16384);
                const uint64_t db = MAJOR == 1 ? desc_mn(sb + st * 32768, 8192) : umma_desc_sw128(sb + st * 32768);
                const uint32_t ka = MAJOR >= 1 ? 128 : 2, kb = MAJOR == 1 ? 128 : 2;   // descriptor step per K = 16
                const uint32_t d = tmem + (i % n_acc) * N;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (CG == 1) umma_bf16_ss(d, da + ka * k, db + kb * k, IDESC, 1);
                    else umma2(d, da + ka * k, db + kb * k, IDESC, 1);
                }
            }
            if (CG == 1) umma_commit(bar);
            else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
        }
        __syncwarp();
    }
    if (warp == 1) { mbar_wait(bar, 0); }
    long long t1 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) *cycles = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync();
    if (warp == 0) {
        if (CG == 1) tmem_dealloc(tmem, 512);
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int N, int CG, int MAJOR = 0>
void run(int n_acc) {
    const int iters = 4000, smem = 1024 + 4 * 16384 + 4 * 32768 + 64;
    auto k = probe<N, CG, MAJOR>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    unsigned long long* cyc; cudaMalloc(&cyc, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int sms = 148;
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(CG == 2 ? sms : sms); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {unsigned(CG), 1, 1};
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a);
        cudaError_t e = cudaLaunchKernelEx(&cfg, k, iters, n_acc, cyc);
        cudaEventRecord(b); cudaEventSynchronize(b);
        if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("N=%d CG=%d launch failed: %s\n", N, CG, cudaGetErrorString(e)); return; }
    }
    float ms; cudaEventElapsedTime(&ms, a, b);
    unsigned long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // per CTA-pair (CG=2) the MMA is 256 x N x 16; CTAs that issue: sms (CG=1) or sms/2 leaders (CG=2)
    const double flops = 2.0 * (CG == 2 ? 256 : 128) * N * 16 * 4.0 * iters * (CG == 2 ? sms / 2 : sms);
    printf("major=%d N=%3d cta_group=%d accs=%d: %.3f ms  %.0f TFLOP/s  %.1f cycles per MMA (clock64)\n", MAJOR, N, CG, n_acc, ms, flops / ms / 1e9,
           double(h) / (4.0 * iters));
}

int main() {
    run<64, 1>(1); run<64, 1>(2); run<64, 1>(4); run<128, 1>(1); run<128, 1>(2); run<256, 1>(1); run<256, 1>(2);
    run<64, 2>(1); run<64, 2>(2); run<128, 2>(2); run<256, 2>(2);
    run<64, 1, 1>(2); run<128, 1, 1>(2); run<192, 1, 1>(2); run<256, 1, 1>(2);      // wgrad: both operands MN-major
    run<64, 1, 2>(2); run<256, 1, 2>(2);                                           // A MN-major only
    run<256, 2, 1>(2);
    return 0;
}
