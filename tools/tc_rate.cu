// MMA issue-rate probe (mode 0: MMAs only; 1: + a commit per 12 MMAs; 2: + two barrier probes and a fence): one CTA per SM loops tcgen05.mma kind::tf32 (M=128, K=8) over fixed
// shared-memory operands, N = 128 or 256, and reports cycles per MMA and TF32 TFLOP/s.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) rate(int iters, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t bar2[4], bar3;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (16384 * 2 + N * 128 * 2) / 4; e += 128) ((float*)base)[e] = 0.001f * (e & 255);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[i])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(base), b0 = a0 + 32768;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t acc = tb + (uint32_t)((it & 1) * N);
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = make_desc(a0 + (p == 2 ? 16384 : 0) + k * 32);
                    const uint64_t bd = make_desc(b0 + (p == 1 ? N * 128 : 0) + k * 32);
                    const uint32_t en = (p | k) ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(acc),
                        "l"(ad), "l"(bd), "r"(idesc), "r"(en)
                        : "memory");
                }
            if (MODE >= 1)  // per-tile commit, as the scan does
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[it & 3])) : "memory");
            if (MODE >= 2) {  // plus two barrier probes that succeed at once and the fence
                for (int w = 0; w < 2; ++w) {
                    uint32_t ok = 0;
                    while (!ok)
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                                     : "=r"(ok) : "r"(smem_u32(&bar3)), "r"(1) : "memory");
                }
                asm volatile("tcgen05.fence::after_thread_sync;");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        if (blockIdx.x == 0) cycles[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(512));
}
template <int N, int MODE>
void run() {
    long long* dc; cudaMalloc(&dc, 8);
    const int smem = 1024 + 32768 + 2 * N * 128, iters = 4000;
    cudaFuncSetAttribute(rate<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate<N, MODE><<<148, 128, smem>>>(10, dc); cudaDeviceSynchronize();
    cudaEventRecord(e0); rate<N, MODE><<<148, 128, smem>>>(iters, dc); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    const double flop = 2.0 * 128 * N * 8 * 12 * (double)iters * 148;
    printf("mode %d N=%d: %.1f cycles per MMA (M128 N%d K8), %.1f TFLOP/s TF32 over 148 SMs, %s\n", MODE, N, (double)c / (12.0 * iters), N,
           flop / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() { run<128, 0>(); run<128, 1>(); run<128, 2>(); run<256, 2>(); return 0; }
