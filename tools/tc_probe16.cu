// Probe: D[128 x 128] = A[128 x 32] . B[128 x 32]^T with tcgen05.mma kind::f16 (FP16 hi/lo split,
// FP32 accumulate), operands K-major SWIZZLE_64B (64-byte rows = 32 halves).  Validates layout,
// descriptors and accuracy against float64 for inputs scaled to |v| <= 64.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <vector>
#include <random>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // K-major, SWIZZLE_64B: LBO = 1, SBO = 512 B between 8-row groups, version 1, layout type 4
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)4 << 61);
}
constexpr int M = 128, N = 128, K = 32;
__global__ void __launch_bounds__(128, 1) probe(const float* __restrict__ A, const float* __restrict__ B,
                                                float* __restrict__ D, int passes) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    __half* Ah = (__half*)base;                 // 128 rows x 64 B = 8 KB per image
    __half* Al = (__half*)(base + 8192);
    __half* Bh = (__half*)(base + 16384);
    __half* Bl = (__half*)(base + 24576);
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < M * K; e += 128) {
        const int r = e / K, k = e % K;
        // 8-row groups of 512 B; 16-byte chunk c (8 halves) of row r stored at chunk c ^ ((r >> 1) & 3)
        const int chunk = k >> 3, within = k & 7;
        const int off = ((r >> 3) * 512 + (r & 7) * 64 + (((chunk ^ ((r >> 1) & 3)) & 3) << 4) + within * 2) >> 1;
        const float a = A[e], b = B[e];
        const __half ah = __float2half_rn(a), bh = __float2half_rn(b);
        Ah[off] = ah;
        Al[off] = __float2half_rn(a - __half2float(ah));
        Bh[off] = bh;
        Bl[off] = __float2half_rn(b - __half2float(bh));
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        // c_format F32 (1), a/b format F16 (0), K-major, N, M
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a_addr[3] = {smem_u32(Ah), smem_u32(Ah), smem_u32(Al)};
        const uint32_t b_addr[3] = {smem_u32(Bh), smem_u32(Bl), smem_u32(Bh)};
        int first = 1;
        for (int p = 0; p < passes; ++p)
            for (int k = 0; k < K / 16; ++k) {   // K = 16 halves = 32 B per MMA
                const uint64_t ad = make_desc(a_addr[p] + k * 32), bd = make_desc(b_addr[p] + k * 32);
                const uint32_t acc = first ? 0u : 1u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tb),
                             "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
                first = 0;
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int row = warp * 32 + lane;
#pragma unroll
        for (int j = 0; j < 32; ++j) D[row * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(128));
}
int main() {
    std::mt19937 rng(1);
    std::normal_distribution<float> nd(0.f, 8.f);
    std::vector<float> A(M * K), B(N * K), D(M * N);
    for (auto& v : A) v = nd(rng);
    for (auto& v : B) v = nd(rng);
    for (int i = 0; i < 40; ++i) { A[i * 7] *= 1e-3f; B[i * 11] *= 1e-4f; }   // some tiny values (FP16 subnormal lo parts)
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const int smem = 4 * 8192 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int passes = 1; passes <= 3; passes += 2) {
        cudaMemset(dD, 0, D.size() * 4);
        probe<<<1, 128, smem>>>(dA, dB, dD, passes);
        cudaError_t e = cudaDeviceSynchronize();
        printf("passes=%d: %s\n", passes, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0, worst_rel = 0; int bad = 0;
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < N; ++j) {
                double ref = 0, na = 0, nb = 0;
                for (int k = 0; k < K; ++k) { ref += (double)A[i * K + k] * B[j * K + k]; na += (double)A[i*K+k]*A[i*K+k]; nb += (double)B[j*K+k]*B[j*K+k]; }
                const double err = fabs(D[i * N + j] - ref), rel = err / (sqrt(na) * sqrt(nb));
                if (err > worst) worst = err;
                if (rel > worst_rel) worst_rel = rel;
                if (rel > 1e-2) ++bad;
            }
        printf("  max |err| = %.3e, max err/(|a||b|) = %.3e = 2^%.1f, gross mismatches %d\n", worst, worst_rel, log2(worst_rel), bad);
    }
    return 0;
}
