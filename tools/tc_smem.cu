// Shared-memory bandwidth probe for the tensor-core scan: one CTA per SM issues the scan's 6 FP16 MMAs per
// tile (M = 128, N = 128, K = 16) with the A operand from shared memory (SS) or from tensor memory (TS),
// alone, next to a thread streaming 16 KB bulk copies into shared memory (the candidate ring), and next to
// four warps of LDS/STS traffic (the selection warps' queues and heaps).  Timing only: operands are junk.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // K-major, SWIZZLE_64B (the scan's layout)
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)4 << 61);
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t par) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(par) : "memory");
}
// TS: A from tensor memory.  LOAD bit 0: bulk copies, bit 1: LDS/STS warps.
template <int N, int TS, int LOAD, int SYNC>
__global__ void __launch_bounds__(192, 1) rate(int iters, long long* cycles, const unsigned char* src) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    // A hi/lo 16 KB | B hi/lo N*128 B | copy ring 4 x 16 KB | scratch 32 KB
    unsigned char* ring = base + 16384 + N * 128;
    unsigned char* scratch = ring + 65536;
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t bar, cbar[4], tb4[4], okbar;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (16384 + N * 128) / 2; e += 192) ((unsigned short*)base)[e] = (unsigned short)(0x2000 + (e & 255));
    if (tid == 0) {
        stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[i])));
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tb4[i])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&okbar)));
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
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(base), b0 = a0 + 16384;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t acc = tb + (uint32_t)((it & 1) * N);
            // SYNC bit 1: two barrier probes that succeed at once (parity of the phase before the first), bit 2: the fence
            if (SYNC & 2) { wait_bar(&okbar, 1); wait_bar(&okbar, 1); }
            if (SYNC & 4) asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint64_t ad = make_desc(a0 + (p == 2 ? 8192 : 0) + k * 32);
                    const uint64_t bd = make_desc(b0 + (p == 1 ? N * 64 : 0) + k * 32);
                    const uint32_t at = tb + 2 * N + (p == 2 ? 16 : 0) + k * 8;   // TS: 8 columns per K = 16 halves
                    const uint32_t en = (p | k) ? 1u : 0u;
                    if (TS)
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(acc),
                            "r"(at), "l"(bd), "r"(idesc), "r"(en)
                            : "memory");
                    else
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(acc),
                            "l"(ad), "l"(bd), "r"(idesc), "r"(en)
                            : "memory");
                }
            // SYNC bit 0: a commit per tile, bit 3: a second one
            if (SYNC & 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&tb4[it & 3])) : "memory");
            if (SYNC & 8) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&tb4[(it + 2) & 3])) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        wait_bar(&bar, 0);
        if (blockIdx.x == 0) cycles[0] = clock64() - t0;
        stop = 1;
    } else if (warp == 1) {
        if ((LOAD & 1) && (tid & 31) == 0) {   // the candidate ring: 16 KB per tile, 4 copies in flight
            long long n = 0;
            while (!stop) {
                const int s = (int)(n & 3);
                if (n >= 4) wait_bar(&cbar[s], (uint32_t)((n >> 2) - 1) & 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar[s])), "r"(16384) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(ring + s * 16384)),
                             "l"(src + ((size_t)blockIdx.x * 4096 + (n & 4095)) * 16384), "r"(16384), "r"(smem_u32(&cbar[s]))
                             : "memory");
                ++n;
            }
            if (blockIdx.x == 0) cycles[1] = n;
            for (long long m = (n > 4 ? n - 4 : 0); m < n; ++m) wait_bar(&cbar[m & 3], (uint32_t)(m >> 2) & 1);
        }
    } else {
        if (LOAD & 4) {   // the selection warps' read-out: 4 x 32 columns of this warp's 32 lanes, over and over
            uint32_t acc = 0;
            const uint32_t trow = tb + ((uint32_t)((warp & 3) * 32) << 16);
            while (!__shfl_sync(0xFFFFFFFFu, (int)stop, 0)) {   // warp-uniform exit: the loads are .sync.aligned
                uint32_t r[32];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                                 : "r"(trow + q * 32));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc ^= r[j];
                }
            }
            if (acc == 0x12345) cycles[3] = acc;
        } else if (LOAD & 2) {   // selection-like traffic: per thread one 8-byte load + store per ~8 instructions
            uint2* my = (uint2*)scratch + (tid - 64);
            uint2 v = make_uint2(tid, 0);
            while (!stop) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint2 r = my[j * 128];
                    v.x += r.x * 3 + j;
                    v.y ^= r.y + v.x;
                    my[((j + 5) & 15) * 128] = v;
                }
            }
            if (v.x == 0x12345) cycles[2] = v.y;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "n"(512));
}
template <int N, int TS, int LOAD, int SYNC>
void run(const unsigned char* src) {
    long long* dc; cudaMalloc(&dc, 32); cudaMemset(dc, 0, 32);
    const int smem = 1024 + 16384 + N * 128 + 65536 + 32768, iters = 20000;
    cudaFuncSetAttribute(rate<N, TS, LOAD, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate<N, TS, LOAD, SYNC><<<148, 192, smem>>>(10, dc, src); cudaDeviceSynchronize();
    cudaEventRecord(e0); rate<N, TS, LOAD, SYNC><<<148, 192, smem>>>(iters, dc, src); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c[4]; cudaMemcpy(c, dc, 32, cudaMemcpyDeviceToHost);
    const double flop = 2.0 * 128 * N * 16 * 6 * (double)iters * 148;
    printf("N=%d A from %s, load %d, sync %d: %.1f cycles per tile of 6 MMAs (ideal %d), %.0f TFLOP/s, copies per tile %.2f, %s\n", N,
           TS ? "TMEM" : "smem", LOAD, SYNC, (double)c[0] / iters, 3 * N, flop / ms / 1e9, (double)c[1] / iters,
           cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
    cudaFree(dc);
}
int main() {
    unsigned char* src; cudaMalloc(&src, (size_t)148 * 4096 * 16384); cudaMemset(src, 0, (size_t)148 * 4096 * 16384);
    run<128, 0, 0, 0>(src); run<128, 0, 1, 0>(src); run<128, 0, 3, 0>(src); run<128, 1, 3, 0>(src);
    // what the scan's issuing loop adds per tile: commit(s), two barrier probes, the fence; and the read-out next to it
    run<128, 0, 0, 1>(src); run<128, 0, 0, 9>(src); run<128, 0, 0, 2>(src); run<128, 0, 0, 4>(src); run<128, 0, 0, 7>(src);
    run<128, 0, 0, 15>(src); run<128, 0, 4, 0>(src); run<128, 0, 4, 15>(src); run<128, 0, 5, 15>(src);
    run<256, 0, 0, 15>(src); run<256, 0, 5, 15>(src);
    return 0;
}
