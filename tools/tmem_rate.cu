// TMEM read-out probe: 4 warps (one per lane quarter) loop tcgen05.ld.32x32b.x32 over 128 columns
// (x4 per "tile") and report cycles per 128 x 128 accumulator tile.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BATCH>
__global__ void __launch_bounds__(128, 1) rd(int iters, long long* cycles, float* sink) {
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base + ((uint32_t)(warp * 32) << 16);
    float acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t base = tb + (uint32_t)((it & 3) * 128);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(base + c * 32));
            if (BATCH == 1 || c == 3) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
        }
    }
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}
template <int BATCH>
void run(const char* name) {
    long long* dc; float* ds; cudaMalloc(&dc, 8); cudaMalloc(&ds, 16);
    const int iters = 20000;
    rd<BATCH><<<148, 128>>>(10, dc, ds); cudaDeviceSynchronize();
    rd<BATCH><<<148, 128>>>(iters, dc, ds); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("%s: %.0f cycles per 128x128 tile read-out (4 warps x 4 x LDTM.x32), %s\n", name, (double)c / iters, cudaGetErrorString(cudaGetLastError()));
}
int main() { run<1>("wait after every load"); run<4>("4 loads in flight, one wait"); return 0; }
