// Locality-preserving staging order for the tensor-core scan (scan_tc.cu, SWEEP): profiles are grouped by
// a small integer key (the nearest of a few hundred sampled profiles) with a counting sort, so that the
// rows of a tile -- and of neighbouring tiles -- are close in profile space.  The reference has no
// counterpart: it is an execution detail of the scan; results are reported under the original row numbers.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) key_hist_kernel(const int32_t* __restrict__ key, long long N, int nkeys,
                                                       int32_t* __restrict__ count, int32_t* __restrict__ bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int k = key[i];
    if (k < 0 || k >= nkeys) {                         // (a row without a valid key, e.g. a non-finite profile)
        *bad = 1;
        k = min(max(k, 0), nkeys - 1);                 // still placed: perm must stay a permutation
    }
    atomicAdd(&count[k], 1);
}

// one block: start[k] = sum of count[< k], start[nkeys] = total; cursor = copy of start
__global__ void __launch_bounds__(1024) key_scan_kernel(const int32_t* __restrict__ count, int nkeys,
                                                        int32_t* __restrict__ start, int32_t* __restrict__ cursor) {
    __shared__ int32_t part[1024];
    const int tid = threadIdx.x;
    const int per = (nkeys + 1023) / 1024;
    const int a = tid * per, b = min(nkeys, a + per);
    int32_t sum = 0;
    for (int k = a; k < b; ++k) sum += count[k];
    part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int32_t v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int32_t run = part[tid] - sum;                     // exclusive prefix of this thread's keys
    for (int k = a; k < b; ++k) {
        start[k] = run;
        cursor[k] = run;
        run += count[k];
    }
    if (tid == 1023) start[nkeys] = part[1023];
}

__global__ void __launch_bounds__(256) key_scatter_kernel(const int32_t* __restrict__ key, long long N, int nkeys,
                                                          int32_t* __restrict__ cursor, uint32_t* __restrict__ perm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int k = min(max(key[i], 0), nkeys - 1);
    perm[atomicAdd(&cursor[k], 1)] = (uint32_t)i;      // order inside a key group is arbitrary (and irrelevant)
}

}  // namespace

extern "C" size_t gmd_order_workspace_bytes(int nkeys) {
    return nkeys > 0 ? (size_t)(3 * (size_t)nkeys + 2) * sizeof(int32_t) : 0;
}

extern "C" int gmd_order_by_key(const int32_t* key, int64_t N, int nkeys, uint32_t* perm, void* workspace,
                                size_t workspace_bytes, void* stream) {
    GMD_REQUIRE(key && perm && workspace, GMD_ERR_BAD_ARG, "order_by_key: null pointer");
    GMD_REQUIRE(N > 0 && N < 0xFFFFFFFFll && nkeys >= 1 && nkeys <= (1 << 20), GMD_ERR_BAD_ARG,
                "order_by_key: bad sizes (N=%lld nkeys=%d)", (long long)N, nkeys);
    GMD_REQUIRE(workspace_bytes >= gmd_order_workspace_bytes(nkeys), GMD_ERR_WORKSPACE,
                "order_by_key: workspace %zu B < required %zu B", workspace_bytes, gmd_order_workspace_bytes(nkeys));
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* count = (int32_t*)workspace;              // [nkeys]
    int32_t* start = count + nkeys;                    // [nkeys + 1]
    int32_t* cursor = start + nkeys + 1;               // [nkeys]
    int32_t* bad = cursor + nkeys;                     // [1]
    GMD_CUDA_CHECK(cudaMemsetAsync(workspace, 0, gmd_order_workspace_bytes(nkeys), st));
    const unsigned grid = (unsigned)((N + 255) / 256);
    key_hist_kernel<<<grid, 256, 0, st>>>(key, N, nkeys, count, bad);
    GMD_CUDA_CHECK(cudaGetLastError());
    key_scan_kernel<<<1, 1024, 0, st>>>(count, nkeys, start, cursor);
    GMD_CUDA_CHECK(cudaGetLastError());
    key_scatter_kernel<<<grid, 256, 0, st>>>(key, N, nkeys, cursor, perm);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}
