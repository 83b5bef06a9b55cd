// float64 pairwise kernels for the reference's small/medium call sites: materialising
// cdist / pdist, radius-neighbour CSR lists, gathered pair distances and the median of a
// never-materialised distance block (radix select).  All use the reference's arithmetic
// (scipy direct differences in float64, sequential over the dimensions, no FMA) so values
// are bit-identical to scipy's; they are bounded by the float64 pipe / HBM writes, not by
// the FP32 scan.
#include "common.cuh"
#include <math_constants.h>

#include <algorithm>

namespace {

constexpr int PT = 32;  // pair tile edge

__device__ __forceinline__ double acc_step(double s, double a, double b, int metric) {
    const double t = __dsub_rn(a, b);
    return metric == GMD_L2 ? __dadd_rn(s, __dmul_rn(t, t)) : __dadd_rn(s, fabs(t));
}
__device__ __forceinline__ double finish(double s, int metric) { return metric == GMD_L2 ? sqrt(s) : s; }

// A 32x32 tile of pair distances: thread (tx, ty) of a 32x8 block holds rows ty, ty+8, ty+16,
// ty+24 of A against column tx of B.  Rows are staged transposed, 32 dimensions at a time.
struct PairTile {
    double v[4];
};

__device__ __forceinline__ PairTile pair_tile(const double* __restrict__ A, long long mA, long long i0,
                                              const double* __restrict__ B, long long mB, long long j0, int D,
                                              int metric, double (*As)[PT + 1], double (*Bs)[PT + 1]) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    PairTile r;
    r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0.0;
    for (int d0 = 0; d0 < D; d0 += PT) {
        const int dn = min(PT, D - d0);
        __syncthreads();
        for (int e = ty * PT + tx; e < PT * dn; e += PT * 8) {
            const int row = e / dn, d = e - row * dn;
            As[d][row] = (i0 + row < mA) ? A[(size_t)(i0 + row) * D + d0 + d] : 0.0;
            Bs[d][row] = (j0 + row < mB) ? B[(size_t)(j0 + row) * D + d0 + d] : 0.0;
        }
        __syncthreads();
        for (int d = 0; d < dn; ++d) {
            const double b = Bs[d][tx];
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) r.v[r4] = acc_step(r.v[r4], As[d][ty + 8 * r4], b, metric);
        }
    }
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) r.v[r4] = finish(r.v[r4], metric);
    return r;
}

__global__ void __launch_bounds__(256) cdist_kernel(const double* __restrict__ A, long long mA,
                                                    const double* __restrict__ B, long long mB, int D, int metric,
                                                    double* __restrict__ out) {
    __shared__ double As[PT][PT + 1], Bs[PT][PT + 1];
    const long long i0 = (long long)blockIdx.y * PT, j0 = (long long)blockIdx.x * PT;
    const PairTile t = pair_tile(A, mA, i0, B, mB, j0, D, metric, As, Bs);
    const long long j = j0 + threadIdx.x;
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) {
        const long long i = i0 + threadIdx.y + 8 * r4;
        if (i < mA && j < mB) out[(size_t)i * mB + j] = t.v[r4];
    }
}

// condensed pdist: entry (i, j), i < j, at  i*n - i(i+1)/2 + (j - i - 1)  (scipy order,
// the inverse of small2indices, mstore.py:1880-1888)
__global__ void __launch_bounds__(256) pdist_kernel(const double* __restrict__ X, long long n, int D, int metric,
                                                    double* __restrict__ out) {
    __shared__ double As[PT][PT + 1], Bs[PT][PT + 1];
    if (blockIdx.x < blockIdx.y) return;  // tile entirely below the diagonal
    const long long i0 = (long long)blockIdx.y * PT, j0 = (long long)blockIdx.x * PT;
    const PairTile t = pair_tile(X, n, i0, X, n, j0, D, metric, As, Bs);
    const long long j = j0 + threadIdx.x;
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) {
        const long long i = i0 + threadIdx.y + 8 * r4;
        if (i < j && j < n) out[(size_t)(i * n - i * (i + 1) / 2 + (j - i - 1))] = t.v[r4];
    }
}

// ---- radius neighbours: rows of where(D[i, :] <= r)[0], ascending (cluster.py:534, :571)
// One warp walks 4 query rows over all candidate tiles in ascending order, so the CSR fill is
// ordered without a sort.  FILL = false counts, FILL = true writes indices (+ distances).
template <bool FILL>
__global__ void __launch_bounds__(256) radius_kernel(const double* __restrict__ X, long long n, int D, int metric,
                                                     double radius, long long* __restrict__ counts,
                                                     const long long* __restrict__ indptr,
                                                     int32_t* __restrict__ indices, double* __restrict__ dists) {
    extern __shared__ double sm[];  // Bs[D][33] then As[32][D]
    double* Bs = sm;
    double* As = sm + (size_t)D * (PT + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long i0 = (long long)blockIdx.x * PT;
    for (int e = threadIdx.x; e < PT * D; e += 256) {
        const long long row = i0 + e / D;
        As[e] = row < n ? X[(size_t)i0 * D + e] : 0.0;
    }
    long long off[4] = {0, 0, 0, 0};
    if (FILL) {
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
            const long long i = i0 + warp * 4 + r4;
            off[r4] = i < n ? indptr[i] : 0;
        }
    }
    for (long long j0 = 0; j0 < n; j0 += PT) {
        __syncthreads();
        for (int e = threadIdx.x; e < PT * D; e += 256) {
            const int row = e / D, d = e - row * D;
            Bs[d * (PT + 1) + row] = (j0 + row < n) ? X[(size_t)j0 * D + e] : 0.0;
        }
        __syncthreads();
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        for (int d = 0; d < D; ++d) {
            const double b = Bs[d * (PT + 1) + lane];
#pragma unroll
            for (int r4 = 0; r4 < 4; ++r4) s[r4] = acc_step(s[r4], As[(warp * 4 + r4) * D + d], b, metric);
        }
        const long long j = j0 + lane;
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
            const long long i = i0 + warp * 4 + r4;
            const double dist = finish(s[r4], metric);
            const bool hit = (i < n) && (j < n) && (dist <= radius);
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, hit);
            if (FILL && hit) {
                const long long pos = off[r4] + __popc(bal & ((1u << lane) - 1u));
                indices[pos] = (int32_t)j;
                if (dists) dists[pos] = dist;
            }
            off[r4] += __popc(bal);
        }
    }
    if (!FILL && lane == 0) {
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
            const long long i = i0 + warp * 4 + r4;
            if (i < n) counts[i] = off[r4];
        }
    }
}

// distances of the CSR pairs (i, indices[p]) in another profile space (cluster.py:547, :583)
__global__ void __launch_bounds__(256) gather_dist_kernel(const double* __restrict__ X, long long n, int D, int metric,
                                                          const long long* __restrict__ indptr,
                                                          const int32_t* __restrict__ indices,
                                                          double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const double* a = X + (size_t)i * D;
    for (long long p = indptr[i] + lane; p < indptr[i + 1]; p += 32) {
        const double* b = X + (size_t)indices[p] * D;
        double s = 0.0;
        for (int d = 0; d < D; ++d) s = acc_step(s, a[d], b[d], metric);
        out[p] = finish(s, metric);
    }
}

// ---- median of a distance block by radix select on the float64 bit pattern (>= 0, so the
// bits order like the values).  state: [0] prefix bits, [1] rank wanted inside the prefix,
// [2] count(< v_hi), [3] bits of max(value < v_hi); hist: 2048 bins.
constexpr int RBITS = 11;
constexpr int RBINS = 1 << RBITS;

__global__ void __launch_bounds__(256) select_hist_kernel(const double* __restrict__ A, long long mA,
                                                          const double* __restrict__ B, long long mB, int D,
                                                          int metric, int self, int shift,
                                                          const unsigned long long* __restrict__ state,
                                                          unsigned int* __restrict__ hist) {
    __shared__ double As[PT][PT + 1], Bs[PT][PT + 1];
    __shared__ unsigned int h[RBINS];
    if (self && blockIdx.x < blockIdx.y) return;
    const int tid = threadIdx.y * PT + threadIdx.x;
    for (int e = tid; e < RBINS; e += 256) h[e] = 0;
    const long long i0 = (long long)blockIdx.y * PT, j0 = (long long)blockIdx.x * PT;
    const PairTile t = pair_tile(A, mA, i0, B, mB, j0, D, metric, As, Bs);  // syncs inside
    const unsigned long long prefix = state[0];
    const long long j = j0 + threadIdx.x;
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) {
        const long long i = i0 + threadIdx.y + 8 * r4;
        const bool valid = (i < mA) && (j < mB) && (!self || i < j);
        if (valid) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(t.v[r4]);
            const int top = shift + RBITS;  // bits at and above `top` are already decided
            if (top >= 64 || (bits >> top) == (prefix >> top)) atomicAdd(&h[(bits >> shift) & (RBINS - 1)], 1u);
        }
    }
    __syncthreads();
    for (int e = tid; e < RBINS; e += 256)
        if (h[e]) atomicAdd(&hist[e], h[e]);
}

__global__ void select_pick_kernel(unsigned long long* __restrict__ state, unsigned int* __restrict__ hist,
                                   int shift) {
    // single thread: 2048 bins
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long want = state[1], run = 0;
    int bin = RBINS - 1;
    for (int b = 0; b < RBINS; ++b) {
        const unsigned long long c = hist[b];
        if (want < run + c) {
            bin = b;
            break;
        }
        run += c;
    }
    state[0] |= ((unsigned long long)bin) << shift;
    state[1] = want - run;
    for (int b = 0; b < RBINS; ++b) hist[b] = 0;
}

__global__ void __launch_bounds__(256) select_below_kernel(const double* __restrict__ A, long long mA,
                                                           const double* __restrict__ B, long long mB, int D,
                                                           int metric, int self,
                                                           unsigned long long* __restrict__ state) {
    __shared__ double As[PT][PT + 1], Bs[PT][PT + 1];
    if (self && blockIdx.x < blockIdx.y) return;
    const long long i0 = (long long)blockIdx.y * PT, j0 = (long long)blockIdx.x * PT;
    const PairTile t = pair_tile(A, mA, i0, B, mB, j0, D, metric, As, Bs);
    const unsigned long long vhi = state[0];
    const long long j = j0 + threadIdx.x;
    unsigned long long cnt = 0, mx = 0;
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) {
        const long long i = i0 + threadIdx.y + 8 * r4;
        const bool valid = (i < mA) && (j < mB) && (!self || i < j);
        const unsigned long long bits = (unsigned long long)__double_as_longlong(t.v[r4]);
        if (valid && bits < vhi) {
            cnt += 1;
            mx = max(mx, bits);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    }
    if (threadIdx.x == 0 && cnt) {
        atomicAdd(&state[2], cnt);
        atomicMax(&state[3], mx);
    }
}

__global__ void select_finish_kernel(const unsigned long long* __restrict__ state, unsigned long long total,
                                     double* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double vhi = __longlong_as_double((long long)state[0]);
    if (total & 1ull) {
        out[0] = vhi;
    } else {
        // np.median: mean of ranks total/2 - 1 and total/2
        const double vlo = (state[2] == total / 2) ? __longlong_as_double((long long)state[3]) : vhi;
        out[0] = (vlo + vhi) / 2.0;
    }
}

}  // namespace

// ---- angle between coverage profiles (refine.py:1545-1546, 1577-1578):
//   arccos(np.dot(cov[r1], cov[r2]) / (normCoverages[r1] * normCoverages[r2]))
// for a small block of pairs.  The denominator is the product of the caller's normCoverages (the
// reference loads them from its database, profileManager.py:205; they are NOT recomputed from the
// rows), rounded once; the dot product is the sequential fused-multiply-add chain that OpenBLAS'
// ddot runs on rows shorter than its SIMD block (the 3..10-sample rows this path sees).  Pairs the
// reference drops with a FloatingPointError (argument outside [-1, 1], 0/0 or x/0 for a zero norm)
// come back as NaN.  normA / normB null: norms of the rows themselves (sqrt of the same chain).
namespace {
__device__ __forceinline__ double fma_dot(const double* __restrict__ a, const double* __restrict__ b, int D) {
    double s = 0.0;
    for (int d = 0; d < D; ++d) s = __fma_rn(a[d], b[d], s);
    return s;
}
__global__ void __launch_bounds__(256) angle_block_kernel(const double* __restrict__ A, long long mA,
                                                          const double* __restrict__ B, long long mB, int D,
                                                          const double* __restrict__ normA,
                                                          const double* __restrict__ normB,
                                                          double* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= mA * mB) return;
    const long long i = e / mB, j = e % mB;
    const double* a = A + (size_t)i * D;
    const double* b = B + (size_t)j * D;
    const double na = normA ? normA[i] : sqrt(fma_dot(a, a, D));
    const double nb = normB ? normB[j] : sqrt(fma_dot(b, b, D));
    const double den = __dmul_rn(na, nb);
    const double arg = __ddiv_rn(fma_dot(a, b, D), den);
    // den == 0: divide-by-zero / invalid in the reference; |arg| > 1: invalid in arccos; NaN compares false
    out[e] = (den != 0.0 && arg >= -1.0 && arg <= 1.0) ? acos(arg) : CUDART_NAN;
}
}  // namespace

extern "C" int gmd_angle_block_f64(const double* A, int64_t mA, const double* B, int64_t mB, int D,
                                   const double* normA, const double* normB, double* out, void* stream) {
    GMD_REQUIRE(A && B && out && mA > 0 && mB > 0 && D > 0, GMD_ERR_BAD_ARG, "angle_block: bad arguments");
    const long long n = (long long)mA * mB;
    angle_block_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A, mA, B, mB, D, normA, normB,
                                                                                      out);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

extern "C" int gmd_cdist_f64(const double* A, int64_t mA, const double* B, int64_t mB, int D, int metric,
                             double* out, void* stream) {
    GMD_REQUIRE(A && B && out && mA > 0 && mB > 0 && D >= 1, GMD_ERR_BAD_ARG, "cdist: bad arguments");
    GMD_REQUIRE(metric == GMD_L2 || metric == GMD_L1, GMD_ERR_BAD_ARG, "cdist: bad metric %d", metric);
    dim3 grid((unsigned)((mB + PT - 1) / PT), (unsigned)((mA + PT - 1) / PT));
    GMD_REQUIRE(grid.y <= 65535, GMD_ERR_UNSUPPORTED, "cdist: more than %d rows in XA", 65535 * PT);
    cdist_kernel<<<grid, dim3(PT, 8), 0, (cudaStream_t)stream>>>(A, mA, B, mB, D, metric, out);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

extern "C" int gmd_pdist_f64(const double* X, int64_t n, int D, int metric, double* out, void* stream) {
    GMD_REQUIRE(X && out && n > 1 && D >= 1, GMD_ERR_BAD_ARG, "pdist: bad arguments");
    GMD_REQUIRE(metric == GMD_L2 || metric == GMD_L1, GMD_ERR_BAD_ARG, "pdist: bad metric %d", metric);
    dim3 grid((unsigned)((n + PT - 1) / PT), (unsigned)((n + PT - 1) / PT));
    GMD_REQUIRE(grid.y <= 65535, GMD_ERR_UNSUPPORTED, "pdist: more than %d rows", 65535 * PT);
    pdist_kernel<<<grid, dim3(PT, 8), 0, (cudaStream_t)stream>>>(X, n, D, metric, out);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

static int radius_launch(bool fill, const double* X, int64_t n, int D, int metric, double radius, long long* counts,
                         const long long* indptr, int32_t* indices, double* dists, void* stream) {
    const size_t smem = ((size_t)D * (PT + 1) + (size_t)PT * D) * sizeof(double);
    const unsigned grid = (unsigned)((n + PT - 1) / PT);
    if (fill) {
        GMD_CUDA_CHECK(cudaFuncSetAttribute(radius_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        radius_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(X, n, D, metric, radius, counts, indptr, indices,
                                                                       dists);
    } else {
        GMD_CUDA_CHECK(cudaFuncSetAttribute(radius_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        radius_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(X, n, D, metric, radius, counts, indptr,
                                                                        indices, dists);
    }
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

extern "C" int gmd_radius_count_f64(const double* X, int64_t n, int D, int metric, double radius, int64_t* counts,
                                    void* stream) {
    GMD_REQUIRE(X && counts && n > 0 && D >= 1 && D <= GMD_MAX_D, GMD_ERR_BAD_ARG, "radius_count: bad arguments");
    GMD_REQUIRE(metric == GMD_L2 || metric == GMD_L1, GMD_ERR_BAD_ARG, "radius_count: bad metric %d", metric);
    GMD_REQUIRE(radius == radius, GMD_ERR_NONFINITE, "radius_count: radius is NaN");
    return radius_launch(false, X, n, D, metric, radius, (long long*)counts, nullptr, nullptr, nullptr, stream);
}

extern "C" int gmd_radius_fill_f64(const double* X, int64_t n, int D, int metric, double radius,
                                   const int64_t* indptr, int32_t* indices, double* dists, void* stream) {
    GMD_REQUIRE(X && indptr && indices && n > 0 && D >= 1 && D <= GMD_MAX_D, GMD_ERR_BAD_ARG,
                "radius_fill: bad arguments");
    GMD_REQUIRE(metric == GMD_L2 || metric == GMD_L1, GMD_ERR_BAD_ARG, "radius_fill: bad metric %d", metric);
    return radius_launch(true, X, n, D, metric, radius, nullptr, (const long long*)indptr, indices, dists, stream);
}

extern "C" int gmd_gather_dist_f64(const double* X, int64_t n, int D, int metric, const int64_t* indptr,
                                   const int32_t* indices, double* out, void* stream) {
    GMD_REQUIRE(X && indptr && indices && out && n > 0 && D >= 1, GMD_ERR_BAD_ARG, "gather_dist: bad arguments");
    gather_dist_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(X, n, D, metric,
                                                                                 (const long long*)indptr, indices, out);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

// ---- many small medians in one launch: one block per job, the whole radix select inside the
// block (shared-memory histogram, no host round trips).  The reference takes these medians
// bin by bin in Python loops (getKCut refine.py:1479-1488, findMergeGroups :718, :876) on bins
// of 10..10^4 contigs, where a launch-per-pass select is slower than scipy.
namespace {
__device__ __forceinline__ double pair_dist64(const double* __restrict__ a, const double* __restrict__ b, int D,
                                              int metric) {
    double s = 0.0;
    for (int d = 0; d < D; ++d) s = acc_step(s, a[d], b[d], metric);
    return finish(s, metric);
}

__global__ void __launch_bounds__(256) median_batch_kernel(const double* __restrict__ X, int D, int metric,
                                                           const int32_t* __restrict__ rows,
                                                           const int32_t* __restrict__ jobs,  // a_off, a_len, b_off, b_len (<0: pdist)
                                                           double* __restrict__ out) {
    __shared__ unsigned int h[RBINS];
    __shared__ unsigned long long s_prefix, s_want, s_cnt, s_max;
    const int job = blockIdx.x, tid = threadIdx.x;
    const int a_off = jobs[4 * job], mA = jobs[4 * job + 1];
    const bool self = jobs[4 * job + 3] < 0;
    const int b_off = self ? a_off : jobs[4 * job + 2], mB = self ? mA : jobs[4 * job + 3];
    const unsigned long long total =
        self ? (unsigned long long)mA * (unsigned long long)(mA > 0 ? mA - 1 : 0) / 2 : (unsigned long long)mA * mB;
    if (total == 0) {
        if (tid == 0) out[job] = CUDART_NAN;  // np.median of nothing / the reference's "return None"
        return;
    }
    const long long span = (long long)mA * mB;  // pdist walks the square and skips j <= i
    if (tid == 0) {
        s_prefix = 0ull;
        s_want = total / 2;
    }
    for (int shift = 55; shift >= 0; shift -= RBITS) {
        for (int e = tid; e < RBINS; e += 256) h[e] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const int top = shift + RBITS;
        for (long long p = tid; p < span; p += 256) {
            const int i = (int)(p / mB), j = (int)(p - (long long)i * mB);
            if (self && j <= i) continue;
            const double d = pair_dist64(X + (size_t)rows[a_off + i] * D, X + (size_t)rows[b_off + j] * D, D, metric);
            const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
            if (top >= 64 || (bits >> top) == (prefix >> top)) atomicAdd(&h[(bits >> shift) & (RBINS - 1)], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long want = s_want, run = 0;
            int bin = RBINS - 1;
            for (int b = 0; b < RBINS; ++b) {
                const unsigned long long c = h[b];
                if (want < run + c) {
                    bin = b;
                    break;
                }
                run += c;
            }
            s_prefix = prefix | ((unsigned long long)bin << shift);
            s_want = want - run;
            s_cnt = 0ull;
            s_max = 0ull;
        }
        __syncthreads();
    }
    const unsigned long long vhi = s_prefix;
    if (total & 1ull) {
        if (tid == 0) out[job] = __longlong_as_double((long long)vhi);
        return;
    }
    // even count: np.median averages ranks total/2 - 1 and total/2
    unsigned long long cnt = 0, mx = 0;
    for (long long p = tid; p < span; p += 256) {
        const int i = (int)(p / mB), j = (int)(p - (long long)i * mB);
        if (self && j <= i) continue;
        const double d = pair_dist64(X + (size_t)rows[a_off + i] * D, X + (size_t)rows[b_off + j] * D, D, metric);
        const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
        if (bits < vhi) {
            cnt += 1;
            mx = max(mx, bits);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    }
    if ((tid & 31) == 0 && cnt) {
        atomicAdd(&s_cnt, cnt);
        atomicMax(&s_max, mx);
    }
    __syncthreads();
    if (tid == 0) {
        const double hi = __longlong_as_double((long long)vhi);
        const double lo = (s_cnt == total / 2) ? __longlong_as_double((long long)s_max) : hi;
        out[job] = (lo + hi) / 2.0;
    }
}
}  // namespace

extern "C" int gmd_median_batch_f64(const double* X, int64_t N, int D, int metric, const int32_t* rows,
                                    const int32_t* jobs, int n_jobs, double* out, void* stream) {
    GMD_REQUIRE(X && rows && jobs && out && N > 0 && D >= 1 && n_jobs >= 0, GMD_ERR_BAD_ARG,
                "median_batch: bad arguments");
    GMD_REQUIRE(metric == GMD_L2 || metric == GMD_L1, GMD_ERR_BAD_ARG, "median_batch: bad metric %d", metric);
    if (n_jobs == 0) return GMD_OK;
    median_batch_kernel<<<(unsigned)n_jobs, 256, 0, (cudaStream_t)stream>>>(X, D, metric, rows, jobs, out);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

extern "C" size_t gmd_median_workspace_bytes(void) { return 4 * sizeof(unsigned long long) + RBINS * sizeof(unsigned int); }

extern "C" int gmd_median_dist_f64(const double* A, int64_t mA, const double* B, int64_t mB, int D, int metric,
                                   double* out_median, void* workspace, void* stream) {
    GMD_REQUIRE(A && out_median && workspace && mA > 0 && D >= 1, GMD_ERR_BAD_ARG, "median_dist: bad arguments");
    GMD_REQUIRE(metric == GMD_L2 || metric == GMD_L1, GMD_ERR_BAD_ARG, "median_dist: bad metric %d", metric);
    const int self = (B == nullptr);
    if (self) {
        B = A;
        mB = mA;
    }
    GMD_REQUIRE(mB > 0, GMD_ERR_BAD_ARG, "median_dist: empty XB");
    const unsigned long long total =
        self ? (unsigned long long)mA * (unsigned long long)(mA - 1) / 2 : (unsigned long long)mA * (unsigned long long)mB;
    GMD_REQUIRE(total > 0, GMD_ERR_BAD_ARG, "median_dist: no pairs (a bin of one contig has no pdist)");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* state = (unsigned long long*)workspace;
    unsigned int* hist = (unsigned int*)(state + 4);
    const unsigned long long init[4] = {0ull, total / 2, 0ull, 0ull};
    GMD_CUDA_CHECK(cudaMemcpyAsync(state, init, sizeof(init), cudaMemcpyHostToDevice, st));
    GMD_CUDA_CHECK(cudaMemsetAsync(hist, 0, RBINS * sizeof(unsigned int), st));
    dim3 grid((unsigned)((mB + PT - 1) / PT), (unsigned)((mA + PT - 1) / PT));
    GMD_REQUIRE(grid.y <= 65535, GMD_ERR_UNSUPPORTED, "median_dist: more than %d rows in XA", 65535 * PT);
    const dim3 block(PT, 8);
    // 64 bits = 9 + 5 x 11, most significant digit first
    for (int shift = 55; shift >= 0; shift -= RBITS) {
        select_hist_kernel<<<grid, block, 0, st>>>(A, mA, B, mB, D, metric, self, shift, state, hist);
        select_pick_kernel<<<1, 32, 0, st>>>(state, hist, shift);
    }
    if (!(total & 1ull)) select_below_kernel<<<grid, block, 0, st>>>(A, mA, B, mB, D, metric, self, state);
    select_finish_kernel<<<1, 32, 0, st>>>(state, total, out_median);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

// ---- small helpers of the host layer (so that no eager PyTorch kernel sits on the pairwise paths)
namespace {
__global__ void __launch_bounds__(256) nonfinite_kernel(const double* __restrict__ X, long long n, int32_t* __restrict__ flag) {
    bool bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        bad |= !isfinite(X[i]);
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0) atomicExch(flag, 1);
}

// one block: out[0] = 0, out[i + 1] = in[0] + ... + in[i]
__global__ void __launch_bounds__(1024) scan_i64_kernel(const long long* __restrict__ in, long long n,
                                                        long long* __restrict__ out) {
    __shared__ long long part[1024];
    const int tid = threadIdx.x;
    const long long per = (n + 1023) / 1024;
    const long long a = min(n, tid * per), b = min(n, a + per);
    long long sum = 0;
    for (long long i = a; i < b; ++i) sum += in[i];
    part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const long long v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    long long run = part[tid] - sum;
    if (tid == 0) out[0] = 0;
    for (long long i = a; i < b; ++i) {
        run += in[i];
        out[i + 1] = run;
    }
}
}  // namespace

// ---- best-matching unit of a FEW vectors, straight in float64: argmin(cdist(nodes, [v])) as the reference
// calls it per contig (torusMesh.py:127, som.py:351).  One block per vector, scipy's arithmetic, first minimum.
// One launch and no staging: what a single bestMatch call costs is launch latency, not arithmetic.
namespace {
__global__ void __launch_bounds__(256) bmu_small_kernel(const double* __restrict__ nodes, long long M,
                                                       const double* __restrict__ X, int D, int32_t* __restrict__ out_idx,
                                                       double* __restrict__ out_dist, int32_t* __restrict__ flag) {
    __shared__ double sd[8];
    __shared__ long long si[8];
    const double* v = X + (size_t)blockIdx.x * D;
    double best = CUDART_INF;
    long long bi = 0x7fffffffffffffffll;
    bool bad = false;
    for (long long j = threadIdx.x; j < M; j += blockDim.x) {
        const double* c = nodes + (size_t)j * D;
        double s = 0.0;
        for (int d = 0; d < D; ++d) {
            const double t = __dsub_rn(c[d], v[d]);
            s = __dadd_rn(s, __dmul_rn(t, t));
        }
        const double dist = sqrt(s);
        bad |= !(dist == dist) || dist == CUDART_INF;
        if (dist < best || (dist == best && j < bi)) best = dist, bi = j;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xFFFFFFFFu, best, o);
        const long long oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
        if (od < best || (od == best && oi < bi)) best = od, bi = oi;
    }
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0 && flag) atomicExch(flag, 1);
    if ((threadIdx.x & 31) == 0) sd[threadIdx.x >> 5] = best, si[threadIdx.x >> 5] = bi;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (sd[w] < best || (sd[w] == best && si[w] < bi)) best = sd[w], bi = si[w];
        out_idx[blockIdx.x] = (int32_t)bi;
        if (out_dist) out_dist[blockIdx.x] = best;
    }
}
}  // namespace

extern "C" int gmd_bmu_small_f64(const double* nodes, int64_t M, const double* X, int64_t N, int D, int32_t* out_idx,
                                 double* out_dist, int32_t* flag, void* stream) {
    GMD_REQUIRE(nodes && X && out_idx && M > 0 && M < 0x7fffffffll && N > 0 && N <= 65535 && D >= 1, GMD_ERR_BAD_ARG,
                "bmu_small: bad arguments");
    if (flag) GMD_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int32_t), (cudaStream_t)stream));
    bmu_small_kernel<<<(unsigned)N, 256, 0, (cudaStream_t)stream>>>(nodes, M, X, D, out_idx, out_dist, flag);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

// ---- k nearest candidates of a FEW query rows, straight in float64: KDTree(data).query(x, k) as
// findMergeGroups calls it once per bin (refine.py:653-665).  One block per query: all distances into shared
// memory (scipy's arithmetic), then k rounds of "smallest (distance, index) above the last one taken".
namespace {
constexpr int KNS_MAX_NC = 16384;   // candidates whose distances fit a block's shared memory
__global__ void __launch_bounds__(256) knn_small_kernel(const double* __restrict__ Q, const double* __restrict__ C,
                                                       int Nc, int D, int k, int metric, int32_t* __restrict__ out_idx,
                                                       double* __restrict__ out_dist, int32_t* __restrict__ flag) {
    extern __shared__ double kd[];                    // [Nc] distances of this query
    __shared__ double rd[8];
    __shared__ int ri[8];
    __shared__ double last_d;
    __shared__ int last_i;
    const double* q = Q + (size_t)blockIdx.x * D;
    bool bad = false;
    for (int j = threadIdx.x; j < Nc; j += blockDim.x) {
        const double* c = C + (size_t)j * D;
        double s = 0.0;
        if (metric == GMD_L2) {
            for (int d = 0; d < D; ++d) {
                const double t = __dsub_rn(q[d], c[d]);
                s = __dadd_rn(s, __dmul_rn(t, t));
            }
            s = sqrt(s);
        } else {
            for (int d = 0; d < D; ++d) s = __dadd_rn(s, fabs(__dsub_rn(q[d], c[d])));
        }
        bad |= !(s == s) || s == CUDART_INF;
        kd[j] = s;
    }
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0 && flag) atomicExch(flag, 1);
    if (threadIdx.x == 0) last_d = -1.0, last_i = -1;
    __syncthreads();
    for (int r = 0; r < k; ++r) {
        const double pd = last_d;
        const int pi = last_i;
        double best = CUDART_INF;
        int bi = 0x7fffffff;
        for (int j = threadIdx.x; j < Nc; j += blockDim.x) {
            const double v = kd[j];
            const bool after = v > pd || (v == pd && j > pi);       // not taken yet
            if (after && (v < best || (v == best && j < bi))) best = v, bi = j;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xFFFFFFFFu, best, o);
            const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
            if (od < best || (od == best && oi < bi)) best = od, bi = oi;
        }
        if ((threadIdx.x & 31) == 0) rd[threadIdx.x >> 5] = best, ri[threadIdx.x >> 5] = bi;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w)
                if (rd[w] < best || (rd[w] == best && ri[w] < bi)) best = rd[w], bi = ri[w];
            out_idx[(size_t)blockIdx.x * k + r] = bi;
            out_dist[(size_t)blockIdx.x * k + r] = best;
            last_d = best, last_i = bi;
        }
        __syncthreads();
    }
}
}  // namespace

extern "C" int gmd_knn_small_max_candidates(void) { return KNS_MAX_NC; }

extern "C" int gmd_knn_small_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D, int k, int metric,
                                 int32_t* out_idx, double* out_dist, int32_t* flag, void* stream) {
    GMD_REQUIRE(Q && C && out_idx && out_dist && Nq > 0 && Nq <= 65535 && Nc > 0 && Nc <= KNS_MAX_NC && D >= 1 && k >= 1 &&
                    k <= Nc,
                GMD_ERR_BAD_ARG, "knn_small: bad arguments (Nq=%lld Nc=%lld k=%d)", (long long)Nq, (long long)Nc, k);
    GMD_REQUIRE(metric == GMD_L2 || metric == GMD_L1, GMD_ERR_BAD_ARG, "knn_small: bad metric %d", metric);
    if (flag) GMD_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int32_t), (cudaStream_t)stream));
    const size_t smem = (size_t)Nc * sizeof(double);
    GMD_CUDA_CHECK(cudaFuncSetAttribute(knn_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_small_kernel<<<(unsigned)Nq, 256, smem, (cudaStream_t)stream>>>(Q, C, (int)Nc, D, k, metric, out_idx, out_dist, flag);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

extern "C" int gmd_zero_async(void* p, size_t bytes, void* stream) {
    GMD_REQUIRE(p || bytes == 0, GMD_ERR_BAD_ARG, "zero_async: null pointer");
    if (bytes) GMD_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, (cudaStream_t)stream));
    return GMD_OK;
}

extern "C" int gmd_nonfinite_f64(const double* X, int64_t n, int32_t* flag, void* stream) {
    GMD_REQUIRE(X && flag && n >= 0, GMD_ERR_BAD_ARG, "nonfinite: bad arguments");
    if (n == 0) return GMD_OK;
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8);
    nonfinite_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, n, flag);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

extern "C" int gmd_indptr_from_counts_i64(const int64_t* counts, int64_t n, int64_t* indptr, void* stream) {
    GMD_REQUIRE(counts && indptr && n > 0, GMD_ERR_BAD_ARG, "indptr_from_counts: bad arguments");
    scan_i64_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((const long long*)counts, n, (long long*)indptr);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}
