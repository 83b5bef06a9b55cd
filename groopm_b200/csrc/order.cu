// Locality-preserving staging order for the tensor-core scans (scan_tc.cu SWEEP, scan_tc_prune.cu):
// profiles are grouped by a small integer key with a counting sort, so that the rows of a tile -- and of
// neighbouring tiles -- are close in profile space.  The reference has no counterpart: it is an execution
// detail of the scan; results are reported under the original row numbers.
//
// Two keyings:
//  * gmd_order_by_key: any caller-supplied key (round 1: the nearest of ~N/1024 sampled profiles).
//  * the GROUPED order of the pruned scan: "leaders" are picked from a strided sample of the rows by
//    leader clustering (a sample becomes a leader unless an earlier leader lies within R = 2 x the
//    median nearest-sample distance), every row is keyed by (nearest leader, distance shell), and the
//    counting sort pads every leader's group to whole 128-row tiles, so a tile holds rows of ONE
//    group and the group's rows of similar distance to the leader sit together.  Balls around the
//    tiles and groups then give the scan rigorous lower bounds on whole tiles of candidates.
#include <algorithm>

#include "common.cuh"

// distance shells of a group: count, and width in units of the leader radius R
// (in high dimensions d(q, c)^2 ~ |q - L|^2 + |c - L|^2 inside a group, so the rows nearest their leader are near
// everybody: thin shells put them first in the group's window and the thresholds tighten sooner -- 8 shells of
// R / 2: scan 5.03 ms, 64 of R / 16: 4.80 ms at 1M x 30)
#ifndef GMD_SHELLS
#define GMD_SHELLS 64
#define GMD_SHELL_WIDTH 0.0625
#endif

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


// ---------------------------------------------------------------- leader clustering of a row sample
constexpr int LEAD_THREADS = 1024;
constexpr int LEAD_PROBES = 256;       // samples whose nearest-sample distance fixes the radius

// sample i = row i * stride of X, centred, FP32 -> S[ns][D]
__global__ void __launch_bounds__(256) gather_samples_kernel(const double* __restrict__ X, long long stride, int ns,
                                                            int D, const double* __restrict__ center,
                                                            float* __restrict__ S) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)ns * D) return;
    const long long i = e / D;
    const int d = (int)(e - i * D);
    S[e] = (float)(X[(size_t)(i * stride) * D + d] - (center ? center[d] : 0.0));
}

// distance from probe p (sample p * (ns / nprobe)) to its nearest DISTINCT sample; one block per probe
__global__ void __launch_bounds__(1024) sample_nn_kernel(const float* __restrict__ S, int ns, int D, int nprobe,
                                                        float* __restrict__ nn) {
    __shared__ float part[32];
    __shared__ float as[GMD_MAX_D];
    const int w = blockIdx.x;
    const int p = w * (ns / nprobe);
    for (int d = threadIdx.x; d < D; d += blockDim.x) as[d] = S[(size_t)p * D + d];
    __syncthreads();
    float best = __int_as_float(0x7f800000);
    const bool pairs = (D & 1) == 0;                   // rows of an even D are 8-byte aligned
    for (int j = threadIdx.x; j < ns; j += blockDim.x) {
        const float* b = S + (size_t)j * D;
        float d2 = 0.0f;
        if (pairs) {
#pragma unroll 5
            for (int d = 0; d < D; d += 2) {
                const float2 v = *reinterpret_cast<const float2*>(b + d);
                const float t0 = as[d] - v.x, t1 = as[d + 1] - v.y;
                d2 = fmaf(t0, t0, d2);
                d2 = fmaf(t1, t1, d2);
            }
        } else {
#pragma unroll 6
            for (int d = 0; d < D; ++d) {
                const float t = as[d] - b[d];
                d2 = fmaf(t, t, d2);
            }
        }
        if (d2 > 0.0f) best = fminf(best, d2);
    }
    for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int x = 1; x < (int)(blockDim.x >> 5); ++x) best = fminf(best, part[x]);
        nn[w] = sqrtf(best);
    }
}

// Leader clustering of the samples: in sample order, a sample becomes a leader unless an earlier LEADER lies
// within R = rfac x median(nn) of it (at most cap leaders).  Greedy and order dependent -- but only the
// decision is: which pairs of samples lie within R of each other is a plain all-pairs question, answered
// by the whole GPU into a bit matrix (adjacency_kernel), and the greedy pass (resolve_kernel, one block)
// then walks the samples with 32-bit ANDs against the mask of the leaders so far.
//   pass A   the first n1 samples among themselves;
//   cover    every other sample against pass A's leaders (cover_kernel, whole GPU);
//   pass C   the samples still uncovered, among themselves (they are far from all of pass A's leaders).
// info[0] = leaders, info[1] = float bits of R, info[2] = pass A's leaders, info[3] = uncovered samples.

// R from the nearest-sample distances of the probes: the median of the finite ones, by rank counting
__global__ void __launch_bounds__(LEAD_PROBES) radius_kernel(const float* __restrict__ nn, int nprobe, float rfac,
                                                             int32_t* __restrict__ info) {
    __shared__ float vals[LEAD_PROBES];
    const int tid = threadIdx.x;
    const bool probe = tid < nprobe;
    const float v = probe ? nn[tid] : __int_as_float(0x7f800000);
    vals[tid] = v;
    if (tid == 0) info[0] = 0, info[1] = 0, info[2] = 0, info[3] = 0;
    const int nfin = __syncthreads_count(probe && v < __int_as_float(0x7f800000));
    if (probe && nfin > 0) {
        int rank = 0;
        for (int j = 0; j < nprobe; ++j) rank += (vals[j] < v || (vals[j] == v && j < tid)) ? 1 : 0;
        if (rank == nfin / 2) info[1] = __float_as_int(rfac * v);
    }
}

// bits[i][c] bit l  <-  |s_i - s_j| <= R for j = 32 c + l, over the n samples list[0..n) (list null: samples
// 0..n-1; n_dev non-null: n is read from the device).  One block per 8 rows i; a warp takes 32 columns at a time,
// lane = column, the 8 rows broadcast from shared memory, one ballot per row.
__global__ void __launch_bounds__(256) adjacency_kernel(const float* __restrict__ S, const int32_t* __restrict__ list,
                                                       const int32_t* __restrict__ n_dev, int n_max, int D,
                                                       const int32_t* __restrict__ info, uint32_t* __restrict__ bits,
                                                       int wpr) {
    extern __shared__ float rows[];                      // [8][D]
    const int n = n_dev ? min(*n_dev, n_max) : n_max;
    const int i0 = blockIdx.x * 8;
    if (i0 >= n) return;
    const int nr = min(8, n - i0);
    for (int e = threadIdx.x; e < nr * D; e += blockDim.x) {
        const int r = e / D, d = e - r * D;
        const int si = list ? list[i0 + r] : i0 + r;
        rows[e] = S[(size_t)si * D + d];
    }
    __syncthreads();
    const float R = __int_as_float(info[1]), R2 = R * R;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // only the columns below the block's last row matter (a sample is tested against EARLIER leaders)
    const int ncol = min(n, i0 + nr);
    for (int c = warp; c * 32 < ncol; c += 8) {
        const int j = c * 32 + lane;
        float d2[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) d2[r] = 0.0f;
        if (j < n) {
            const float* b = S + (size_t)(list ? list[j] : j) * D;
#pragma unroll 4
            for (int d = 0; d < D; ++d) {
                const float v = b[d];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float t = rows[r * D + d] - v;
                    d2[r] = fmaf(t, t, d2[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const unsigned w = __ballot_sync(0xFFFFFFFFu, j < n && d2[r] <= R2);
            if (lane == 0 && r < nr) bits[(size_t)(i0 + r) * wpr + c] = w;
        }
    }
}

// The greedy pass over the n samples of list (null: 0..n-1), one block.  Sample i is a leader iff no bit of row i
// of the bit matrix below column i meets the mask of the leaders so far.  Warp 0 decides a chunk of 32 samples
// at a time from shared memory while the other warps fetch the next chunk into the second buffer (one block
// barrier per chunk, no global-memory round trip on the sequential path):
//   lane r scans row r against the leaders of the EARLIER chunks, all 32 rows side by side;
//   what is left is the chunk's own 32 x 32 block: a 32-step loop on registers (row r's word by shuffle,
//   the chunk's leader bits in a uniform register).
// The leaders' rows are copied by the whole block once all decisions are made.  New leaders are appended to
// L / lead_row from info[0] on.
__global__ void __launch_bounds__(LEAD_THREADS) resolve_kernel(const float* __restrict__ S, const int32_t* __restrict__ list,
                                                              const int32_t* __restrict__ n_dev, int n_max, int D,
                                                              const uint32_t* __restrict__ bits, int wpr, int cap,
                                                              long long stride, float* __restrict__ L,
                                                              int32_t* lead_row, int32_t* __restrict__ info,
                                                              int pass) {
    extern __shared__ uint32_t rsm[];
    const int pitch = wpr | 1;                           // odd row pitch: lane r reading word w of row r is conflict-free
    uint32_t* mask = rsm;                                // [wpr] leaders among the samples decided so far
    uint32_t* chunk = rsm + wpr;                         // [2][32][pitch]
    __shared__ int nlead_s;
    const int n = n_dev ? min(*n_dev, n_max) : n_max;
    const int tid = threadIdx.x, lane = tid & 31;
    const int base = info[0];
    for (int w = tid; w < wpr; w += LEAD_THREADS) mask[w] = 0u;
    if (tid == 0) nlead_s = base;
    auto fetch = [&](int i0, uint32_t* buf, int t0, int nt) {
        const int nr = min(32, n - i0);
        const int nwords = (i0 >> 5) + 1;                // columns up to the chunk's own word
        for (int e = t0; e < nr * nwords; e += nt) {
            const int r = e / nwords, w = e - r * nwords;
            buf[r * pitch + w] = bits[(size_t)(i0 + r) * wpr + w];
        }
    };
    if (n > 0) fetch(0, chunk, tid, LEAD_THREADS);
    __syncthreads();
    int nl = base;                                       // warp 0's running count (the same in all its lanes)
    for (int i0 = 0, c = 0; i0 < n; i0 += 32, ++c) {
        const uint32_t* cur = chunk + (size_t)(c & 1) * 32 * pitch;
        if (tid >= 32) {
            if (i0 + 32 < n) fetch(i0 + 32, chunk + (size_t)((c + 1) & 1) * 32 * pitch, tid - 32, LEAD_THREADS - 32);
        } else if (nl < cap) {
            const int nr = min(32, n - i0), wc = i0 >> 5;
            const uint32_t* row = cur + lane * pitch;
            uint32_t hit = 0u;
#pragma unroll 4
            for (int w = 0; w < wc; ++w) hit |= row[w] & mask[w];
            const uint32_t own = lane < nr ? (row[wc] & ((1u << lane) - 1u)) : 0u;
            const uint32_t cand = __ballot_sync(0xFFFFFFFFu, lane < nr && hit == 0u);
            uint32_t lm = 0u;
            int room = cap - nl;
            for (int r = 0; r < nr; ++r) {
                const uint32_t o = __shfl_sync(0xFFFFFFFFu, own, r);
                if (((cand >> r) & 1u) && (o & lm) == 0u && room > 0) lm |= 1u << r, --room;
            }
            if ((lm >> lane) & 1u)
                lead_row[nl + __popc(lm & ((1u << lane) - 1u))] =
                    (int32_t)((long long)(list ? list[i0 + lane] : i0 + lane) * stride);
            if (lane == 0) mask[wc] = lm;
            nl += __popc(lm);
            if (lane == 0) nlead_s = nl;
        }
        __syncthreads();
    }
    // the new leaders' rows (lead_row holds sample index x stride, written above by this block)
    const int total = nlead_s;
    for (int e = tid; e < (total - base) * D; e += LEAD_THREADS) {
        const int x = e / D, d = e - x * D;
        const long long si = (long long)lead_row[base + x] / stride;
        L[(size_t)(base + x) * D + d] = S[(size_t)si * D + d];
    }
    if (tid == 0) {
        info[0] = total;
        if (pass == 0) info[2] = total;
    }
}

// flags[i] = 1 for every sample i >= n1 that no leader of pass A covers; one warp per sample
__global__ void __launch_bounds__(256) cover_kernel(const float* __restrict__ S, int ns, int n1, int D,
                                                   const float* __restrict__ L, const int32_t* __restrict__ info,
                                                   unsigned char* __restrict__ flags) {
    const int i = n1 + (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (i >= ns) return;
    const int nl = info[0];
    const float R = __int_as_float(info[1]), R2 = R * R;
    const float* a = S + (size_t)i * D;
    bool hit = false;
    for (int j0 = 0; j0 < nl && !hit; j0 += 32) {
        const int j = j0 + lane;
        float d2 = __int_as_float(0x7f800000);
        if (j < nl) {
            const float* l = L + (size_t)j * D;
            d2 = 0.0f;
#pragma unroll 6
            for (int d = 0; d < D; ++d) {
                const float t = a[d] - l[d];
                d2 = fmaf(t, t, d2);
            }
        }
        hit = __any_sync(0xFFFFFFFFu, d2 <= R2);
    }
    if (lane == 0) flags[i] = hit ? 0 : 1;
}

// ordered list of the uncovered samples i >= n1 (one block); info[3] = their number
__global__ void __launch_bounds__(LEAD_THREADS) uncovered_kernel(const unsigned char* __restrict__ flags, int ns, int n1,
                                                                int32_t* __restrict__ list, int32_t* __restrict__ info) {
    __shared__ int part[LEAD_THREADS];
    const int tid = threadIdx.x;
    const int per = (ns - n1 + LEAD_THREADS - 1) / LEAD_THREADS;
    const int a = min(ns, n1 + tid * per), b = min(ns, a + per);
    int c = 0;
    for (int i = a; i < b; ++i) c += flags[i] ? 1 : 0;
    part[tid] = c;
    __syncthreads();
    for (int o = 1; o < LEAD_THREADS; o <<= 1) {
        const int v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int pos = part[tid] - c;
    for (int i = a; i < b; ++i)
        if (flags[i]) list[pos++] = i;
    if (tid == LEAD_THREADS - 1) info[3] = part[tid];
}

// leader rows of X, float64 row-major (the candidate set of the nearest-leader search)
__global__ void __launch_bounds__(256) gather_rows_kernel(const double* __restrict__ X, const int32_t* __restrict__ rows,
                                                         int n, int D, double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * D) return;
    const int i = e / D, d = e - i * D;
    out[e] = X[(size_t)rows[i] * D + d];
}

// key = leader * NB + distance shell (width R / 2; the last shell is open); grad[leader] = largest
// member distance (non-negative floats order like their bit patterns; rounded up)
__global__ void __launch_bounds__(256) group_key_kernel(const int32_t* __restrict__ lead, const double* __restrict__ dist,
                                                       long long N, int nlead, int NB, const int32_t* __restrict__ info,
                                                       int32_t* __restrict__ key, float* __restrict__ grad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int g = min(max(lead[i], 0), nlead - 1);
    const double d = dist[i];
    const float R = __int_as_float(info[1]);
    int b = NB - 1;
    if (d == d && R > 0.0f) {
        // (the shell only orders rows inside a group: single precision is plenty)
        const float s = (float)d / (R * (float)GMD_SHELL_WIDTH);
        b = s < (float)(NB - 1) ? (int)s : NB - 1;
    } else if (d == 0.0) {
        b = 0;
    }
    key[i] = g * NB + b;
    if (grad && d == d) {
        // 500 addresses for a million rows: look first (a stale smaller value only costs the atomic it would have cost)
        const int bits = __float_as_int(__double2float_ru(d));
        if (bits > __ldcg(reinterpret_cast<const int*>(grad) + g)) atomicMax(reinterpret_cast<int*>(grad) + g, bits);
    }
}

// One block.  Padded starts of a (group, shell) counting sort: group g's rows begin on a tile boundary.
// start / cursor [nkeys], gstart[ngroups + 1] in TILES; gstart[ngroups] * tile = positions used.
// A warp per group for the passes over the shells (coalesced reads of its NB <= 64 counts, sums and offsets by
// shuffles), the block for the running sum over the groups.
constexpr int SCAN_GROUPS_MAX = 4096;                  // = the cap of gmd_group_params
__global__ void __launch_bounds__(1024) key_scan_padded_kernel(const int32_t* __restrict__ count, int ngroups, int NB,
                                                               int tile, int pad, int32_t* __restrict__ start,
                                                               int32_t* __restrict__ cursor,
                                                               int32_t* __restrict__ gstart) {
    __shared__ int32_t gt[SCAN_GROUPS_MAX];            // tiles of every group, then their exclusive running sum
    __shared__ int32_t part[1024];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int g = warp; g < ngroups; g += 32) {
        int32_t c = (lane < NB ? count[g * NB + lane] : 0) + (lane + 32 < NB ? count[g * NB + lane + 32] : 0);
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
        if (lane == 0) gt[g] = ((c + tile - 1) / tile + pad - 1) / pad * pad;   // whole tiles, a multiple of pad of them
    }
    __syncthreads();
    const int per = (ngroups + 1023) / 1024;
    const int a = min(ngroups, tid * per), b = min(ngroups, a + per);
    int32_t sum = 0;
    for (int g = a; g < b; ++g) sum += gt[g];
    part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int32_t v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int32_t run = part[tid] - sum;                     // first tile of this thread's first group
    for (int g = a; g < b; ++g) {
        const int32_t t = gt[g];
        gt[g] = run;
        gstart[g] = run;
        run += t;
    }
    if (tid == 1023) gstart[ngroups] = part[1023];
    __syncthreads();
    for (int g = warp; g < ngroups; g += 32) {
        const int32_t c0 = lane < NB ? count[g * NB + lane] : 0, c1 = lane + 32 < NB ? count[g * NB + lane + 32] : 0;
        int32_t s0 = c0, s1 = c1;                      // inclusive sums of the two halves of the shells
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t v0 = __shfl_up_sync(0xFFFFFFFFu, s0, o), v1 = __shfl_up_sync(0xFFFFFFFFu, s1, o);
            if (lane >= o) s0 += v0, s1 += v1;
        }
        const int32_t half = __shfl_sync(0xFFFFFFFFu, s0, 31), pos = gt[g] * tile;
        if (lane < NB) start[g * NB + lane] = cursor[g * NB + lane] = pos + s0 - c0;
        if (lane + 32 < NB) start[g * NB + lane + 32] = cursor[g * NB + lane + 32] = pos + half + s1 - c1;
    }
}

// group of every tile of a grouped order (the group of its first row; -1 for a tile of padding only)
__global__ void __launch_bounds__(256) tile_group_kernel(const uint32_t* __restrict__ perm, const int32_t* __restrict__ key,
                                                        int ntiles, int tile, int NB, int32_t* __restrict__ tgroup) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const uint32_t r = perm[(size_t)t * tile];
    tgroup[t] = r == GMD_EMPTY_IDX ? -1 : key[r] / NB;
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

// ---------------------------------------------------------------- grouped order (pruned scan)
extern "C" int gmd_group_shells(void) { return GMD_SHELLS; }

// sample count and leader capacity for N rows: ~8 samples per group of ~1000 rows, so that every
// cluster of that size is met by the sample
extern "C" void gmd_group_params(int64_t N, int* ns, int* cap) {
    int64_t s = N / 128;
    s = std::max<int64_t>(256, std::min<int64_t>(16384, s)) & ~(int64_t)31;
    if (s > N) s = N;
    *ns = (int)s;
    *cap = (int)std::max<int64_t>(1, std::min<int64_t>(4096, std::max<int64_t>(64, s / 4)));
    if (*cap > *ns) *cap = *ns;
}

namespace {
int lead_n1(int ns) { return std::min(ns, 2048); }       // pass A: the samples clustered before any leader is known
size_t lead_bits_words(int ns) {
    // the larger of the two bit matrices: pass A (n1 x n1) and pass C (at most ns - n1 uncovered samples)
    const size_t n1 = lead_n1(ns), n2 = (size_t)ns - n1;
    return std::max(n1 * ((n1 + 31) / 32), n2 * ((n2 + 31) / 32));
}
}  // namespace

extern "C" size_t gmd_group_leaders_workspace_bytes(int ns, int D) {
    // samples, nearest-sample distances, uncovered flags, list of uncovered samples, bit matrix
    return ((size_t)ns * D + LEAD_PROBES) * sizeof(float) + (((size_t)ns + 15) & ~(size_t)15) + (size_t)ns * sizeof(int32_t) +
           lead_bits_words(ns) * sizeof(uint32_t);
}

extern "C" int gmd_group_leaders_f64(const double* X, int64_t N, int D, const double* center, int ns, int cap, float rfac,
                                     float* L, int32_t* lead_row, double* lead_rows, int32_t* info, void* workspace,
                                     size_t workspace_bytes, int32_t* nlead_host, void* stream) {
    GMD_REQUIRE(X && L && lead_row && lead_rows && info && workspace && nlead_host, GMD_ERR_BAD_ARG,
                "group_leaders: null pointer");
    GMD_REQUIRE(N > 0 && D >= 1 && D <= GMD_MAX_D && ns >= 1 && ns <= N && ns <= 16384 && cap >= 1 && rfac > 0.0f,
                GMD_ERR_BAD_ARG, "group_leaders: bad sizes (N=%lld ns=%d cap=%d)", (long long)N, ns, cap);
    GMD_REQUIRE(workspace_bytes >= gmd_group_leaders_workspace_bytes(ns, D), GMD_ERR_WORKSPACE,
                "group_leaders: workspace %zu B < required %zu B", workspace_bytes,
                gmd_group_leaders_workspace_bytes(ns, D));
    cudaStream_t st = (cudaStream_t)stream;
    float* S = (float*)workspace;
    float* nn = S + (size_t)ns * D;
    unsigned char* flags = (unsigned char*)(nn + LEAD_PROBES);               // [ns]
    int32_t* list = (int32_t*)(flags + (((size_t)ns + 15) & ~(size_t)15));   // [ns]
    uint32_t* bits = (uint32_t*)(list + ns);
    const long long stride = N / ns;
    const int nprobe = std::min(ns, LEAD_PROBES);
    const long long ne = (long long)ns * D;
    gather_samples_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(X, stride, ns, D, center, S);
    GMD_CUDA_CHECK(cudaGetLastError());
    sample_nn_kernel<<<(unsigned)nprobe, 1024, 0, st>>>(S, ns, D, nprobe, nn);
    GMD_CUDA_CHECK(cudaGetLastError());
    radius_kernel<<<1, LEAD_PROBES, 0, st>>>(nn, nprobe, rfac, info);
    GMD_CUDA_CHECK(cudaGetLastError());
    const size_t adj_smem = (size_t)8 * D * sizeof(float);
    auto pass = [&](const int32_t* lst, const int32_t* n_dev, int n_max, int which) -> int {
        if (n_max <= 0) return GMD_OK;
        const int wpr = (n_max + 31) / 32;
        adjacency_kernel<<<(unsigned)((n_max + 7) / 8), 256, adj_smem, st>>>(S, lst, n_dev, n_max, D, info, bits, wpr);
        GMD_CUDA_CHECK(cudaGetLastError());
        const size_t rs = ((size_t)wpr + 64 * (size_t)(wpr | 1)) * sizeof(uint32_t);
        GMD_CUDA_CHECK(cudaFuncSetAttribute(resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs));
        resolve_kernel<<<1, LEAD_THREADS, rs, st>>>(S, lst, n_dev, n_max, D, bits, wpr, cap, stride, L, lead_row, info, which);
        GMD_CUDA_CHECK(cudaGetLastError());
        return GMD_OK;
    };
    const int n1 = lead_n1(ns);
    int rc;
    if ((rc = pass(nullptr, nullptr, n1, 0)) != GMD_OK) return rc;
    if (n1 < ns) {
        cover_kernel<<<(unsigned)(((size_t)(ns - n1) * 32 + 255) / 256), 256, 0, st>>>(S, ns, n1, D, L, info, flags);
        GMD_CUDA_CHECK(cudaGetLastError());
        uncovered_kernel<<<1, LEAD_THREADS, 0, st>>>(flags, ns, n1, list, info);
        GMD_CUDA_CHECK(cudaGetLastError());
        if ((rc = pass(list, info + 3, ns - n1, 1)) != GMD_OK) return rc;
    }
    int32_t h[2] = {0, 0};
    GMD_CUDA_CHECK(cudaMemcpyAsync(h, info, sizeof(h), cudaMemcpyDeviceToHost, st));
    GMD_CUDA_CHECK(cudaStreamSynchronize(st));
    GMD_REQUIRE(h[0] >= 1 && h[0] <= cap, GMD_ERR_CUDA, "group_leaders: %d leaders", h[0]);
    *nlead_host = h[0];
    gather_rows_kernel<<<(unsigned)((h[0] * D + 255) / 256), 256, 0, st>>>(X, lead_row, h[0], D, lead_rows);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

extern "C" size_t gmd_group_order_workspace_bytes(int ngroups) {
    return ngroups > 0 ? (size_t)(3 * (size_t)ngroups * gmd_group_shells() + 2) * sizeof(int32_t) : 0;
}

// Positions a grouped order of N rows in `ngroups` groups can take: every group is padded to whole tiles.
extern "C" int64_t gmd_group_positions(int64_t N, int ngroups, int tile) {
    return (N / tile + ngroups + 1) * (int64_t)tile;
}

extern "C" int gmd_group_order(const int32_t* lead, const double* dist, int64_t N, int ngroups, const int32_t* info,
                               int tile, int32_t* key, float* grad, uint32_t* perm, int64_t npos, int32_t* gstart,
                               int32_t* tgroup, void* workspace, size_t workspace_bytes, void* stream) {
    return gmd_group_order_pad(lead, dist, N, ngroups, info, tile, 1, key, grad, perm, npos, gstart, tgroup, workspace,
                               workspace_bytes, stream);
}

// pad: every group takes a multiple of `pad` tiles (a scan whose CTA owns `pad` neighbouring query tiles then never
// has a CTA straddling two groups); gstart and tgroup stay in units of one tile.
extern "C" int gmd_group_order_pad(const int32_t* lead, const double* dist, int64_t N, int ngroups, const int32_t* info,
                                   int tile, int pad, int32_t* key, float* grad, uint32_t* perm, int64_t npos,
                                   int32_t* gstart, int32_t* tgroup, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    GMD_REQUIRE(lead && dist && info && key && perm && gstart && tgroup && workspace, GMD_ERR_BAD_ARG,
                "group_order: null pointer");
    GMD_REQUIRE(pad >= 1 && pad <= 8, GMD_ERR_BAD_ARG, "group_order: bad pad %d", pad);
    GMD_REQUIRE(N > 0 && N < 0x7FFFFFFFll && ngroups >= 1 && tile >= 1 && npos >= gmd_group_positions(N, ngroups, tile * pad) &&
                    npos < 0xFFFFFFFFll,
                GMD_ERR_BAD_ARG, "group_order: bad sizes (N=%lld groups=%d positions=%lld)", (long long)N, ngroups,
                (long long)npos);
    GMD_REQUIRE(workspace_bytes >= gmd_group_order_workspace_bytes(ngroups), GMD_ERR_WORKSPACE,
                "group_order: workspace %zu B < required %zu B", workspace_bytes,
                gmd_group_order_workspace_bytes(ngroups));
    GMD_REQUIRE(ngroups <= SCAN_GROUPS_MAX && gmd_group_shells() <= 64, GMD_ERR_UNSUPPORTED,
                "group_order: %d groups exceed %d", ngroups, SCAN_GROUPS_MAX);
    cudaStream_t st = (cudaStream_t)stream;
    const int NB = gmd_group_shells(), nkeys = ngroups * NB;
    int32_t* count = (int32_t*)workspace;              // [nkeys]
    int32_t* start = count + nkeys;                    // [nkeys]
    int32_t* cursor = start + nkeys;                   // [nkeys]
    int32_t* bad = cursor + nkeys;                     // [1]
    GMD_CUDA_CHECK(cudaMemsetAsync(workspace, 0, gmd_group_order_workspace_bytes(ngroups), st));
    GMD_CUDA_CHECK(cudaMemsetAsync(perm, 0xFF, (size_t)npos * sizeof(uint32_t), st));
    if (grad) GMD_CUDA_CHECK(cudaMemsetAsync(grad, 0, (size_t)ngroups * sizeof(float), st));
    const unsigned grid = (unsigned)((N + 255) / 256);
    group_key_kernel<<<grid, 256, 0, st>>>(lead, dist, N, ngroups, NB, info, key, grad);
    GMD_CUDA_CHECK(cudaGetLastError());
    key_hist_kernel<<<grid, 256, 0, st>>>(key, N, nkeys, count, bad);
    GMD_CUDA_CHECK(cudaGetLastError());
    key_scan_padded_kernel<<<1, 1024, 0, st>>>(count, ngroups, NB, tile, pad, start, cursor, gstart);
    GMD_CUDA_CHECK(cudaGetLastError());
    key_scatter_kernel<<<grid, 256, 0, st>>>(key, N, nkeys, cursor, perm);
    GMD_CUDA_CHECK(cudaGetLastError());
    const int ntiles = (int)(npos / tile);
    tile_group_kernel<<<(ntiles + 255) / 256, 256, 0, st>>>(perm, key, ntiles, tile, NB, tgroup);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}
