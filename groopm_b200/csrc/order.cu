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
__global__ void __launch_bounds__(256) sample_nn_kernel(const float* __restrict__ S, int ns, int D, int nprobe,
                                                       float* __restrict__ nn) {
    __shared__ float part[8];
    const int w = blockIdx.x;
    const int p = w * (ns / nprobe);
    const float* a = S + (size_t)p * D;
    float best = __int_as_float(0x7f800000);
    for (int j = threadIdx.x; j < ns; j += blockDim.x) {
        const float* b = S + (size_t)j * D;
        float d2 = 0.0f;
#pragma unroll 6
        for (int d = 0; d < D; ++d) {
            const float t = a[d] - b[d];
            d2 = fmaf(t, t, d2);
        }
        if (d2 > 0.0f) best = fminf(best, d2);
    }
    for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int x = 1; x < 8; ++x) best = fminf(best, part[x]);
        nn[w] = sqrtf(best);
    }
}

// Leader clustering of the samples: a sample becomes a leader unless an earlier leader lies within
// R = rfac x median(nn) of it (at most cap leaders).  Greedy and order dependent, hence one block -- but only
// for the part that has to be: pass A clusters the first n1 samples (all their candidate leaders are
// unknown), cover_kernel then tests every other sample against pass A's leaders on the whole GPU, and pass C
// clusters the samples that are still uncovered, testing them against pass C's own leaders only.
// Inside a pass the samples go through in batches of B: all threads test the batch against the leaders
// that exist (thread = one sample x a share of the leaders; the batch sits dimension-major in shared
// memory, the leader rows are warp-wide broadcasts), then warp 0 walks the uncovered samples in order and
// tests each against the leaders the batch itself has produced so far.
// out: L [cap][D] centred FP32 leader rows, lead_row[cap] (row of X), info[0] = leaders,
// info[1] = float bits of R.
// pass A: flags null, samples [0, n1), j0 = 0, computes R.  pass C: samples {i >= n1 : flags[i] != 0} in order
// (compacted by the block first), tested against the leaders from index info[2] on (pass A's count).
__global__ void __launch_bounds__(LEAD_THREADS) leader_kernel(const float* __restrict__ S, int ns, int n1,
                                                             long long stride, int D, const float* __restrict__ nn,
                                                             int nprobe, float rfac, int cap, int B, int capS,
                                                             const unsigned char* __restrict__ flags,
                                                             int32_t* __restrict__ list, float* __restrict__ L,
                                                             int32_t* __restrict__ lead_row,
                                                             int32_t* __restrict__ info) {
    extern __shared__ float lsm[];
    float* sT = lsm;                                   // [D][B] the batch, dimension-major
    int* cov = reinterpret_cast<int*>(sT + (size_t)D * B);   // [B]
    int* fresh = cov + B;                              // [B] batch members that became leaders (batch index)
    int* sidx = fresh + B;                             // [B] sample index of a batch member
    // the first capS leaders are mirrored in shared memory: the global rows are written while the kernel runs,
    // and a freshly written line is an L2 round trip for every reader
    float* Ls = reinterpret_cast<float*>(sidx + B);    // [capS][D]
    __shared__ float vals[LEAD_PROBES];
    __shared__ int part[LEAD_THREADS];
    __shared__ int nlead_s, nlist_s;
    __shared__ float R2_s;
    const int tid = threadIdx.x, lane = tid & 31;
    int j0 = 0, nlist;
    if (!flags) {
        // ---- pass A: median of the finite nearest-sample distances, by rank counting
        const bool probe = tid < nprobe;
        const float v = probe ? nn[tid] : __int_as_float(0x7f800000);
        if (tid < LEAD_PROBES) vals[tid] = v;
        if (tid == 0) nlead_s = 0, R2_s = 0.0f;
        const int nfin = __syncthreads_count(probe && v < __int_as_float(0x7f800000));
        if (probe && nfin > 0) {
            int rank = 0;
            for (int j = 0; j < nprobe; ++j) rank += (vals[j] < v || (vals[j] == v && j < tid)) ? 1 : 0;
            if (rank == nfin / 2) {
                const float R = rfac * v;
                R2_s = R * R;
                info[1] = __float_as_int(R);
            }
        } else if (tid == 0 && nfin == 0) {
            info[1] = 0;
        }
        nlist = n1;
    } else {
        // ---- pass C: ordered compaction of the uncovered samples
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
        if (tid == LEAD_THREADS - 1) nlist_s = part[tid];
        if (tid == 0) {
            nlead_s = info[0];
            const float R = __int_as_float(info[1]);
            R2_s = R * R;
        }
        __syncthreads();
        nlist = nlist_s;
        j0 = info[0];
        for (int e = tid; e < min(j0, capS) * D; e += LEAD_THREADS) Ls[e] = L[e];   // pass A's leaders
    }
    __syncthreads();
    const float R2 = R2_s;
    const int shares = LEAD_THREADS / B;               // threads per sample, each takes every shares-th leader
    const int b = tid % B, share = tid / B;
    for (int i0 = 0; i0 < nlist; i0 += B) {
        const int nb = min(B, nlist - i0);
        for (int e = tid; e < B; e += LEAD_THREADS) {
            cov[e] = 0;
            if (e < nb) sidx[e] = flags ? list[i0 + e] : i0 + e;
        }
        __syncthreads();
        for (int e = tid; e < nb * D; e += LEAD_THREADS) {
            const int r = e / D, d = e - r * D;
            sT[(size_t)d * B + r] = S[(size_t)sidx[r] * D + d];
        }
        __syncthreads();
        const int nl = nlead_s;
        if (b < nb) {
            bool hit = false;
            for (int j = j0 + share; j < nl && !hit; j += shares) {
                const float* l = j < capS ? Ls + (size_t)j * D : L + (size_t)j * D;   // one row for the whole warp: broadcast
                float d2 = 0.0f;
#pragma unroll 6
                for (int d = 0; d < D; ++d) {
                    const float t = sT[(size_t)d * B + b] - l[d];
                    d2 = fmaf(t, t, d2);
                }
                hit = d2 <= R2;
            }
            if (hit) cov[b] = 1;
        }
        __syncthreads();
        if (tid < 32) {
            int nf = 0;                                // leaders this batch has produced (warp-uniform)
            for (int r = 0; r < nb; ++r) {
                if (cov[r]) continue;
                if (nlead_s + nf >= cap) break;
                bool near = false;
                for (int x = lane; x < nf; x += 32) {
                    const int o = fresh[x];
                    float d2 = 0.0f;
#pragma unroll 6
                    for (int d = 0; d < D; ++d) {
                        const float t = sT[(size_t)d * B + r] - sT[(size_t)d * B + o];
                        d2 = fmaf(t, t, d2);
                    }
                    near |= d2 <= R2;
                }
                if (__any_sync(0xFFFFFFFFu, near)) continue;
                const int j = nlead_s + nf;
                for (int d = lane; d < D; d += 32) {
                    const float v = sT[(size_t)d * B + r];
                    L[(size_t)j * D + d] = v;
                    if (j < capS) Ls[(size_t)j * D + d] = v;
                }
                if (lane == 0) {
                    lead_row[j] = (int32_t)((long long)sidx[r] * stride);
                    fresh[nf] = r;
                }
                __syncwarp();
                ++nf;
            }
            if (lane == 0) nlead_s += nf;
        }
        __syncthreads();
    }
    if (tid == 0) {
        info[0] = nlead_s;
        if (!flags) info[2] = nlead_s;
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
        const double s = d / (0.5 * (double)R);
        b = s < (double)(NB - 1) ? (int)s : NB - 1;
    } else if (d == 0.0) {
        b = 0;
    }
    key[i] = g * NB + b;
    if (grad && d == d) atomicMax(reinterpret_cast<int*>(grad) + g, __float_as_int(__double2float_ru(d)));
}

// One block.  Padded starts of a (group, shell) counting sort: group g's rows begin on a tile boundary.
// start / cursor [nkeys], gstart[ngroups + 1] in TILES; gstart[ngroups] * tile = positions used.
__global__ void __launch_bounds__(1024) key_scan_padded_kernel(const int32_t* __restrict__ count, int ngroups, int NB,
                                                               int tile, int32_t* __restrict__ start,
                                                               int32_t* __restrict__ cursor,
                                                               int32_t* __restrict__ gstart) {
    __shared__ int32_t part[1024];
    const int tid = threadIdx.x;
    const int per = (ngroups + 1023) / 1024;
    const int a = min(ngroups, tid * per), b = min(ngroups, a + per);
    int32_t sum = 0;                                   // padded sizes of this thread's groups, in tiles
    for (int g = a; g < b; ++g) {
        int32_t c = 0;
        for (int x = 0; x < NB; ++x) c += count[g * NB + x];
        sum += (c + tile - 1) / tile;
    }
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
        gstart[g] = run;
        int32_t pos = run * tile, c = 0;
        for (int x = 0; x < NB; ++x) {
            start[g * NB + x] = pos;
            cursor[g * NB + x] = pos;
            pos += count[g * NB + x];
            c += count[g * NB + x];
        }
        run += (c + tile - 1) / tile;
    }
    if (tid == 1023) gstart[ngroups] = part[1023];
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
extern "C" int gmd_group_shells(void) { return 8; }

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

extern "C" size_t gmd_group_leaders_workspace_bytes(int ns, int D) {
    // samples, nearest-sample distances, uncovered flags, list of uncovered samples
    return ((size_t)ns * D + LEAD_PROBES) * sizeof(float) + (((size_t)ns + 15) & ~(size_t)15) + (size_t)ns * sizeof(int32_t);
}

extern "C" int gmd_group_leaders_f64(const double* X, int64_t N, int D, const double* center, int ns, int cap, float rfac,
                                     float* L, int32_t* lead_row, double* lead_rows, int32_t* info, void* workspace,
                                     size_t workspace_bytes, int32_t* nlead_host, void* stream) {
    GMD_REQUIRE(X && L && lead_row && lead_rows && info && workspace && nlead_host, GMD_ERR_BAD_ARG,
                "group_leaders: null pointer");
    GMD_REQUIRE(N > 0 && D >= 1 && D <= GMD_MAX_D && ns >= 1 && ns <= N && cap >= 1 && rfac > 0.0f, GMD_ERR_BAD_ARG,
                "group_leaders: bad sizes (N=%lld ns=%d cap=%d)", (long long)N, ns, cap);
    GMD_REQUIRE(workspace_bytes >= gmd_group_leaders_workspace_bytes(ns, D), GMD_ERR_WORKSPACE,
                "group_leaders: workspace %zu B < required %zu B", workspace_bytes,
                gmd_group_leaders_workspace_bytes(ns, D));
    cudaStream_t st = (cudaStream_t)stream;
    float* S = (float*)workspace;
    float* nn = S + (size_t)ns * D;
    const long long stride = N / ns;
    const int nprobe = std::min(ns, LEAD_PROBES);
    const long long ne = (long long)ns * D;
    gather_samples_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, st>>>(X, stride, ns, D, center, S);
    GMD_CUDA_CHECK(cudaGetLastError());
    sample_nn_kernel<<<(unsigned)nprobe, 256, 0, st>>>(S, ns, D, nprobe, nn);
    GMD_CUDA_CHECK(cudaGetLastError());
    // batch: [D][B] floats of shared memory.  Pass A meets mostly new clusters -- its cost is warp 0's walk over
    // the batch against the leaders the batch has produced, so its batches are small; pass C meets mostly
    // members of the few clusters pass A missed, which the parallel test against pass C's leaders covers.
    const int B = D <= 48 ? 256 : (D <= 96 ? 128 : 64), BA = 32;
    const size_t batch_b = ((size_t)D * B + 3 * (size_t)B) * sizeof(float);
    const int capS = (int)std::min<size_t>((size_t)cap, (200 * 1024 - batch_b) / ((size_t)D * sizeof(float)));
    const size_t lsm = batch_b + (size_t)capS * D * sizeof(float);
    GMD_CUDA_CHECK(cudaFuncSetAttribute(leader_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm));
    unsigned char* flags = (unsigned char*)(nn + LEAD_PROBES);  // [ns]
    int32_t* list = (int32_t*)(flags + (((size_t)ns + 15) & ~(size_t)15));   // [ns]
    const int n1 = std::min(ns, 1024);                        // pass A: the samples clustered before any is known
    leader_kernel<<<1, LEAD_THREADS, lsm, st>>>(S, ns, n1, stride, D, nn, nprobe, rfac, cap, BA, capS, nullptr, list, L,
                                               lead_row, info);
    GMD_CUDA_CHECK(cudaGetLastError());
    if (n1 < ns) {
        cover_kernel<<<(unsigned)(((size_t)(ns - n1) * 32 + 255) / 256), 256, 0, st>>>(S, ns, n1, D, L, info, flags);
        GMD_CUDA_CHECK(cudaGetLastError());
        leader_kernel<<<1, LEAD_THREADS, lsm, st>>>(S, ns, n1, stride, D, nn, nprobe, rfac, cap, B, capS, flags, list, L,
                                                   lead_row, info);
        GMD_CUDA_CHECK(cudaGetLastError());
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
    GMD_REQUIRE(lead && dist && info && key && perm && gstart && tgroup && workspace, GMD_ERR_BAD_ARG,
                "group_order: null pointer");
    GMD_REQUIRE(N > 0 && N < 0x7FFFFFFFll && ngroups >= 1 && tile >= 1 && npos >= gmd_group_positions(N, ngroups, tile) &&
                    npos < 0xFFFFFFFFll,
                GMD_ERR_BAD_ARG, "group_order: bad sizes (N=%lld groups=%d positions=%lld)", (long long)N, ngroups,
                (long long)npos);
    GMD_REQUIRE(workspace_bytes >= gmd_group_order_workspace_bytes(ngroups), GMD_ERR_WORKSPACE,
                "group_order: workspace %zu B < required %zu B", workspace_bytes,
                gmd_group_order_workspace_bytes(ngroups));
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
    key_scan_padded_kernel<<<1, 1024, 0, st>>>(count, ngroups, NB, tile, start, cursor, gstart);
    GMD_CUDA_CHECK(cudaGetLastError());
    key_scatter_kernel<<<grid, 256, 0, st>>>(key, N, nkeys, cursor, perm);
    GMD_CUDA_CHECK(cudaGetLastError());
    const int ntiles = (int)(npos / tile);
    tile_group_kernel<<<(ntiles + 255) / 256, 256, 0, st>>>(perm, key, ntiles, tile, NB, tgroup);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}
