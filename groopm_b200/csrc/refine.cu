// float64 re-rank of the scan's candidates + exactness certificate + exact fallback.
//
// The FP32 scan only PROPOSES kk >= k candidates per (query, split).  Here every
// proposed pair is re-measured from the original float64 rows with the reference's
// own arithmetic -- sqrt(sum((a-b)^2)) / sum(|a-b|), direct differences, no FMA
// contraction (scipy cdist/pdist, e.g. torusMesh.py:127, cluster.py:495) -- and
// ranked by (distance, index), the np.argmin / lowest-index tie rule.
//
// Certificate.  A candidate the scan did NOT propose sits in some full split list
// whose kk-th FP32 value is tau_s, so its FP32 value is >= tau = min_s tau_s and its
// true float64 distance is >= lower(tau) where lower() subtracts a bound on
// (a) float64->float32 rounding of the centred rows, |d - d~| <= u(||q|| + ||c||),
// and (b) the accumulation error of the scan's form (for the tensor-core form: derived below without
// assumptions about the hardware's internal summation; tests/test_prune_gpu.py attacks it).  If the k-th float64
// distance is strictly below lower(tau) the top-k is provably the exact one.
// Otherwise the row goes to exact_rows_kernel, a float64 scan of every candidate.
#include "common.cuh"
#include <math_constants.h>

#include <stdlib.h>

#include <algorithm>

namespace {

constexpr double U32 = 5.9604644775390625e-08;  // 2^-24

// scipy's arithmetic: the terms in dimension order, every operation rounded on its own (no FMA).  The rows
// are read eight values at a time, all loads of a batch in flight together (a gathered candidate row
// costs an L2 round trip per dependent load otherwise), as 128-bit loads where the rows allow it.
__device__ __forceinline__ double dist64(const double* __restrict__ a, const double* __restrict__ b, int D,
                                         int metric) {
    double s = 0.0;
    int d = 0;
    const bool wide = (D % 2 == 0) && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) % 16 == 0);
    for (; d + 8 <= D; d += 8) {
        double av[8], bv[8];
        if (wide) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double2 x = *reinterpret_cast<const double2*>(a + d + 2 * j);
                const double2 y = *reinterpret_cast<const double2*>(b + d + 2 * j);
                av[2 * j] = x.x, av[2 * j + 1] = x.y, bv[2 * j] = y.x, bv[2 * j + 1] = y.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) av[j] = a[d + j], bv[j] = b[d + j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double t = __dsub_rn(av[j], bv[j]);
            s = __dadd_rn(s, metric == GMD_L2 ? __dmul_rn(t, t) : fabs(t));
        }
    }
    for (; d < D; ++d) {
        const double t = __dsub_rn(a[d], b[d]);
        s = __dadd_rn(s, metric == GMD_L2 ? __dmul_rn(t, t) : fabs(t));
    }
    return metric == GMD_L2 ? sqrt(s) : s;
}

__device__ __forceinline__ bool key_less(double da, uint32_t ia, double db, uint32_t ib) {
    return da < db || (da == db && ia < ib);
}

// G lanes per query row (G = 32, or 8 when there are at most 8 proposals per row: the k = 1
// scan), 32 / G rows per warp.
//
// NCH > 0 (G = 16, D even, D <= 32 NCH, rows 16-byte aligned) reads the candidate rows ACROSS the lanes of a
// group instead of one row per lane.  The gather of one row per lane is bound by the L1 tag stage (ncu, 1 M x 40
// proposals x 30 dimensions: l1tex 94 % busy, 13 sectors per request, one tag look-up per 16 bytes loaded = 15 per
// row); with lane p loading the 16 bytes of dimension pair p a row costs the 2-3 lines it spans.  The lane that
// loads a pair also holds the query's pair in registers and forms the term ((a - b)^2 or |a - b|); the terms of 16
// rows go through shared memory (row stride an odd number of 16-byte units: conflict-free both ways) to the lane
// that owns the row, which adds them in dimension order -- the same operations in the same order as dist64.
template <int G, int NCH, int METRIC>
__global__ void __launch_bounds__(128) refine_kernel(const double* __restrict__ Q, long long Nq,
                                                     const double* __restrict__ C, int D,
                                                     const double* __restrict__ center, int k, int kk, int nsplit,
                                                     int metric, int form, const uint32_t* __restrict__ cand_idx,
                                                     const float* __restrict__ cand_val,
                                                     const float* __restrict__ cand_tau,
                                                     const float* __restrict__ stats, int32_t* __restrict__ out_idx,
                                                     double* __restrict__ out_dist, int32_t* fail_count,
                                                     int32_t* __restrict__ fail_rows,
                                                     const uint32_t* __restrict__ qorder, long long nqpos) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    constexpr int RPW = 32 / G;  // rows per warp
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const int m = kk * nsplit;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, sl = lane % G;
    const int slot = warp * RPW + grp;                 // row slot inside the block
    const int nslots = (blockDim.x >> 5) * RPW;
    double* sd = reinterpret_cast<double*>(sm_raw) + (size_t)slot * m;
    uint32_t* si = reinterpret_cast<uint32_t*>(reinterpret_cast<double*>(sm_raw) + (size_t)nslots * m) + (size_t)slot * m;
    // qorder (nullable): the queries are visited in the scan's staging order (position -> row,
    // GMD_EMPTY_IDX = padding), so that consecutive blocks re-rank against the same neighbourhood of
    // candidate rows and find them in L2
    const long long qp = (long long)blockIdx.x * nslots + slot;
    long long q = qp;
    bool live = qp < Nq;                               // dead groups still take part in the shuffles
    if (qorder) {
        const uint32_t r = qp < nqpos ? qorder[qp] : GMD_EMPTY_IDX;
        live = r != GMD_EMPTY_IDX;
        q = live ? (long long)r : 0;
    }
    const double* qrow = Q + (size_t)(live ? q : 0) * D;
    const uint32_t* ci = cand_idx + (size_t)(live ? q : 0) * m;

    // the bound on everything that was not proposed: the smallest of the splits' taus
    float tau32 = CUDART_INF_F;
    if (live)
        for (int sp = sl; sp < nsplit; sp += G) tau32 = fminf(tau32, cand_tau[(size_t)q * nsplit + sp]);
    for (int o = G / 2; o > 0; o >>= 1) tau32 = fminf(tau32, __shfl_xor_sync(FULL, tau32, o));

    if constexpr (NCH > 0) {
        static_assert(NCH == 0 || G == 16, "the row-across-lanes gather is written for 16 lanes per query");
        const int P = D >> 1;                          // dimension pairs per row
        const int PS2 = 2 * (P | 1);                   // row stride of the term block, doubles
        size_t off = ((size_t)nslots * m * 3 + 1) / 2 * sizeof(double);
        off = (off + 15) & ~(size_t)15;
        double* tt = reinterpret_cast<double*>(sm_raw + off) + (size_t)slot * 16 * PS2;
        // lanes past the last pair (and empty proposals, and dead groups) load a valid address and drop the result
        double2 qv[NCH];
        const char* cb[NCH];
        unsigned keep[NCH];
        unsigned tw[NCH];
        const unsigned rowb = (unsigned)D * 8u;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int pp = sl + 16 * ch;
            keep[ch] = pp < P ? 1u : 0u;
            const int pc = keep[ch] ? pp : P - 1;
            qv[ch] = *reinterpret_cast<const double2*>(qrow + 2 * pc);
            cb[ch] = reinterpret_cast<const char*>(C + 2 * pc);
            tw[ch] = (unsigned)__cvta_generic_to_shared(tt + 2 * pc);
        }
        const float* cv = cand_val + (size_t)(live ? q : 0) * m;
        for (int b0 = 0; b0 < m; b0 += 16) {
            uint32_t my = GMD_EMPTY_IDX;
            if (live && b0 + sl < m) {
                my = ci[b0 + sl];
                if (nsplit > 1 && my != GMD_EMPTY_IDX && cv[b0 + sl] > tau32) my = GMD_EMPTY_IDX;
            }
            const uint32_t my_row = my == GMD_EMPTY_IDX ? 0u : my;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (b0 + 8 * h >= m) break;            // uniform
                double2 v[8][NCH];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const unsigned long long ro =
                        (unsigned long long)__shfl_sync(FULL, my_row, 8 * h + j, 16) * (unsigned long long)rowb;
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch)   // volatile: all the loads of a half issue before its first store
                        asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];"
                                     : "=d"(v[j][ch].x), "=d"(v[j][ch].y) : "l"(cb[ch] + ro));
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) {
                        const double t0 = __dsub_rn(qv[ch].x, v[j][ch].x), t1 = __dsub_rn(qv[ch].y, v[j][ch].y);
                        const double2 w = METRIC == GMD_L2 ? make_double2(__dmul_rn(t0, t0), __dmul_rn(t1, t1))
                                                           : make_double2(fabs(t0), fabs(t1));
                        // predicated, not branched around: the lanes past the last pair did the same arithmetic
                        asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; @p st.shared.v2.f64 [%1], {%2, %3}; }"
                                     :: "r"(keep[ch]), "r"(tw[ch] + (unsigned)((8 * h + j) * PS2 * 8)), "d"(w.x), "d"(w.y)
                                     : "memory");
                    }
                }
            }
            __syncwarp();
            double d = CUDART_INF;
            if (my != GMD_EMPTY_IDX) {
                const double2* row = reinterpret_cast<const double2*>(tt + (size_t)sl * PS2);
                double acc = 0.0;
#pragma unroll 5
                for (int pp = 0; pp < P; ++pp) {
                    const double2 w = row[pp];
                    acc = __dadd_rn(acc, w.x);
                    acc = __dadd_rn(acc, w.y);
                }
                d = METRIC == GMD_L2 ? sqrt(acc) : acc;
            }
            if (live && b0 + sl < m) {
                sd[b0 + sl] = d;
                si[b0 + sl] = my;
            }
            __syncwarp();
        }
    } else {
        if (live) {
            const float* cv = cand_val + (size_t)q * m;
            for (int x = sl; x < m; x += G) {
                uint32_t idx = ci[x];
                // With several splits a proposal of one split may lie above another split's tau: like anything
                // that was not proposed, its float64 distance is then >= lower(tau), it cannot be among a
                // certified top-k, and its row is not read.  (The split with the smallest tau holds kk >= k
                // proposals at or below it, so k finite entries always remain.)
                if (nsplit > 1 && idx != GMD_EMPTY_IDX && cv[x] > tau32) idx = GMD_EMPTY_IDX;
                double d = CUDART_INF;
                if (idx != GMD_EMPTY_IDX) d = dist64(qrow, C + (size_t)idx * D, D, metric);
                sd[x] = d;
                si[x] = idx;
            }
        }
        __syncwarp();
    }

    // ---- ranks.  Fast path: a lane keeps its (up to four) entries in registers and counts, for every entry of the
    // row, the distances strictly below its own; if the counts of the real entries are all different (an OR of
    // one bit per count) no two distances tie and the count IS the rank by (distance, index).  Rows with equal
    // distances (duplicate contigs) and rows with more than 64 proposals take the general loop.
    double kth = CUDART_INF;
    bool ranked = false;
    if (m <= 4 * G && m <= 64) {
        double dx[4];
        uint32_t ix[4];
        int rk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = sl + j * G;
            const bool in = live && x < m;
            dx[j] = in ? sd[x] : CUDART_INF;
            ix[j] = in ? si[x] : GMD_EMPTY_IDX;
            rk[j] = 0;
        }
        const int nx = (m + G - 1) / G;                // uniform
        if (nx <= 2) {
#pragma unroll 4
            for (int y = 0; y < m; ++y) {
                const double dy = sd[y];
#pragma unroll
                for (int j = 0; j < 2; ++j) rk[j] += dy < dx[j] ? 1 : 0;
            }
        } else if (nx == 3) {
#pragma unroll 4
            for (int y = 0; y < m; ++y) {
                const double dy = sd[y];
#pragma unroll
                for (int j = 0; j < 3; ++j) rk[j] += dy < dx[j] ? 1 : 0;
            }
        } else {
#pragma unroll 4
            for (int y = 0; y < m; ++y) {
                const double dy = sd[y];
#pragma unroll
                for (int j = 0; j < 4; ++j) rk[j] += dy < dx[j] ? 1 : 0;
            }
        }
        unsigned long long mask = 0ull;
        int real = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (ix[j] != GMD_EMPTY_IDX) mask |= 1ull << rk[j], ++real;
        unsigned lo = (unsigned)mask, hi = (unsigned)(mask >> 32);
        for (int o = G / 2; o > 0; o >>= 1) {
            lo |= __shfl_xor_sync(FULL, lo, o);
            hi |= __shfl_xor_sync(FULL, hi, o);
            real += __shfl_xor_sync(FULL, real, o);
        }
        ranked = __popc(lo) + __popc(hi) == real;      // uniform inside the group
        if (ranked && live) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (ix[j] != GMD_EMPTY_IDX && rk[j] < k) {
                    out_idx[(size_t)q * k + rk[j]] = (int32_t)ix[j];
                    out_dist[(size_t)q * k + rk[j]] = dx[j];
                    if (rk[j] == k - 1) kth = dx[j];
                }
        }
    }
    if (live && !ranked) {
        for (int x = sl; x < m; x += G) {
            const double d = sd[x];
            const uint32_t idx = si[x];
            int r = 0;
            for (int y = 0; y < m; ++y) r += key_less(sd[y], si[y], d, idx) ? 1 : 0;
            if (r < k && idx != GMD_EMPTY_IDX) {
                out_idx[(size_t)q * k + r] = (int32_t)idx;
                out_dist[(size_t)q * k + r] = d;
                if (r == k - 1) kth = d;
            }
        }
    }
    for (int o = G / 2; o > 0; o >>= 1) kth = fmin(kth, __shfl_xor_sync(FULL, kth, o));

    // ---- certificate
    double n2 = 0.0, n1 = 0.0;
    if (live) {
        for (int d = sl; d < D; d += G) {
            const double v = qrow[d] - (center ? center[d] : 0.0);
            n2 += v * v;
            n1 += fabs(v);
        }
    }
    for (int o = G / 2; o > 0; o >>= 1) {
        n2 += __shfl_xor_sync(FULL, n2, o);
        n1 += __shfl_xor_sync(FULL, n1, o);
    }
    bool ok = true;
    if (tau32 != CUDART_INF_F) {
        const double tau = (double)tau32;
        double lower;
        if (form == GMD_FORM_EXPANDED) {
            const double nn = (sqrt(n2) + sqrt((double)stats[0])) * 1.000001;
            const double e2 = (2.0 * D + 8.0) * U32 * nn * nn;
            lower = sqrt(fmax(tau - e2, 0.0)) - 1.01 * U32 * nn;
        } else if (form == GMD_FORM_TC) {
            // 3 x FP16 dot product of rows scaled by 2^e.  A bound that does not lean on how the tensor core
            // accumulates (that is undocumented): with v = hi + lo + r, |r| <= 2^-22 |v| (two roundings to an 11-bit
            // significand), the computed  hi.hi' + hi.lo' + lo.hi'  differs from q.c by
            //   the two representation residues and the dropped lo.lo' term   <= 3 x 2^-22 |q||c|
            //   FP32 accumulation: the 16-bit x 16-bit products are exact in FP32; if EVERY one of the 96 nkb
            //   additions into the accumulator rounded (or truncated: one ulp) on a partial sum <= |q||c|   <= 96 nkb 2^-23 |q||c|
            // (measured, tools/tc_probe16.cu: 2^-21.9 |q||c|, forty times less).  The ranked value is -2 x that, its norm
            // column carries 2^-22 ||c||^2, the write-out rounds twice more: with |q||c| <= nn^2 / 4 the sum is below
            // (2 (3 x 2^-22 + 96 nkb 2^-23) / 4 + 2^-21) nn^2 <= 128 nkb 2^-24 nn^2.  FP16 subnormals add an
            // absolute 2^-25 per element in scaled units: 2^-24 sqrt(D) nn 2^-e on the value ranked by.
            const double nn = (sqrt(n2) + sqrt((double)fmaxf(stats[0], stats[2]))) * 1.000001;
            const double floor_term = ldexp(sqrt((double)D) * nn, -24 - gmd_tc_exponent(stats[0], stats[2]));
            const double e2 = 128.0 * ((D + 32) / 32) * U32 * nn * nn + floor_term;
            lower = sqrt(fmax(tau - e2, 0.0)) - 1.01 * U32 * nn;
        } else if (form == GMD_FORM_DIRECT_L2) {
            const double nn = (sqrt(n2) + sqrt((double)stats[0])) * 1.000001;
            lower = sqrt(tau * (1.0 - (D + 6.0) * U32)) - 1.01 * U32 * nn;
        } else {
            const double nn = (n1 + (double)stats[1]) * 1.000001;
            lower = tau * (1.0 - (D + 4.0) * U32) - 1.01 * U32 * nn;
        }
        ok = kth < lower;
    }
    if (live && !ok && sl == 0) {
        const int fs = atomicAdd(fail_count, 1);
        fail_rows[fs] = (int32_t)q;
    }
}

// exact float64 scan of all candidates for the rows the certificate rejected
constexpr int EX_CAP = 2048;
constexpr int EX_THREADS = 1024;  // one block owns a row: memory-level parallelism is what it lacks

struct KeyD {
    double d;
    uint32_t i;
};

__device__ __forceinline__ KeyD block_min_key(KeyD v, KeyD* red) {
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xFFFFFFFFu, v.d, o);
        const uint32_t oi = __shfl_xor_sync(0xFFFFFFFFu, v.i, o);
        if (key_less(od, oi, v.d, v.i)) {
            v.d = od;
            v.i = oi;
        }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    KeyD r = red[0];
    for (int w = 1; w < EX_THREADS / 32; ++w)
        if (key_less(red[w].d, red[w].i, r.d, r.i)) r = red[w];
    return r;
}

// Batched form of the exact scan: one block per chunk of EXB_CH candidates (staged once, coalesced),
// every rejected row visits it, EXB_RB rows at a time.  The re-ranked k-th distance of a row is an
// upper bound on its true k-th distance, so only candidates within it are kept: they go to the row's
// hit list in global memory, which exact_finish_kernel ranks.  The candidate set is read once per
// launch instead of once per rejected row; rows whose list overflows (tie masses) and rows beyond
// the workspace's capacity are left to exact_rows_kernel.
constexpr int EXB_CH = 128;       // candidates per block
constexpr int EXB_RB = 8;         // rejected rows per pass: 4 per thread
constexpr int EXB_THREADS = 256;  // (candidate, half of the row group)
constexpr int EXB_HC = 256;       // hit list capacity per row
constexpr int EXB_FCAP = 4096;    // most rows taken by the batched pass

__global__ void __launch_bounds__(EXB_THREADS) exact_batch_kernel(const double* __restrict__ Q,
                                                                 const double* __restrict__ C, long long Nc, int D,
                                                                 int k, int metric,
                                                                 const int32_t* __restrict__ fail_count,
                                                                 const int32_t* __restrict__ fail_rows, int fcap,
                                                                 const double* __restrict__ out_dist,
                                                                 int32_t* __restrict__ hit_cnt,
                                                                 double* __restrict__ hit_d,
                                                                 uint32_t* __restrict__ hit_i) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int nfail = min(*fail_count, fcap);
    if (nfail == 0) return;
    const int Dp = D | 1;                                  // odd row pitch: conflict-free column walks
    double* cs = reinterpret_cast<double*>(sm_raw);        // [EXB_CH][Dp]
    double* qs = cs + (size_t)EXB_CH * Dp;                 // [EXB_RB][D]
    double* ubs = qs + (size_t)EXB_RB * D;                 // [EXB_RB]
    const int tid = threadIdx.x;
    const long long c0 = (long long)blockIdx.x * EXB_CH;
    const int n = (int)min((long long)EXB_CH, Nc - c0);
    const double* cg = C + (size_t)c0 * D;
    for (int e = tid; e < n * D; e += EXB_THREADS) {
        const int r = e / D;
        cs[r * Dp + (e - r * D)] = cg[e];
    }
    const int c = tid & (EXB_CH - 1), rh = tid / EXB_CH;   // rows rh * 4 .. rh * 4 + 3 of the group
    const double* cr = cs + (size_t)c * Dp;
    const double* q0 = qs + (size_t)(rh * 4) * D;
    for (int w0 = 0; w0 < nfail; w0 += EXB_RB) {
        __syncthreads();
        for (int e = tid; e < EXB_RB * D; e += EXB_THREADS) {
            const int r = e / D, w = w0 + r;
            qs[e] = w < nfail ? Q[(size_t)fail_rows[w] * D + (e - r * D)] : 0.0;
        }
        if (tid < EXB_RB) ubs[tid] = (w0 + tid) < nfail ? out_dist[(size_t)fail_rows[w0 + tid] * k + k - 1] : -1.0;
        __syncthreads();
        if (c >= n) continue;
        // scipy's arithmetic (dist64), four rows side by side
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        if (metric == GMD_L2) {
            for (int d = 0; d < D; ++d) {
                const double v = cr[d];
                const double t0 = __dsub_rn(q0[d], v), t1 = __dsub_rn(q0[D + d], v);
                const double t2 = __dsub_rn(q0[2 * D + d], v), t3 = __dsub_rn(q0[3 * D + d], v);
                s0 = __dadd_rn(s0, __dmul_rn(t0, t0));
                s1 = __dadd_rn(s1, __dmul_rn(t1, t1));
                s2 = __dadd_rn(s2, __dmul_rn(t2, t2));
                s3 = __dadd_rn(s3, __dmul_rn(t3, t3));
            }
            s0 = sqrt(s0), s1 = sqrt(s1), s2 = sqrt(s2), s3 = sqrt(s3);
        } else {
            for (int d = 0; d < D; ++d) {
                const double v = cr[d];
                s0 = __dadd_rn(s0, fabs(__dsub_rn(q0[d], v)));
                s1 = __dadd_rn(s1, fabs(__dsub_rn(q0[D + d], v)));
                s2 = __dadd_rn(s2, fabs(__dsub_rn(q0[2 * D + d], v)));
                s3 = __dadd_rn(s3, fabs(__dsub_rn(q0[3 * D + d], v)));
            }
        }
        const double sv[4] = {s0, s1, s2, s3};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int w = w0 + rh * 4 + j;
            if (w < nfail && sv[j] <= ubs[rh * 4 + j]) {
                const int slot = atomicAdd(&hit_cnt[w], 1);
                if (slot < EXB_HC) {
                    hit_d[(size_t)w * EXB_HC + slot] = sv[j];
                    hit_i[(size_t)w * EXB_HC + slot] = (uint32_t)(c0 + c);
                }
            }
        }
    }
}

// ranks a row's hit list by (distance, index) and writes its k best
__global__ void __launch_bounds__(EXB_HC) exact_finish_kernel(int k, const int32_t* __restrict__ fail_count,
                                                             const int32_t* __restrict__ fail_rows, int fcap,
                                                             const int32_t* __restrict__ hit_cnt,
                                                             const double* __restrict__ hit_d,
                                                             const uint32_t* __restrict__ hit_i,
                                                             int32_t* __restrict__ out_idx,
                                                             double* __restrict__ out_dist) {
    __shared__ double bd[EXB_HC];
    __shared__ uint32_t bi[EXB_HC];
    const int nfail = min(*fail_count, fcap);
    for (int w = blockIdx.x; w < nfail; w += gridDim.x) {
        const int n = hit_cnt[w];
        if (n > EXB_HC) continue;
        const long long q = fail_rows[w];
        __syncthreads();
        if ((int)threadIdx.x < n) {
            bd[threadIdx.x] = hit_d[(size_t)w * EXB_HC + threadIdx.x];
            bi[threadIdx.x] = hit_i[(size_t)w * EXB_HC + threadIdx.x];
        }
        __syncthreads();
        if ((int)threadIdx.x < n) {
            const double d = bd[threadIdx.x];
            const uint32_t idx = bi[threadIdx.x];
            int r = 0;
            for (int y = 0; y < n; ++y) r += key_less(bd[y], bi[y], d, idx) ? 1 : 0;
            if (r < k) {
                out_idx[(size_t)q * k + r] = (int32_t)idx;
                out_dist[(size_t)q * k + r] = d;
            }
        }
    }
}

__global__ void __launch_bounds__(EX_THREADS) exact_rows_kernel(const double* __restrict__ Q,
                                                               const double* __restrict__ C, long long Nc, int D,
                                                               int k, int metric,
                                                               const int32_t* __restrict__ fail_count,
                                                               const int32_t* __restrict__ fail_rows,
                                                               const int32_t* __restrict__ hit_cnt, int fcap,
                                                               int32_t* __restrict__ out_idx,
                                                               double* __restrict__ out_dist) {
    __shared__ double qs[GMD_MAX_D];
    __shared__ double bd[EX_CAP];
    __shared__ uint32_t bi[EX_CAP];
    __shared__ KeyD red[EX_THREADS / 32];
    __shared__ int cnt;
    const int nfail = *fail_count;
    for (int w = blockIdx.x; w < nfail; w += gridDim.x) {
        if (hit_cnt && w < fcap && hit_cnt[w] <= EXB_HC) continue;   // finished by the batched pass
        const long long q = fail_rows[w];
        __syncthreads();
        for (int d = threadIdx.x; d < D; d += EX_THREADS) qs[d] = Q[(size_t)q * D + d];
        if (threadIdx.x == 0) cnt = 0;
        __syncthreads();
        // the re-ranked k-th distance is a valid upper bound on the true k-th distance
        const double ub = out_dist[(size_t)q * k + k - 1];
        for (long long c = threadIdx.x; c < Nc; c += EX_THREADS) {
            const double d = dist64(qs, C + (size_t)c * D, D, metric);
            if (d <= ub) {
                const int slot = atomicAdd(&cnt, 1);
                if (slot < EX_CAP) {
                    bd[slot] = d;
                    bi[slot] = (uint32_t)c;
                }
            }
        }
        __syncthreads();
        const int n = cnt;
        if (n <= EX_CAP) {
            for (int x = threadIdx.x; x < n; x += EX_THREADS) {
                int r = 0;
                for (int y = 0; y < n; ++y) r += key_less(bd[y], bi[y], bd[x], bi[x]) ? 1 : 0;
                if (r < k) {
                    out_idx[(size_t)q * k + r] = (int32_t)bi[x];
                    out_dist[(size_t)q * k + r] = bd[x];
                }
            }
        } else {
            // pathological tie mass (> EX_CAP candidates within the bound): k selection passes
            KeyD prev = {-1.0, 0u};
            bool first = true;
            for (int r = 0; r < k; ++r) {
                KeyD best = {CUDART_INF, GMD_EMPTY_IDX};
                for (long long c = threadIdx.x; c < Nc; c += EX_THREADS) {
                    const double d = dist64(qs, C + (size_t)c * D, D, metric);
                    const bool after = first || key_less(prev.d, prev.i, d, (uint32_t)c);
                    if (after && key_less(d, (uint32_t)c, best.d, best.i)) {
                        best.d = d;
                        best.i = (uint32_t)c;
                    }
                }
                best = block_min_key(best, red);
                if (threadIdx.x == 0) {
                    out_idx[(size_t)q * k + r] = (int32_t)best.i;
                    out_dist[(size_t)q * k + r] = best.d;
                }
                prev = best;
                first = false;
            }
        }
    }
}

// One proposal per row, no certificate: the lowest (FP32 value, index) of the row's m proposals and the float64
// distance to THAT candidate (scipy's arithmetic).  For callers that need a good assignment with an exact distance
// to it, not the proven nearest (the grouped staging: any assignment gives valid balls).
__global__ void __launch_bounds__(256) pick_kernel(const double* __restrict__ X, long long N,
                                                   const double* __restrict__ C, int D, int m,
                                                   const uint32_t* __restrict__ cand_idx,
                                                   const float* __restrict__ cand_val, int32_t* __restrict__ out_idx,
                                                   double* __restrict__ out_dist) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= N) return;
    float bv = CUDART_INF_F;
    uint32_t bi = GMD_EMPTY_IDX;
    for (int x = 0; x < m; ++x) {
        const uint32_t i = cand_idx[(size_t)q * m + x];
        const float v = cand_val[(size_t)q * m + x];
        if (i != GMD_EMPTY_IDX && (bi == GMD_EMPTY_IDX || v < bv || (v == bv && i < bi))) bv = v, bi = i;
    }
    if (bi == GMD_EMPTY_IDX) bi = 0;                   // (cannot happen with at least one candidate)
    out_idx[q] = (int32_t)bi;
    out_dist[q] = dist64(X + (size_t)q * D, C + (size_t)bi * D, D, GMD_L2);
}

}  // namespace

extern "C" int gmd_pick_proposal_f64(const double* X, int64_t N, const double* C, int D, int m,
                                     const uint32_t* cand_idx, const float* cand_val, int32_t* out_idx,
                                     double* out_dist, void* stream) {
    GMD_REQUIRE(X && C && cand_idx && cand_val && out_idx && out_dist && N > 0 && D >= 1 && m >= 1, GMD_ERR_BAD_ARG,
                "pick_proposal: bad arguments");
    pick_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, N, C, D, m, cand_idx, cand_val, out_idx,
                                                                             out_dist);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

extern "C" int gmd_refine_topk_order_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D,
                                         const double* center, int k, int kk, int nsplit, int metric, int form,
                                         const uint32_t* cand_idx, const float* cand_val, const float* cand_tau,
                                         const float* stats, int32_t* out_idx, double* out_dist, int32_t* fail_count,
                                         int32_t* fail_rows, const uint32_t* qorder, int64_t nqpos, void* stream) {
    GMD_REQUIRE(Q && C && cand_idx && cand_val && cand_tau && stats && out_idx && out_dist && fail_count && fail_rows,
                GMD_ERR_BAD_ARG, "refine: null pointer");
    GMD_REQUIRE(Nq > 0 && Nc > 0 && D >= 1 && k >= 1 && kk >= k && nsplit >= 1 && k <= Nc, GMD_ERR_BAD_ARG,
                "refine: bad sizes (Nq=%lld Nc=%lld D=%d k=%d kk=%d nsplit=%d)", (long long)Nq, (long long)Nc, D, k,
                kk, nsplit);
    const int wpb = 4;
    // lanes per query row: 8 for the k = 1 scan's 8 proposals; otherwise the group size that leaves the fewest
    // idle lanes over the ceil(m / G) rounds of the gather (40 proposals: 16 lanes x 3 rounds, not 32 x 2)
    const int m = kk * nsplit;
    int G = 32;
    if (m <= 8) {
        G = 8;
    } else {
        int best = 1 << 30;
        for (int g : {16, 8, 32}) {
            // (16 lanes measured a little ahead of 8 at equal work: 2.66 vs 2.77 ms for 1M x 40 proposals)
            const int waste = (m + g - 1) / g * g - m - (g == 16 ? 8 : 0);
            if (waste < best) best = waste, G = g;
        }
    }
    if (const char* env = getenv("GMD_REFINE_G")) {    // developer switch
        const int g = atoi(env);
        if (g == 8 || g == 16 || g == 32) G = g;
    }
    // rows read across the lanes of a group (see refine_kernel): even D up to 64, 16-byte aligned rows
    int nch = 0;
    // (not for the k = 1 scan's 8 proposals: half the lanes of a group would idle; measured slower)
    if (m > 8 && D % 2 == 0 && D <= 64 && ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(C)) % 16) == 0)
        nch = D <= 32 ? 1 : 2;
    if (const char* env = getenv("GMD_REFINE_ACROSS"))  // developer switch
        if (atoi(env) == 0) nch = 0;
    if (nch) G = 16;
    const int rows_per_block = wpb * (32 / G);
    size_t smem = (size_t)rows_per_block * kk * nsplit * (sizeof(double) + sizeof(uint32_t));
    if (nch) {
        smem = ((size_t)rows_per_block * m * 3 + 1) / 2 * sizeof(double);
        smem = (smem + 15) & ~(size_t)15;
        smem += (size_t)rows_per_block * 16 * 2 * ((D / 2) | 1) * sizeof(double);
    }
    GMD_REQUIRE(smem <= 200 * 1024, GMD_ERR_UNSUPPORTED, "refine: kk*nsplit=%d too large", kk * nsplit);
    const bool l2 = metric == GMD_L2;
    auto kern = nch == 1 ? (l2 ? refine_kernel<16, 1, GMD_L2> : refine_kernel<16, 1, GMD_L1>)
              : nch == 2 ? (l2 ? refine_kernel<16, 2, GMD_L2> : refine_kernel<16, 2, GMD_L1>)
              : G == 8   ? refine_kernel<8, 0, -1>
              : G == 16  ? refine_kernel<16, 0, -1>
                         : refine_kernel<32, 0, -1>;
    GMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nslots = qorder ? nqpos : Nq;
    GMD_REQUIRE(nslots > 0, GMD_ERR_BAD_ARG, "refine: empty query order");
    const unsigned grid = (unsigned)((nslots + rows_per_block - 1) / rows_per_block);
    kern<<<grid, wpb * 32, smem, (cudaStream_t)stream>>>(Q, Nq, C, D, center, k, kk, nsplit, metric, form, cand_idx,
                                                        cand_val, cand_tau, stats, out_idx, out_dist, fail_count,
                                                        fail_rows, qorder, nqpos);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

extern "C" int gmd_refine_topk_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D,
                                   const double* center, int k, int kk, int nsplit, int metric, int form,
                                   const uint32_t* cand_idx, const float* cand_val, const float* cand_tau,
                                   const float* stats, int32_t* out_idx, double* out_dist, int32_t* fail_count, int32_t* fail_rows,
                                   void* stream) {
    return gmd_refine_topk_order_f64(Q, Nq, C, Nc, D, center, k, kk, nsplit, metric, form, cand_idx, cand_val, cand_tau,
                                     stats, out_idx, out_dist, fail_count, fail_rows, nullptr, 0, stream);
}

extern "C" size_t gmd_exact_workspace_bytes(int64_t Nq) {
    const size_t rows = (size_t)std::min<int64_t>(std::max<int64_t>(Nq, 0), EXB_FCAP);
    return rows * (sizeof(int32_t) + (size_t)EXB_HC * (sizeof(double) + sizeof(uint32_t)));
}

extern "C" int gmd_exact_rows_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D, int k, int metric,
                                  const int32_t* fail_count, const int32_t* fail_rows, int32_t* out_idx,
                                  double* out_dist, void* workspace, size_t workspace_bytes, void* stream) {
    GMD_REQUIRE(Q && C && fail_count && fail_rows && out_idx && out_dist, GMD_ERR_BAD_ARG, "exact_rows: null pointer");
    GMD_REQUIRE(Nq > 0 && Nc > 0 && D >= 1 && D <= GMD_MAX_D && k >= 1 && k <= Nc, GMD_ERR_BAD_ARG,
                "exact_rows: bad sizes");
    int dev = 0, sms = 0;
    GMD_CUDA_CHECK(cudaGetDevice(&dev));
    GMD_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t st = (cudaStream_t)stream;
    // batched pass over as many rejected rows as the workspace holds hit lists for
    const size_t per_row = sizeof(int32_t) + (size_t)EXB_HC * (sizeof(double) + sizeof(uint32_t));
    const int fcap = workspace ? (int)std::min<size_t>(std::min<int64_t>(Nq, EXB_FCAP), workspace_bytes / per_row) : 0;
    int32_t* hit_cnt = nullptr;
    if (fcap > 0) {
        double* hit_d = (double*)workspace;                                  // [fcap][EXB_HC]
        uint32_t* hit_i = (uint32_t*)(hit_d + (size_t)fcap * EXB_HC);        // [fcap][EXB_HC]
        hit_cnt = (int32_t*)(hit_i + (size_t)fcap * EXB_HC);                 // [fcap]
        GMD_CUDA_CHECK(cudaMemsetAsync(hit_cnt, 0, sizeof(int32_t) * fcap, st));
        const size_t smem = ((size_t)EXB_CH * (D | 1) + (size_t)EXB_RB * D + EXB_RB) * sizeof(double);
        GMD_CUDA_CHECK(cudaFuncSetAttribute(exact_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)((Nc + EXB_CH - 1) / EXB_CH);
        exact_batch_kernel<<<grid, EXB_THREADS, smem, st>>>(Q, C, Nc, D, k, metric, fail_count, fail_rows, fcap,
                                                            out_dist, hit_cnt, hit_d, hit_i);
        GMD_CUDA_CHECK(cudaGetLastError());
        exact_finish_kernel<<<sms * 2, EXB_HC, 0, st>>>(k, fail_count, fail_rows, fcap, hit_cnt, hit_d, hit_i, out_idx,
                                                       out_dist);
        GMD_CUDA_CHECK(cudaGetLastError());
    }
    exact_rows_kernel<<<sms * 2, EX_THREADS, 0, st>>>(Q, C, Nc, D, k, metric, fail_count, fail_rows, hit_cnt, fcap,
                                                      out_idx, out_dist);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}
