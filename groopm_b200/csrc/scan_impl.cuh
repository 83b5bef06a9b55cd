// The hot kernel: FP32 distance tiles + fused warp-level top-kk selection.
//
// One CTA owns BQ = 32*NW query rows (resident in shared memory for the whole
// scan) and streams a contiguous range of candidate tiles through an NST-stage
// shared-memory ring filled by a dedicated producer warp with bulk async copies
// (cp.async.bulk -> SASS UBLKCP) that complete on mbarriers.  Each compute warp
// owns 32 of the CTA's queries outright -- its 32q x 64c warp tile is an 8x8
// register tile per lane -- so the running top-kk sets (shared-memory max-heaps of
// u64 keys = order-preserving(value) << 32 | index) are warp-private and the
// warps never synchronise with one another.  The fast path per candidate tile is
// D x 64 FFMA + a min-tree against the per-query threshold; only when some
// value beats a threshold does the warp take the insert path (hits are queued per
// query, then lane L updates the heap of the warp's query L).
//
// Candidate tiles are visited in ascending index order and the threshold test is
// strict, so equal values keep the lowest index (np.argmin's first-occurrence
// rule, torusMesh.py:127) no matter how the work is split.
#pragma once
#include "common.cuh"

namespace gmd_scan {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared, completing on an mbarrier (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

struct ScanParams {
    const float* qtiles;
    const float* ctiles;
    long long Nq, Nc;
    int ntiles_q, ntiles_c;
    int nsplit, kk;
    long long c_index_offset;
    uint32_t* cand_idx;
    float* cand_val;
    float* cand_tau;
};

__device__ __forceinline__ float sel8(const float (&a)[8], int j) {
    float v = a[0];
#pragma unroll
    for (int t = 1; t < 8; ++t) v = (j == t) ? a[t] : v;
    return v;
}

// Per-query running top-kk set: a binary MAX-heap of u64 keys in shared memory
// (root = current kk-th best = the threshold).  Each compute warp owns 32 queries
// and, on the rare path, lane L maintains the heap of the warp's query L by itself,
// so up to 32 inserts proceed in parallel with no shuffles; the float64 re-rank
// orders the survivors anyway.  Empty slots hold GMD_EMPTY_KEY (largest key), so a
// freshly initialised array is already a valid heap.
__device__ __forceinline__ void heap_replace_root(uint64_t* h, int kk, uint64_t key) {
    int pos = 0;
    while (true) {
        const int l = 2 * pos + 1;
        if (l >= kk) break;
        const uint64_t kl = h[l];
        const uint64_t kr = (l + 1 < kk) ? h[l + 1] : 0ull;
        const bool right = kr > kl;
        const uint64_t kc = right ? kr : kl;
        if (kc <= key) break;
        h[pos] = kc;
        pos = right ? l + 1 : l;
    }
    h[pos] = key;
}

constexpr int QCAP = 8;  // pending-hit queue slots per query and round

// d-loop shape: a straight-line prefix of PRE steps (d = 0 initialises the accumulators)
// then a rolled loop whose body is U steps, so the hot loop stays inside the
// instruction cache (a fully unrolled D=30 body is 34 KB of SASS and starves issue).
__host__ __device__ constexpr int pick_unroll(int D) {
    if (D <= 9) return D;  // whole thing straight-line
    int best_u = 8, best_rem = (D - 1) % 8;
    for (int u = 9; u >= 4; --u) {
        const int rem = (D - 1) % u;
        if (rem < best_rem) {
            best_rem = rem;
            best_u = u;
        }
    }
    return best_u;
}

struct Frag {
    float4 qa, qb, ca, cb;
};

template <int FORM, bool FIRST>
__device__ __forceinline__ void tile_step(float (&acc)[8][8], const Frag& f, const float (&cn)[8]) {
    const float q[8] = {f.qa.x, f.qa.y, f.qa.z, f.qa.w, f.qb.x, f.qb.y, f.qb.z, f.qb.w};
    const float c[8] = {f.ca.x, f.ca.y, f.ca.z, f.ca.w, f.cb.x, f.cb.y, f.cb.z, f.cb.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (FORM == GMD_FORM_EXPANDED) {
                acc[i][j] = fmaf(q[i], c[j], FIRST ? cn[j] : acc[i][j]);
            } else if (FORM == GMD_FORM_DIRECT_L2) {
                const float df = q[i] - c[j];
                acc[i][j] = FIRST ? df * df : fmaf(df, df, acc[i][j]);
            } else {
                const float df = fabsf(q[i] - c[j]);
                acc[i][j] = FIRST ? df : acc[i][j] + df;
            }
        }
}

template <int D, int FORM, int NW, int NST>
__global__ void __launch_bounds__((NW + 1) * 32, 1) scan_topk_kernel(const ScanParams p) {
    constexpr int BQ = NW * 32;
    constexpr int TILE_F = (D + 1) * GMD_TW;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr int U = pick_unroll(D);
    constexpr int PRE = (D <= 9) ? D : 1 + (D - 1) % U;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* Cs = reinterpret_cast<float*>(smem_raw);                // NST x [D+1][64]
    float* Qs = Cs + NST * TILE_F;                                 // [D][BQ]
    float* qn = Qs + D * BQ;                                       // [BQ]  ||q||^2 (form 0)
    float* thr_s = qn + BQ;                                        // [BQ]  heap-root value (+inf until full)
    int* qcnt = reinterpret_cast<int*>(thr_s + BQ);                // [BQ]  pending hits this round
    uint64_t* bars = reinterpret_cast<uint64_t*>(qcnt + BQ);       // full[NST], empty[NST]
    uint64_t* queue = bars + 2 * NST;                              // [BQ][QCAP] pending (value, index)
    uint64_t* lists = queue + BQ * QCAP;                           // [BQ][kkp]  max-heaps

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int kk = p.kk;
    const int kkp = kk | 1;  // odd stride: lanes walking their own heaps hit distinct banks
    const long long q0 = (long long)blockIdx.x * BQ;
    const int split = blockIdx.y;
    const int tps = (p.ntiles_c + p.nsplit - 1) / p.nsplit;
    const int tb = split * tps;
    const int te = min(p.ntiles_c, tb + tps);

    // ---- stage the CTA's queries: Qs[d][r] (x -2 for the expanded form), qn[r]
    {
        const float qscale = (FORM == GMD_FORM_EXPANDED) ? -2.0f : 1.0f;
        const int qt0 = (int)(q0 / GMD_TW);
        for (int e = tid; e < (D + 1) * BQ; e += (NW + 1) * 32) {
            const int d = e / BQ, r = e - d * BQ;
            const int t = qt0 + (r >> 6);
            float v = 0.0f;
            if (t < p.ntiles_q) v = p.qtiles[(size_t)t * TILE_F + d * GMD_TW + (r & 63)];
            if (d < D)
                Qs[d * BQ + r] = v * qscale;
            else
                qn[r] = v;
        }
        for (int e = tid; e < BQ * kkp; e += (NW + 1) * 32) lists[e] = GMD_EMPTY_KEY;
        for (int e = tid; e < BQ; e += (NW + 1) * 32) {
            thr_s[e] = __int_as_float(0x7f800000);
            qcnt[e] = 0;
        }
        if (tid == 0) {
            for (int s = 0; s < NST; ++s) {
                mbar_init(&bars[s], 1);
                mbar_init(&bars[NST + s], NW);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    __syncthreads();
    // no CTA-wide barrier below this line: the producer warp retires early

    if (warp == NW) {
        // ---- producer: keep the ring full
        if (lane == 0) {
            for (int t = tb; t < te; ++t) {
                const int it = t - tb, s = it % NST;
                if (it >= NST) mbar_wait(&bars[NST + s], ((it / NST) - 1) & 1);
                mbar_expect_tx(&bars[s], TILE_F * 4);
                bulk_g2s(Cs + s * TILE_F, p.ctiles + (size_t)t * TILE_F, TILE_F * 4, &bars[s]);
            }
        }
        return;
    }

    // ---- compute warps
    const int lq = lane >> 3, lc = lane & 7;
    const float* Qw = Qs + warp * 32 + lq * 4;
    float* wthr = thr_s + warp * 32;
    int* wcnt = qcnt + warp * 32;
    uint64_t* wqueue = queue + warp * 32 * QCAP;
    uint64_t* wheap = lists + (size_t)warp * 32 * kkp;
    float thr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) thr[i] = __int_as_float(0x7f800000);

    for (int t = tb; t < te; ++t) {
        const int it = t - tb, s = it % NST;
        mbar_wait(&bars[s], (it / NST) & 1);
        const float* Ct = Cs + s * TILE_F + lc * 4;
        auto load = [&](int d) {
            Frag f;
            f.qa = *reinterpret_cast<const float4*>(Qw + d * BQ);
            f.qb = *reinterpret_cast<const float4*>(Qw + d * BQ + 16);
            f.ca = *reinterpret_cast<const float4*>(Ct + d * GMD_TW);
            f.cb = *reinterpret_cast<const float4*>(Ct + d * GMD_TW + 32);
            return f;
        };

        float acc[8][8];
        float cn[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (FORM == GMD_FORM_EXPANDED) {
            const float4 na = *reinterpret_cast<const float4*>(Ct + D * GMD_TW);
            const float4 nb = *reinterpret_cast<const float4*>(Ct + D * GMD_TW + 32);
            cn[0] = na.x, cn[1] = na.y, cn[2] = na.z, cn[3] = na.w;
            cn[4] = nb.x, cn[5] = nb.y, cn[6] = nb.z, cn[7] = nb.w;
        }
        // one step ahead in registers; reading "row D" at the very end is harmless
        // (the norm row of the tile / the qn array follow the last coordinate row)
        Frag cur = load(0);
#pragma unroll
        for (int d = 0; d < PRE; ++d) {
            const Frag nxt = load(d + 1);
            if (d == 0)
                tile_step<FORM, true>(acc, cur, cn);
            else
                tile_step<FORM, false>(acc, cur, cn);
            cur = nxt;
        }
        if (PRE < D) {
#pragma unroll 1
            for (int d0 = PRE; d0 < D; d0 += U) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const Frag nxt = load(d0 + u + 1);
                    tile_step<FORM, false>(acc, cur, cn);
                    cur = nxt;
                }
            }
        }
        // this stage's shared memory is no longer needed by this warp
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[NST + s]);

        // ---- selection fast path: one min-tree per query row against its threshold
        unsigned rf = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float m01 = fminf(acc[i][0], acc[i][1]), m23 = fminf(acc[i][2], acc[i][3]);
            const float m45 = fminf(acc[i][4], acc[i][5]), m67 = fminf(acc[i][6], acc[i][7]);
            const float m = fminf(fminf(m01, m23), fminf(m45, m67));
            rf |= (m < thr[i]) ? (1u << i) : 0u;
        }
        if (__any_sync(FULL, rf != 0)) {
            // ---- rare path.  Phase 1: every lane queues its hits on the owning query;
            // phase 2: lane L drains query L's queue into its heap.  Rounds repeat while a
            // queue overflowed (only during list warm-up).
            const unsigned rows = __reduce_or_sync(FULL, rf);
            const uint32_t cbase = (uint32_t)t * GMD_TW + (lc << 2);
            unsigned pend[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                pend[i] = 0;
                if (rows & (1u << i)) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) pend[i] |= (acc[i][j] < thr[i]) ? (1u << j) : 0u;
                }
            }
            bool more;
            do {
                bool left = false;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (!(rows & (1u << i))) continue;  // warp-uniform
                    const int ql = ((i >> 2) << 4) + (lq << 2) + (i & 3);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (pend[i] & (1u << j)) {
                            const int slot = atomicAdd(&wcnt[ql], 1);
                            if (slot < QCAP) {
                                wqueue[ql * QCAP + slot] = gmd_key(acc[i][j], cbase + ((j >> 2) << 5) + (j & 3));
                                pend[i] &= ~(1u << j);
                            } else {
                                left = true;
                            }
                        }
                    }
                }
                __syncwarp();
                {
                    const int n = min(wcnt[lane], QCAP);
                    if (n > 0) {
                        uint64_t* h = wheap + (size_t)lane * kkp;
                        for (int r = 0; r < n; ++r) {
                            const uint64_t key = wqueue[lane * QCAP + r];
                            if (key < h[0]) heap_replace_root(h, kk, key);
                        }
                        const uint64_t root = h[0];
                        wthr[lane] = (root == GMD_EMPTY_KEY) ? __int_as_float(0x7f800000)
                                                             : gmd_unflip((uint32_t)(root >> 32));
                        wcnt[lane] = 0;
                    }
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) thr[i] = wthr[((i >> 2) << 4) + (lq << 2) + (i & 3)];
                more = __any_sync(FULL, left);
                if (more) {
                    // drop queued-out hits that no longer beat the refreshed thresholds
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (pend[i]) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (!(acc[i][j] < thr[i])) pend[i] &= ~(1u << j);
                        }
                    }
                }
            } while (more);
        }
    }

    // ---- write this warp's proposals: cand_*[q][split][kk], cand_tau[q][split]
    __syncwarp();
    for (int ql = 0; ql < 32; ++ql) {
        const long long q = q0 + warp * 32 + ql;
        if (q >= p.Nq) break;
        const uint64_t* h = wheap + (size_t)ql * kkp;
        const float qnv = qn[warp * 32 + ql];
        const size_t ob = ((size_t)q * p.nsplit + split) * kk;
        for (int x = lane; x < kk; x += 32) {
            const uint64_t key = h[x];
            const uint32_t li = (uint32_t)key;
            float v = gmd_unflip((uint32_t)(key >> 32));
            uint32_t oi = GMD_EMPTY_IDX;
            if (key != GMD_EMPTY_KEY && (long long)li < p.Nc) {
                oi = li + (uint32_t)p.c_index_offset;
                if (FORM == GMD_FORM_EXPANDED) v = fmaxf(v + qnv, 0.0f);
            } else {
                v = __int_as_float(0x7f800000);
            }
            p.cand_idx[ob + x] = oi;
            p.cand_val[ob + x] = v;
        }
        if (lane == 0) {
            // every candidate of this split that is NOT in the heap has an FP32 value >= tau
            float tau = wthr[ql];
            if (FORM == GMD_FORM_EXPANDED && h[0] != GMD_EMPTY_KEY) tau = fmaxf(tau + qnv, 0.0f);
            p.cand_tau[(size_t)q * p.nsplit + split] = tau;
        }
    }
}

template <int D, int FORM, int NW, int NST>
static size_t scan_smem_bytes(int kk) {
    constexpr int BQ = NW * 32;
    constexpr int TILE_F = (D + 1) * GMD_TW;
    return (size_t)(NST * TILE_F + D * BQ + 3 * BQ) * 4 + 2 * NST * 8 + (size_t)BQ * QCAP * 8 + (size_t)BQ * (kk | 1) * 8;
}

template <int D, int FORM, int NW, int NST>
static int scan_launch(const ScanParams& p, cudaStream_t st) {
    constexpr int BQ = NW * 32;
    const size_t smem = scan_smem_bytes<D, FORM, NW, NST>(p.kk);
    GMD_REQUIRE(smem <= 227 * 1024, GMD_ERR_UNSUPPORTED,
                "scan: D=%d kk=%d needs %zu B of shared memory (> 227 KB)", D, p.kk, smem);
    auto kern = scan_topk_kernel<D, FORM, NW, NST>;
    GMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((p.Nq + BQ - 1) / BQ), (unsigned)p.nsplit);
    kern<<<grid, (NW + 1) * 32, smem, st>>>(p);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

}  // namespace gmd_scan

// padded dimensions the scan is instantiated for; gmd_scan_pad_dim(D) is the smallest >= D
#define GMD_SCAN_DIMS(X) X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(20) X(24) X(28) X(30) X(32) X(40) X(48) X(64)
#define GMD_SCAN_DIMS_BIG(X) X(96) X(136)

int gmd_scan_pad_dim(int D);
int gmd_scan_launch_form0(int DP, const gmd_scan::ScanParams& p, cudaStream_t st);
int gmd_scan_launch_form1(int DP, const gmd_scan::ScanParams& p, cudaStream_t st);
int gmd_scan_launch_form2(int DP, const gmd_scan::ScanParams& p, cudaStream_t st);
int gmd_scan_block_queries(int DP);
