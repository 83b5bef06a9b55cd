// The hot kernel: FP32 distance tiles + fused warp-level top-kk selection.
//
// One CTA owns BQ = 32*NW query rows (resident in shared memory for the whole scan) and
// streams a contiguous range of 64-candidate tiles through an NST-stage shared-memory ring.
// A dedicated producer warp fills the ring with bulk async copies (cp.async.bulk -> SASS
// UBLKCP, the TMA engine without a tensor map) that complete on mbarriers; a stage is one
// (tile, dimension chunk), so the ring is the same size for D = 4 and D = 136.  Each compute
// warp owns 32 of the CTA's queries outright: its 32q x 64c warp tile is an 8q x 8c register
// tile per lane, accumulated with Blackwell's packed FP32 FMA (fma.rn.f32x2 -> SASS FFMA2,
// two FMAs per issue slot, the query operand broadcast from one 32-bit register).  The
// running top-kk sets (shared-memory max-heaps of u64 keys = order-preserving(value) << 32 |
// index) are warp-private, so compute warps never synchronise with one another.
//
// Fast path per tile: D x 32 FFMA2 + a min-tree against the per-query threshold.  A lane
// whose tile holds a value under a threshold appends (value, index) to that query's
// shared-memory queue and carries on; the queues are folded into the heaps (lane L owns
// query L) only when one of them is filling up, so heap maintenance is amortised over
// many hits and most tiles with a hit cost one atomic and one store.
//
// Register-tile shape and warp count were chosen by measurement (tools/probe_tile.cu,
// tools/probe_ffma2.cu): 8x8 with 8+1 warps keeps the compute warps
// under the 168-register cap of a 9-warp CTA; 8x16 tiles do not run faster.
//
// Candidates are visited in ascending index order and the threshold test is strict, so
// equal values keep the lowest index (np.argmin's first-occurrence rule,
// torusMesh.py:127) no matter how the work is split.
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
// The producer's wait on an EMPTY barrier: the ring is normally full, so this wait lasts a
// whole tile time; sleep between polls instead of spinning on the issue port that the
// producer shares with two compute warps.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}
// 1-D bulk async copy global -> shared, completing on an mbarrier (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

#define GMD_SMEM_LIMIT (227 * 1024)

struct ScanParams {
    const float* qtiles;
    const float* ctiles;
    long long Nq, Nc;
    int ntiles_q, ntiles_c;  // allocated tile counts
    int nsplit, kk;
    long long c_index_offset;
    uint32_t* cand_idx;
    float* cand_val;
    float* cand_tau;
};

// Per-query running top-kk set: a binary MAX-heap of u64 keys in shared memory
// (root = current kk-th best = the threshold).  On the rare path lane L maintains the heap
// of the warp's query L by itself, so up to 32 inserts proceed in parallel with no
// shuffles; the float64 re-rank orders the survivors anyway.  Empty slots hold
// GMD_EMPTY_KEY (largest key), so a freshly initialised array is already a valid heap.
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

constexpr int QCAP = 16;   // pending-hit queue slots per query
constexpr int QHIGH = 8;   // a queue this full asks for a drain of the warp's queues

// Dimension chunk per stage: whole rows up to 64 dimensions, else an even split
// (96 = 3 x 32, 136 = 4 x 34) so the accumulators persist across the chunks of one tile pair.
__host__ __device__ constexpr int dim_chunk(int D) {
    if (D <= 64) return D;
    for (int c = 34; c >= 16; --c)
        if (D % c == 0) return c;
    return D;
}

// d-loop shape inside a chunk: a straight-line prefix of PRE steps then a rolled loop whose
// body is U steps, so the hot loop stays inside the instruction cache (a fully unrolled
// D=30 body starves issue with instruction fetch).
__host__ __device__ constexpr int pick_unroll(int DC) {
    if (DC <= 9) return DC;  // whole chunk straight-line
    int best_u = 8, best_rem = (DC - 1) % 8;
    for (int u = 9; u >= 4; --u) {
        const int rem = (DC - 1) % u;
        if (rem < best_rem) {
            best_rem = rem;
            best_u = u;
        }
    }
    return best_u;
}

// Operands of one d-step: 8 query values (broadcast operands) and 8 candidate values
// (4 packed pairs).
struct Frag {
    float4 qa, qb;  // q0..q3, q4..q7
    float4 c0, c1;  // candidates 0..3 | 4..7
};

template <int FORM, bool FIRST>
__device__ __forceinline__ void tile_step(float2 (&acc)[8][4], const Frag& f, const float2 (&cn)[4]) {
    const float q[8] = {f.qa.x, f.qa.y, f.qa.z, f.qa.w, f.qb.x, f.qb.y, f.qb.z, f.qb.w};
    const float2 c[4] = {make_float2(f.c0.x, f.c0.y), make_float2(f.c0.z, f.c0.w), make_float2(f.c1.x, f.c1.y),
                         make_float2(f.c1.z, f.c1.w)};
    const float2 m1 = make_float2(-1.0f, -1.0f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 qq = make_float2(q[i], q[i]);  // folds into FFMA2's scalar-broadcast operand
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            if (FORM == GMD_FORM_EXPANDED) {
                acc[i][p] = __ffma2_rn(qq, c[p], FIRST ? cn[p] : acc[i][p]);
            } else {
                const float2 df = __ffma2_rn(c[p], m1, qq);  // q - c
                if (FORM == GMD_FORM_DIRECT_L2) {
                    acc[i][p] = FIRST ? __fmul2_rn(df, df) : __ffma2_rn(df, df, acc[i][p]);
                } else {
                    acc[i][p].x = FIRST ? fabsf(df.x) : acc[i][p].x + fabsf(df.x);
                    acc[i][p].y = FIRST ? fabsf(df.y) : acc[i][p].y + fabsf(df.y);
                }
            }
        }
    }
}

template <int D, int NW, int NST>
struct ScanShape {
    static constexpr int BQ = NW * 32;
    static constexpr int DC = dim_chunk(D);             // dimensions per stage
    static constexpr int NCH = D / DC;                  // stages per tile pair
    static constexpr int TILE_F = (D + 1) * GMD_TW;     // floats of one HBM tile
    static constexpr int STAGE_F = (DC + 1) * GMD_TW;   // floats of one stage (chunk rows + norm row)
    static size_t smem_bytes(int kk) {
        return (size_t)(NST * STAGE_F + D * BQ + 3 * BQ) * 4 + 2 * NST * 8 + (size_t)BQ * QCAP * 8 +
               (size_t)BQ * (kk | 1) * 8;
    }
};

template <int D, int FORM, int NW, int NST>
__global__ void __launch_bounds__((NW + 1) * 32, 1) scan_topk_kernel(const ScanParams p) {
    using S = ScanShape<D, NW, NST>;
    constexpr int BQ = S::BQ, DC = S::DC, NCH = S::NCH, TILE_F = S::TILE_F, STAGE_F = S::STAGE_F;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr int U = pick_unroll(DC);
    constexpr int PRE = (DC <= 9) ? DC : 1 + (DC - 1) % U;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* Cs = reinterpret_cast<float*>(smem_raw);                // NST x [DC+1][64]
    float* Qs = Cs + NST * STAGE_F;                                // [D][BQ]
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
        // ---- producer: keep the ring full.  Stage `it` = (tile it / NCH, dimension chunk it % NCH).
        if (lane == 0) {
            const int nit = (te - tb) * NCH;
            for (int it = 0; it < nit; ++it) {
                const int s = it % NST;
                if (it >= NST) mbar_wait_backoff(&bars[NST + s], ((it / NST) - 1) & 1, 20 * DC);
                const int t = tb + it / NCH, ch = it % NCH;
                const float* ta = p.ctiles + (size_t)t * TILE_F;
                float* dst = Cs + s * STAGE_F;
                if (NCH == 1) {
                    mbar_expect_tx(&bars[s], STAGE_F * 4);  // whole tile, norm row included
                    bulk_g2s(dst, ta, STAGE_F * 4, &bars[s]);
                } else {
                    const uint32_t cb = DC * GMD_TW * 4, nb = GMD_TW * 4;
                    mbar_expect_tx(&bars[s], cb + (ch == 0 ? nb : 0));
                    bulk_g2s(dst, ta + ch * DC * GMD_TW, cb, &bars[s]);
                    if (ch == 0) bulk_g2s(dst + DC * GMD_TW, ta + D * GMD_TW, nb, &bars[s]);
                }
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

    int it = 0;
    for (int t = tb; t < te; ++t) {
        float2 acc2[8][4];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch, ++it) {
            const int s = it % NST;
            mbar_wait(&bars[s], (it / NST) & 1);
            const float* Ct = Cs + s * STAGE_F + lc * 4;
            const float* Qc = Qw + ch * DC * BQ;
            auto load = [&](int d) {
                Frag f;
                f.qa = *reinterpret_cast<const float4*>(Qc + d * BQ);
                f.qb = *reinterpret_cast<const float4*>(Qc + d * BQ + 16);
                f.c0 = *reinterpret_cast<const float4*>(Ct + d * GMD_TW);
                f.c1 = *reinterpret_cast<const float4*>(Ct + d * GMD_TW + 32);
                return f;
            };
            float2 cn[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) cn[x] = make_float2(0.f, 0.f);
            if (FORM == GMD_FORM_EXPANDED && ch == 0) {
                const Frag n = load(DC);  // row DC of the stage holds ||c||^2
                cn[0] = make_float2(n.c0.x, n.c0.y), cn[1] = make_float2(n.c0.z, n.c0.w);
                cn[2] = make_float2(n.c1.x, n.c1.y), cn[3] = make_float2(n.c1.z, n.c1.w);
            }
            // one step ahead in registers; the read one row past the chunk at its end is
            // harmless (norm row of the stage / next rows of Qs, qn)
            Frag cur = load(0);
#pragma unroll
            for (int d = 0; d < PRE; ++d) {
                const Frag nxt = load(d + 1);
                if (d == 0 && ch == 0)
                    tile_step<FORM, true>(acc2, cur, cn);
                else
                    tile_step<FORM, false>(acc2, cur, cn);
                cur = nxt;
            }
            if (PRE < DC) {
#pragma unroll 1
                for (int d0 = PRE; d0 < DC; d0 += U) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const Frag nxt = load(d0 + u + 1);
                        tile_step<FORM, false>(acc2, cur, cn);
                        cur = nxt;
                    }
                }
            }
            // this stage's shared memory is no longer needed by this warp
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[NST + s]);
        }

        // ---- selection fast path: one min-tree per query row against its threshold
        unsigned rf = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float m = fminf(acc2[i][0].x, acc2[i][0].y);
#pragma unroll
            for (int x = 1; x < 4; ++x) m = fminf(m, fminf(acc2[i][x].x, acc2[i][x].y));
            rf |= (m < thr[i]) ? (1u << i) : 0u;
        }
        if (__any_sync(FULL, rf != 0)) {
            // ---- hit path.  A lane with a hit only APPENDS (value, index) to the owning
            // query's queue (one shared-memory atomic + one store); thresholds stay as they
            // are.  Only when some queue is filling up (>= QHIGH) or overflowed does the warp
            // drain: lane L folds query L's queue into its heap and refreshes its threshold.
            // A stale threshold is an upper bound of the true one, so nothing is lost; the
            // heap orders by the full (value, index) key, so the result does not depend on
            // the order in which hits were queued.
            const uint32_t cbase = (uint32_t)t * GMD_TW + (lc << 2);
            unsigned pend[8];
            bool flag = false;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                pend[i] = 0;
                if (rf & (1u << i)) {
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        pend[i] |= (acc2[i][x].x < thr[i]) ? (1u << (2 * x)) : 0u;
                        pend[i] |= (acc2[i][x].y < thr[i]) ? (2u << (2 * x)) : 0u;
                    }
                    const int ql = ((i >> 2) << 4) + (lq << 2) + (i & 3);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (pend[i] & (1u << j)) {
                            const int slot = atomicAdd(&wcnt[ql], 1);
                            if (slot < QCAP) {
                                // candidate j of this lane: group j>>2 (32 rows apart), element j&3
                                const float v = (j & 1) ? acc2[i][j >> 1].y : acc2[i][j >> 1].x;
                                wqueue[ql * QCAP + slot] = gmd_key(v, cbase + ((j >> 2) << 5) + (j & 3));
                                pend[i] &= ~(1u << j);
                                flag |= slot >= QHIGH - 1;
                            } else {
                                flag = true;
                            }
                        }
                    }
                }
            }
            if (__any_sync(FULL, flag)) {
                bool more;
                do {
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
                    // hits that found their queue full: retry those that still beat the
                    // refreshed threshold (only during heap warm-up in practice)
                    bool left = false;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (pend[i]) {
                            const int ql = ((i >> 2) << 4) + (lq << 2) + (i & 3);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                if (pend[i] & (1u << j)) {
                                    const float v = (j & 1) ? acc2[i][j >> 1].y : acc2[i][j >> 1].x;
                                    if (!(v < thr[i])) {
                                        pend[i] &= ~(1u << j);
                                    } else {
                                        const int slot = atomicAdd(&wcnt[ql], 1);
                                        if (slot < QCAP) {
                                            wqueue[ql * QCAP + slot] = gmd_key(v, cbase + ((j >> 2) << 5) + (j & 3));
                                            pend[i] &= ~(1u << j);
                                        } else {
                                            left = true;
                                        }
                                    }
                                }
                            }
                        }
                    }
                    more = __any_sync(FULL, left);
                } while (more);
            }
        }
    }

    // ---- fold what is still queued into the heaps
    __syncwarp();
    {
        const int n = min(wcnt[lane], QCAP);
        uint64_t* h = wheap + (size_t)lane * kkp;
        for (int r = 0; r < n; ++r) {
            const uint64_t key = wqueue[lane * QCAP + r];
            if (key < h[0]) heap_replace_root(h, kk, key);
        }
        const uint64_t root = h[0];
        wthr[lane] = (root == GMD_EMPTY_KEY) ? __int_as_float(0x7f800000) : gmd_unflip((uint32_t)(root >> 32));
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
static int scan_launch(const ScanParams& p, cudaStream_t st) {
    using S = ScanShape<D, NW, NST>;
    const size_t smem = S::smem_bytes(p.kk);
    GMD_REQUIRE(smem <= GMD_SMEM_LIMIT, GMD_ERR_UNSUPPORTED,
                "scan: D=%d kk=%d needs %zu B of shared memory (> 227 KB)", D, p.kk, smem);
    auto kern = scan_topk_kernel<D, FORM, NW, NST>;
    GMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((p.Nq + S::BQ - 1) / S::BQ), (unsigned)p.nsplit);
    kern<<<grid, (NW + 1) * 32, smem, st>>>(p);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

// 8 compute warps (256 queries) per CTA when the heaps fit in shared memory, otherwise 4
template <int D, int FORM>
static int scan_dispatch(const ScanParams& p, cudaStream_t st) {
    if constexpr (D <= 64) {
        if (ScanShape<D, 8, 4>::smem_bytes(p.kk) <= GMD_SMEM_LIMIT) return scan_launch<D, FORM, 8, 4>(p, st);
        return scan_launch<D, FORM, 4, 4>(p, st);
    } else {
        return scan_launch<D, FORM, 4, 3>(p, st);
    }
}

}  // namespace gmd_scan

// padded dimensions the scan is instantiated for; gmd_scan_pad_dim(D) is the smallest >= D
#define GMD_SCAN_DIMS(X) \
    X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(20) X(24) X(28) X(30) X(32) X(40) X(48) X(64) X(96) X(136)

int gmd_scan_pad_dim(int D);
int gmd_scan_launch_form0(int DP, const gmd_scan::ScanParams& p, cudaStream_t st);
int gmd_scan_launch_form1(int DP, const gmd_scan::ScanParams& p, cudaStream_t st);
int gmd_scan_launch_form2(int DP, const gmd_scan::ScanParams& p, cudaStream_t st);
int gmd_scan_block_queries(int DP);
