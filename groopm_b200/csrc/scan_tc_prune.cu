// Pruned tensor-core scan: the tcgen05 distance + top-kk pipeline of scan_tc.cu, visiting only the
// candidate tiles that can still hold a neighbour.
//
// The profiles were staged in the GROUPED order of order.cu: every tile holds rows of one leader's
// group (groups padded to whole tiles, rows of similar distance to the leader together), and prep
// recorded a ball per tile (centroid, largest member distance) and per group (leader, largest member
// distance).  For a query tile Q with ball (m_q, r_q) and any ball (m, r) of candidates,
//     d(q, c) >= |m_q - m| - r_q - r                      for every q in Q, c in the ball,
// so once every query of the CTA holds a running kk-th best distance <= T, a ball whose bound exceeds
// T cannot change any of the CTA's top-kk sets and is never loaded.
//
// A CTA (one query tile) works in rounds:
//   round 0   the window: up to WMAX tiles of the candidate group the query tile itself belongs to
//             (its own neighbourhood) -- after it the thresholds are near-final;
//   bound     T = largest threshold distance over the CTA's real queries; all threads test the group
//             balls, the warps then test the tiles of the surviving groups: bitmask `todo`;
//   rounds    lists of <= LCAP tiles taken from `todo` go through the pipeline until it is empty.
// Producer, MMA issuers and selection threads read the same list in shared memory, so they agree on
// the sequence of tiles without any per-stage metadata; mbarrier phases simply continue across rounds
// (the ring and accumulator stages are numbered by a running tile counter).
//
// Selection is what bounds this kernel once few tiles are visited: a query's own neighbourhood comes first,
// so most of what it is shown enters its top-kk set (~kk (1 + ln(n / kk)) heap updates per query, each a chain
// of dependent shared-memory round trips and compares), and four selection warps -- one per scheduler -- hide
// none of that latency.  For one K block (D <= 31) a CTA therefore takes TWO query tiles (QT = 2): 256 queries,
// eight selection warps (two per scheduler), both tiles multiplied against every candidate stage (the ring
// is read once for two tiles), 2 accumulator stages x 2 query tiles x 128 columns of tensor memory, and an
// accumulator row read in two halves of 64 columns so that the selection threads stay within the registers
// 352 threads leave each other.  The two tiles are neighbours in the grouped order, so they share the window
// and most surviving tiles; a candidate tile is visited if EITHER query tile's bound admits it.
//
// Exactness.  cand_tau stays what refine.cu expects: every candidate that was LOADED but not proposed
// has an FP32 value >= tau.  A candidate that was never loaded has a true float64 distance
//     >= (rigorous bound, all roundings taken against it) > T >= sqrt(tau + |q|^2),
// which is above the certificate's lower(tau), so the certificate's conclusion -- "nothing outside the
// proposals can beat the k-th re-ranked distance" -- holds for it as well.  T only ever shrinks, and
// the test uses the same FP32 expression (fl(thr + |q|^2)) that the write-out stores in cand_tau.
// With data that does not cluster every ball survives and the scan degenerates into the full sweep.
#include <stdlib.h>

#include <algorithm>

#include "tc_common.cuh"

namespace gmd_tc {

constexpr int LCAP = 1024;                     // tiles per round
// Parked hits per selection thread (template parameter TQP): 16 where the shared memory allows, else 12.  Deeper
// queues even out the lanes of a warp (a drain lasts as long as its fullest queue: 9 slots 7.7 ms, 12 5.7 ms,
// 16 5.0 ms for the 1M x 30 scan), but a threshold is only refreshed when its queue is folded and a stale one
// parks more: queues of 48 in an L2-resident scratch took 7.9 ms.
constexpr int WMAX = 32;                       // tiles of the opening window
constexpr int META_MAX = GMD_MAX_D + 4;        // centroid + radius of the query tile in shared memory
constexpr float LB_REL = 1.0f / 16384.0f;      // relative slack on every FP32-evaluated term of a bound
constexpr float LB_ABS = 1.0f / 512.0f;        // absolute slack (scaled units: row norms <= 256): float64 -> FP32
                                               // rounding of the centred rows (<= 2^-15) and FP32 evaluation
// warp roles: 4 QT selection warps (warp w: query tile w / 4, TMEM lane quarter w % 4), then the single-thread
// roles on the highest warp ids: producer, two MMA issuers (even / odd candidate tiles)
__host__ __device__ constexpr int prune_threads(int qt) { return 32 * (4 * qt + 3); }

struct PruneParams {
    const unsigned char* qtiles;
    const unsigned char* ctiles;
    long long Nq_pos, Nc_pos;                  // staged positions (rows + padding) of the two sets
    int ntiles_c;                              // candidate tiles staged
    int kk, ns, D;
    const uint32_t* perm_q;                    // position -> row, GMD_EMPTY_IDX = padding
    const uint32_t* perm_c;
    const int32_t* tgroup_q;                   // [query tiles] group of the tile, -1 = padding only
    const int32_t* gstart_c;                   // [ngroups + 1] first candidate tile of a group
    int ngroups, ng_alloc;
    const float* gmeta;                        // [D + 1][ng_alloc] leader (scaled units), group radius
    const float* tmeta_c;                      // [D + 1][nt_alloc_c] tile centroid, radius (< 0: empty tile)
    int nt_alloc_c;
    const float* tmeta_q;                      // same for the query tiles
    int nt_alloc_q;
    int same;                                  // queries and candidates share one order
    long long c_index_offset;
    const float* stats_c;
    const float* stats_q;
    int ntiles_q;                              // query tiles staged
    uint32_t* cand_idx;                        // [Nq][kk]
    float* cand_val;
    float* cand_tau;                           // [Nq]
    unsigned long long* counters;              // nullable: [0] tiles visited, [1] CTAs, [2] groups surviving
};

__device__ __forceinline__ float ball_bound(float d2, float rsum) {
    return sqrtf(d2) * (1.0f - LB_REL) - rsum * (1.0f + LB_REL) - LB_ABS;
}

template <int NKB, int QT, int TQP>
__global__ void __launch_bounds__(prune_threads(QT), 1) scan_tc_prune_kernel(const PruneParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int PTHREADS = prune_threads(QT);
    constexpr int PW_PROD = 4 * QT, PW_ISSUE = 4 * QT + 1;
    constexpr int NA = QT == 2 ? 2 : NACC;                 // accumulator stages (QT x 128 columns each)
    const long long qt0 = (long long)blockIdx.x * QT;      // first query tile of the CTA
    // groups of the CTA's query tiles (-1: a tile of padding only, or past the end)
    int grpq[QT];
#pragma unroll
    for (int j = 0; j < QT; ++j) grpq[j] = (qt0 + j < p.ntiles_q) ? p.tgroup_q[qt0 + j] : -1;
    bool any_tile = false;
#pragma unroll
    for (int j = 0; j < QT; ++j) any_tile |= grpq[j] >= 0;
    if (!any_tile) return;                                 // nothing but padding (uniform: before any set-up)

    // (an offset from the array, not an integer round trip: the compiler must keep seeing SHARED memory here, or
    // every access below becomes a generic LD.E / ST.E)
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int nkb = NKB;
    constexpr int TILE_B = tile_bytes(nkb);
    const int ns = p.ns, kk = p.kk, KQ = heap8_slots(kk), D = p.D;
    const int nw = (p.ntiles_c + 31) >> 5;
    unsigned char* Aimg = base;                            // QT query tiles
    unsigned char* Bst = base + QT * TILE_B;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Bst + (size_t)ns * KB_B);
    uint64_t* full = bars;
    uint64_t* sfree = bars + NS_MAX;
    uint64_t* tdone = bars + 2 * NS_MAX;
    uint64_t* tempty = bars + 2 * NS_MAX + NACC;
    uint64_t* qbar = bars + 2 * NS_MAX + 2 * NACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS_MAX + 2 * NACC + 2);
    uint2* queues = reinterpret_cast<uint2*>(bars + 2 * NS_MAX + 2 * NACC + 4);   // [QT][TQP][TM]
    float* hvals = reinterpret_cast<float*>(queues + QT * TQP * TM);                // [QT][TM][KQ] values of the top-kk sets
    uint32_t* hidx = reinterpret_cast<uint32_t*>(hvals + (size_t)QT * TM * KQ);    // [QT][TM][KQ] their candidates
    uint32_t* tlist = hidx + (size_t)QT * TM * KQ;                                 // [LCAP] tiles of the round / groups
    float* mq = reinterpret_cast<float*>(tlist + LCAP);                        // [QT][META_MAX] centroid, radius
    float* Tw = mq + QT * META_MAX;                                            // [4 QT] per selection warp
    int* ctl = reinterpret_cast<int*>(Tw + 4 * QT);                            // [2][2] (count, pending) by round parity
    uint32_t* visited = reinterpret_cast<uint32_t*>(ctl + 4);                  // [nw]
    uint32_t* todo = visited + nw;                                             // [nw]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr unsigned FULLW = 0xFFFFFFFFu;
    const float INF = __int_as_float(0x7f800000);
#ifdef GMD_TC_PROBES
    // developer build only: SM cycles of thread 0 (a selection thread) per phase, summed over the CTAs in
    // counters[4..]: wait for an accumulator, read it, trees + vote, parking, drains; drain calls; entries folded
    long long pk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PROBE_T(var) const long long var = clock64()
#define PROBE_ADD(i, v) pk[i] += (v)
#else
#define PROBE_T(var)
#define PROBE_ADD(i, v)
#endif

    if (tid == 0) {
        for (int s = 0; s < ns; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&sfree[s], 1);
        }
        for (int a = 0; a < NA; ++a) {
            mbar_init(&tdone[a], 1);
            mbar_init(&tempty[a], 4 * QT);                 // one arrival per selection warp
        }
        mbar_init(qbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(qbar, QT * TILE_B);
        // (a tile past the end of the query set is a second copy of the first one: its rows count as padding)
        for (int j = 0; j < QT; ++j)
            bulk_g2s(Aimg + (size_t)j * TILE_B, p.qtiles + (size_t)(qt0 + j < p.ntiles_q ? qt0 + j : qt0) * TILE_B, TILE_B, qbar);
    }
    for (int e = tid; e < QT * TM * KQ; e += PTHREADS) hvals[e] = -INF, hidx[e] = GMD_EMPTY_IDX;
    for (int e = tid; e < 2 * nw; e += PTHREADS) visited[e] = 0u;               // visited and todo are adjacent
    for (int e = tid; e < QT * (D + 1); e += PTHREADS) {
        const int j = e / (D + 1), d = e - j * (D + 1);
        mq[j * META_MAX + d] = grpq[j] >= 0 ? p.tmeta_q[(size_t)d * p.nt_alloc_q + qt0 + j] : 0.0f;
    }
    if (tid < 4 * QT) Tw[tid] = INF;                       // no threshold yet: nothing can be pruned
    if (warp == PW_ISSUE) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;

    // ---- round 0: the window, up to WMAX tiles of each query tile's own candidate group
    if (tid == 0) {
        int n = 0;
        for (int j = 0; j < QT; ++j) {
            if (grpq[j] < 0) continue;
            const int gs = p.gstart_c[grpq[j]], ge = p.gstart_c[grpq[j] + 1];
            int ws = gs, we = ge;
            if (ge - gs > WMAX) {
                if (p.same) ws = min(max((int)(qt0 + j) - WMAX / 2, gs), ge - WMAX);
                we = ws + WMAX;
            }
            for (int t = ws; t < we; ++t) {
                if ((visited[t >> 5] >> (t & 31)) & 1u) continue;     // the two tiles usually share their group
                tlist[n++] = (uint32_t)t;
                visited[t >> 5] |= 1u << (t & 31);
            }
        }
        ctl[0] = n;
        ctl[1] = 0;
    }

    // ---- persistent state of the roles
    // producer
    long long p_it = 0;
    int p_s = 0, p_pkb = 0;
    long long p_pt = 0;
    uint32_t p_sph = 0;
    // issuers
    const int nissue = ns > nkb ? 2 : 1;
    const int first = warp - PW_ISSUE;
    int i_s = 0;
    uint32_t i_ph = 0;
    // selection: thread <-> query row `row` (= TMEM lane) of query tile `qsel`
    const int quarter = warp & 3, qsel = (warp >> 2) % QT;
    const int row = quarter * 32 + lane;                   // only meaningful for the selection warps
    float* hv = hvals + ((size_t)qsel * TM + row) * KQ;
    uint32_t* hx = hidx + ((size_t)qsel * TM + row) * KQ;
    uint2* myq = queues + (size_t)qsel * TQP * TM + row;   // slot r at myq[r * TM]: conflict-free
    int qcnt = 0, hn = 0;                                  // parked hits; entries in the top-kk set
    float thr = INF;                                       // threshold (s units)
    float uthr = -INF;                                     // the same in u = -s / 2 units
    const long long qpos = (qt0 + qsel) * TM + row;
    bool qreal = false;
    float qnv = 0.0f;
    if (warp < PW_PROD)
        qreal = grpq[qsel] >= 0 && qpos < p.Nq_pos && (p.perm_q ? p.perm_q[qpos] != GMD_EMPTY_IDX : true);

    long long tbase = 0;                                   // tiles processed before this round (ring / stage numbering)
    long long visited_tiles = 0;
    int round = 0;
    bool bounded = false;                                  // the bound phase has run
    while (true) {
        __syncthreads();
        int cnt = ctl[(round & 1) * 2];
        const int pending = ctl[(round & 1) * 2 + 1];
        if (cnt == 0 && pending == 0) {
            if (bounded) break;
            // ---- bound phase: which unvisited tiles can still matter
            bounded = true;
            int* nsurv = ctl + ((round + 1) & 1) * 2;      // scratch: this slot is rewritten by the list builder below
            if (tid == 0) *nsurv = 0;
            __syncthreads();
            // T: per query tile, the largest threshold distance of a real query (scaled units, published by the
            // selection warps at the end of their last round; -1: no such tile, nothing survives for it)
            float T[QT], rq[QT];
#pragma unroll
            for (int j = 0; j < QT; ++j) {
                T[j] = grpq[j] >= 0 ? fmaxf(fmaxf(Tw[4 * j], Tw[4 * j + 1]), fmaxf(Tw[4 * j + 2], Tw[4 * j + 3])) : -1.0f;
                rq[j] = mq[j * META_MAX + D];
            }
            // does the ball (centre column `col` of a [D + 1][pitch] array, radius r) survive for a query tile?
            auto survives = [&](const float* meta, size_t pitch, size_t col, float r) {
                float d2[QT];
#pragma unroll
                for (int j = 0; j < QT; ++j) d2[j] = 0.0f;
#pragma unroll 6
                for (int d = 0; d < D; ++d) {
                    const float c = meta[(size_t)d * pitch + col];
#pragma unroll
                    for (int j = 0; j < QT; ++j) {
                        const float t = mq[j * META_MAX + d] - c;
                        d2[j] = fmaf(t, t, d2[j]);
                    }
                }
                bool keep = false;
#pragma unroll
                for (int j = 0; j < QT; ++j) keep |= T[j] >= 0.0f && !(ball_bound(d2[j], rq[j] + r) > T[j]);
                return keep;
            };
            // groups: every thread tests its share; the survivors go to tlist (if they fit)
            for (int g = tid; g < p.ngroups; g += PTHREADS) {
                if (survives(p.gmeta, (size_t)p.ng_alloc, (size_t)g, p.gmeta[(size_t)D * p.ng_alloc + g])) {
                    const int slot = atomicAdd(nsurv, 1);
                    if (slot < LCAP) tlist[slot] = (uint32_t)g;
                }
            }
            __syncthreads();
            const int ng = *nsurv;
            const bool all = ng > LCAP;                   // list overflow: walk every group
            const int ngl = all ? p.ngroups : ng;
            // tiles of the surviving groups: a warp per group, a lane per tile
            for (int gi = warp; gi < ngl; gi += PTHREADS / 32) {
                const int g = all ? gi : (int)tlist[gi];
                const int te = p.gstart_c[g + 1];
                for (int t = p.gstart_c[g] + lane; t < te; t += 32) {
                    if ((visited[t >> 5] >> (t & 31)) & 1u) continue;
                    const float rt = p.tmeta_c[(size_t)D * p.nt_alloc_c + t];
                    if (rt < 0.0f) continue;               // a tile of padding only
                    if (survives(p.tmeta_c, (size_t)p.nt_alloc_c, (size_t)t, rt)) atomicOr(&todo[t >> 5], 1u << (t & 31));
                }
            }
            if (tid == 0 && p.counters) atomicAdd(&p.counters[2], (unsigned long long)ng);
            __syncthreads();
        }
        if (cnt == 0) {
            // ---- take the next <= LCAP tiles out of todo (warp 0); everybody meets at the top of the loop
            if (warp == 0) {
                int n = 0;
                bool left = false;
                for (int w0 = 0; w0 < nw; w0 += 32) {
                    const int w = w0 + lane;
                    uint32_t bits = w < nw ? todo[w] : 0u;
                    const uint32_t orig = bits;
                    const int c = __popc(bits);
                    int pre = c;
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(FULLW, pre, o);
                        if (lane >= o) pre += v;
                    }
                    const int total = __shfl_sync(FULLW, pre, 31);
                    int pos = n + pre - c;
                    uint32_t keep = 0u;
                    while (bits) {
                        const int b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        if (pos < LCAP)
                            tlist[pos] = (uint32_t)(w * 32 + b);
                        else
                            keep |= 1u << b;
                        ++pos;
                    }
                    if (w < nw && orig) {
                        todo[w] = keep;
                        visited[w] |= orig & ~keep;
                    }
                    left |= __any_sync(FULLW, keep != 0u);
                    n = min(LCAP, n + total);
                    if (n >= LCAP && w0 + 32 < nw) {       // the list is full: whatever is left waits for the next round
                        left = true;
                        break;
                    }
                }
                if (lane == 0) {
                    ctl[((round + 1) & 1) * 2] = n;
                    ctl[((round + 1) & 1) * 2 + 1] = (left && n > 0) ? 1 : 0;
                }
            }
            ++round;
            continue;
        }

        // ---- the round's cnt tiles through the pipeline
        if (warp == PW_PROD) {
            if (lane == 0) {
                for (int i = 0; i < cnt; ++i) {
                    const unsigned char* src = p.ctiles + (size_t)tlist[i] * TILE_B;
#pragma unroll 1
                    for (int kb = 0; kb < nkb; ++kb, ++p_it) {
                        if (p_it >= ns) {
                            // a stage that held the last K block of a tile is free when the tile's accumulator
                            // is complete (tdone), any other when the MMAs of its K block are (sfree)
                            if (p_pkb == nkb - 1) {
                                mbar_wait_backoff(&tdone[p_pt % NA], (uint32_t)(p_pt / NA) & 1u, 64);
                            } else {
                                mbar_wait_backoff(&sfree[p_s], (p_sph >> p_s) & 1u, 64);
                                p_sph ^= 1u << p_s;
                            }
                            if (++p_pkb == nkb) p_pkb = 0, ++p_pt;
                        }
                        mbar_expect_tx(&full[p_s], (uint32_t)KB_B);
                        bulk_g2s(Bst + (size_t)p_s * KB_B, src + (size_t)kb * KB_B, (uint32_t)KB_B, &full[p_s]);
                        if (++p_s == ns) p_s = 0;
                    }
                }
            }
            __syncwarp();
        } else if (warp >= PW_ISSUE) {
            if (lane == 0 && first < nissue) {
                if (tbase == 0) mbar_wait(qbar, 0);
                const uint32_t idesc = (1u << 4) | ((uint32_t)(TM >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
                const uint64_t a_base = umma_desc(smem_u32(Aimg));
                const uint64_t b_base = umma_desc(smem_u32(Bst));
                for (long long t = tbase; t < tbase + cnt; ++t) {
                    if (nissue == 2 && (int)(t & 1) != first) {    // the other issuer's tile
                        i_s += nkb;
                        while (i_s >= ns) i_s -= ns, i_ph ^= 1u;
                        continue;
                    }
                    const int a = (int)(t % NA);
                    if (t >= NA) mbar_wait(&tempty[a], (uint32_t)((t / NA) - 1) & 1u);
#pragma unroll
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&full[i_s], i_ph);
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        const uint64_t b_hi = b_base + (uint64_t)i_s * (KB_B >> 4), b_lo = b_hi + (IMG_B >> 4);
                        // every query tile of the CTA against this candidate stage: hi.hi + hi.lo + lo.hi
#pragma unroll
                        for (int j = 0; j < QT; ++j) {
                            const uint32_t acc = tmem + (uint32_t)((a * QT + j) * TM);
                            const uint64_t a_hi = a_base + (uint64_t)(j * (TILE_B >> 4)) + (uint64_t)kb * (KB_B >> 4),
                                           a_lo = a_hi + (IMG_B >> 4);
#pragma unroll
                            for (int q = 0; q < 3; ++q)
#pragma unroll
                                for (int k = 0; k < TK / 16; ++k)
                                    umma_f16(acc, (q == 2 ? a_lo : a_hi) + 2 * k, (q == 1 ? b_lo : b_hi) + 2 * k, idesc,
                                             (kb | q | k) ? 1u : 0u);
                        }
                        umma_commit(kb + 1 < nkb ? &sfree[i_s] : &tdone[a]);
                        if (++i_s == ns) i_s = 0, i_ph ^= 1u;
                    }
                }
            }
            __syncwarp();
        } else {
            // ---- selection warps: thread <-> query row <-> TMEM lane
            if (tbase == 0) {
                mbar_wait(qbar, 0);
                qnv = reinterpret_cast<const float*>(Aimg + (size_t)qsel * TILE_B + (size_t)nkb * KB_B)[row];
            }
            auto drain = [&]() {
                PROBE_T(d0);
                PROBE_ADD(5, 1);
                PROBE_ADD(6, qcnt);
                PROBE_ADD(7, __reduce_max_sync(FULLW, (unsigned)qcnt));   // the drain lasts as long as its fullest queue
                const Drain2 ds = drain_fill8(smem_u32(hv), smem_u32(hx), kk, hn, thr, smem_u32(myq), TM * 8, qcnt);
                thr = ds.thr;
                hn = ds.hn;
                qcnt = 0;
                uthr = -0.5f * thr;
                PROBE_ADD(4, clock64() - d0);
            };
            // 64 columns at a time (one tcgen05.ld.x64): maxima of the 8 groups of 8 columns as shallow trees,
            // ONE vote, then 8 values at a time: park what beats the (possibly stale) threshold, drain all the
            // warp's queues side by side when one of them is nearly full
            auto half_row = [&](const uint32_t (&r0)[32], const uint32_t (&r1)[32], uint32_t c0) {
                float gm[8];
                auto tree = [&](const uint32_t (&r)[32], float* g4) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float m0 = fmaxf(fmaxf(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1])), __uint_as_float(r[8 * g + 2]));
                        const float m1 = fmaxf(fmaxf(__uint_as_float(r[8 * g + 3]), __uint_as_float(r[8 * g + 4])), __uint_as_float(r[8 * g + 5]));
                        const float m2 = fmaxf(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
                        g4[g] = fmaxf(fmaxf(m0, m1), m2);
                    }
                };
                tree(r0, gm);
                tree(r1, gm + 4);
                const float cm = fmaxf(fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3])),
                                       fmaxf(fmaxf(gm[4], gm[5]), fmaxf(gm[6], gm[7])));
                if (!__any_sync(FULLW, cm > uthr)) return;
                PROBE_T(p0);
                PROBE_ADD(3, -(long long)0);
#ifdef GMD_TC_PROBES
                const long long dr0 = pk[4];
#endif
                unsigned mine = 0;
#pragma unroll
                for (int g = 0; g < 8; ++g) mine |= (gm[g] > uthr) ? (1u << g) : 0u;
                const unsigned hot = __reduce_or_sync(FULLW, mine);
                auto park = [&](const uint32_t (&r)[32], unsigned hot4, uint32_t cc) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (hot4 & (1u << g)) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float u = __uint_as_float(r[8 * g + j]);
                                if (u > uthr) {
                                    myq[qcnt * TM] = make_uint2(cc + 8 * g + j, __float_as_uint(-2.0f * u));
                                    ++qcnt;
                                }
                            }
                            if (__any_sync(FULLW, qcnt > TQP - 8)) drain();
                        }
                    }
                };
                park(r0, hot & 15u, c0);
                park(r1, hot >> 4, c0 + 32);
#ifdef GMD_TC_PROBES
                pk[3] += (clock64() - p0) - (pk[4] - dr0);
#endif
            };
            for (long long t = tbase; t < tbase + cnt; ++t) {
                const int a = (int)(t % NA);
                PROBE_T(s0);
                mbar_wait(&tdone[a], (uint32_t)(t / NA) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;");
                PROBE_T(s1);
                PROBE_ADD(0, s1 - s0);
                const uint32_t cbase = tlist[t - tbase] * TM;
                const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((a * QT + qsel) * TM);
                uint32_t r0[32], r1[32];
                tmem_ld64_sync(trow, r0, r1);
                PROBE_T(s2);
                PROBE_ADD(1, s2 - s1);
                half_row(r0, r1, cbase);
                PROBE_T(s3);
                tmem_ld64_sync(trow + 64, r0, r1);
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[a]);
                PROBE_T(s4);
                PROBE_ADD(1, s4 - s3);
                half_row(r0, r1, cbase + 64);
                PROBE_ADD(2, (clock64() - s4) + (s3 - s2));       // trees + parking + drains of both halves
            }
            // end of the round: fold what is parked, publish the warp's largest threshold DISTANCE (scaled units;
            // padding rows do not count).  Non-negative floats (and +inf) order like their bit patterns.
            drain();
            float tq = 0.0f;
            if (qreal) tq = sqrtf(fmaxf(thr + qnv, 0.0f)) * (1.0f + 1.0f / 1048576.0f);
            const unsigned tmax = __reduce_max_sync(FULLW, __float_as_uint(tq));
            if (lane == 0) Tw[warp] = __uint_as_float(tmax);
        }
        visited_tiles += cnt;
        tbase += cnt;
        // the list of this round is dead once every role is through: the slot of the NEXT round says "build"
        if (tid == 0) {
            ctl[((round + 1) & 1) * 2] = 0;
            ctl[((round + 1) & 1) * 2 + 1] = pending;
        }
        ++round;
    }

    // ---- write this query's proposals
    if (warp < PW_PROD && qreal) {
        const long long q = p.perm_q ? (long long)p.perm_q[qpos] : qpos;
        const float unscale = ldexpf(1.0f, -2 * gmd_tc_exponent(p.stats_c[0], p.stats_q[0]));
        const size_t ob = (size_t)q * kk;
        // eight entries at a time: the position -> row look-ups of a batch are in flight together
        for (int x0 = 0; x0 < kk; x0 += 8) {
            uint32_t oi[8];
            float vv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int x = x0 + j;
                const uint32_t li = x < hn ? hx[x + 7] : GMD_EMPTY_IDX;
                vv[j] = x < hn ? hv[x + 7] : INF;
                oi[j] = GMD_EMPTY_IDX;
                if (li != GMD_EMPTY_IDX && (long long)li < p.Nc_pos) oi[j] = p.perm_c ? p.perm_c[li] : li;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int x = x0 + j;
                if (x >= kk) break;
                const bool ok = oi[j] != GMD_EMPTY_IDX;
                p.cand_idx[ob + x] = ok ? oi[j] + (uint32_t)p.c_index_offset : GMD_EMPTY_IDX;
                p.cand_val[ob + x] = ok ? fmaxf(vv[j] + qnv, 0.0f) * unscale : INF;
            }
        }
        // a set that never filled lists every candidate it was shown: nothing is bounded by it
        p.cand_tau[q] = hn == kk ? fmaxf(thr + qnv, 0.0f) * unscale : INF;
    }
    if (tid == 0) {
        if (tbase == 0) mbar_wait(qbar, 0);                // nothing was scanned: the query tile's copy must still land
        if (p.counters) {
            atomicAdd(&p.counters[0], (unsigned long long)visited_tiles);
            atomicAdd(&p.counters[1], 1ull);
#ifdef GMD_TC_PROBES
            for (int x = 0; x < 8; ++x) atomicAdd(&p.counters[4 + x], (unsigned long long)pk[x]);
#endif
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == PW_ISSUE) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// leader rows and group radii in the scan's scaled units, dimension-major
__global__ void __launch_bounds__(256) group_meta_kernel(const float* __restrict__ L, const float* __restrict__ grad,
                                                        int ngroups, int D, const float* __restrict__ stats_c,
                                                        const float* __restrict__ stats_q, float* __restrict__ gmeta,
                                                        int ng_alloc) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (D + 1) * ngroups) return;
    const int d = e / ngroups, g = e - d * ngroups;
    const float scale = ldexpf(1.0f, gmd_tc_exponent(stats_c[0], stats_q[0]));
    // the radius is a float64 distance between ORIGINAL rows, rounded up; LB_ABS covers the centring
    gmeta[(size_t)d * ng_alloc + g] = d < D ? L[(size_t)g * D + d] * scale : grad[g] * scale * (1.0f + 1.0f / 1048576.0f);
}

}  // namespace gmd_tc

namespace {
// shared memory of one CTA (qt query tiles) without the ring and the two bitmasks
size_t prune_fixed_smem(int nkb, int kk, int qt, int tqp = 12) {
    using namespace gmd_tc;
    return 1024 + (size_t)qt * tile_bytes(nkb) + 512 + (size_t)qt * tqp * TM * 8 + (size_t)qt * TM * heap8_slots(kk) * 8 +
           (size_t)LCAP * 4 + (size_t)qt * META_MAX * 4 + 128;
}
}  // namespace

extern "C" size_t gmd_tc_prune_max_tiles(int D, int kk) {
    using namespace gmd_tc;
    const int nkb = tc_nkb(D);
    if (nkb > MAX_KB) return 0;
    const size_t need = prune_fixed_smem(nkb, kk, 1) + (size_t)std::max(2, nkb + 1) * KB_B;   // >= nkb + 1 ring stages
    if (need >= GMD_SMEM_LIMIT) return 0;
    return (size_t)(GMD_SMEM_LIMIT - need) / 8 * 32;                  // two bitmask words per 32 tiles
}

extern "C" int gmd_tc_group_meta(const float* L, const float* grad, int ngroups, int D, const float* stats_c,
                                 const float* stats_q, float* gmeta, int ng_alloc, void* stream) {
    GMD_REQUIRE(L && grad && stats_c && stats_q && gmeta && ngroups > 0 && ngroups <= ng_alloc && D >= 1,
                GMD_ERR_BAD_ARG, "tc_group_meta: bad arguments");
    const int n = (D + 1) * ngroups;
    gmd_tc::group_meta_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(L, grad, ngroups, D, stats_c, stats_q,
                                                                                gmeta, ng_alloc);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

extern "C" int gmd_scan_tc_prune_f32(const void* qtiles, int64_t Nq_pos, const void* ctiles, int64_t Nc_pos, int D, int kk,
                                     int64_t c_index_offset, const float* stats_c, const float* stats_q,
                                     const uint32_t* perm_q, const uint32_t* perm_c, const int32_t* tgroup_q,
                                     const int32_t* gstart_c, int ngroups, const float* gmeta, int ng_alloc,
                                     const float* tmeta_q, int nt_alloc_q, const float* tmeta_c, int nt_alloc_c, int same,
                                     uint32_t* cand_idx, float* cand_val, float* cand_tau, uint64_t* counters,
                                     void* stream) {
    GMD_REQUIRE(qtiles && ctiles && stats_c && stats_q && cand_idx && cand_val && cand_tau && tgroup_q && gstart_c &&
                    gmeta && tmeta_q && tmeta_c,
                GMD_ERR_BAD_ARG, "scan_tc_prune: null pointer");
    GMD_REQUIRE(Nq_pos > 0 && Nc_pos > 0 && Nc_pos + c_index_offset < 0xFFFFFFFFll && ngroups > 0, GMD_ERR_BAD_ARG,
                "scan_tc_prune: bad sizes");
    using namespace gmd_tc;
    const int nkb = tc_nkb(D);
    GMD_REQUIRE(D >= 1 && nkb <= MAX_KB && kk >= 1 && kk <= 96, GMD_ERR_UNSUPPORTED, "scan_tc_prune: D=%d kk=%d", D, kk);
    const int ntiles_c = (int)((Nc_pos + TM - 1) / TM);
    const int ntiles_q = (int)((Nq_pos + TM - 1) / TM);
    GMD_REQUIRE(ntiles_c <= nt_alloc_c && ntiles_q <= nt_alloc_q, GMD_ERR_BAD_ARG, "scan_tc_prune: tile meta too small");
    GMD_REQUIRE((size_t)ntiles_c <= gmd_tc_prune_max_tiles(D, kk), GMD_ERR_UNSUPPORTED,
                "scan_tc_prune: %d candidate tiles exceed the shared-memory bitmasks", ntiles_c);
    const int nw = (ntiles_c + 31) / 32;
    // two query tiles per CTA when there is one K block and both fit (GMD_TC_QT=1: developer switch)
    int qt = (nkb == 1 && prune_fixed_smem(nkb, kk, 2) + (size_t)nw * 8 + 2 * (size_t)KB_B <= GMD_SMEM_LIMIT) ? 2 : 1;
    if (const char* env = getenv("GMD_TC_QT")) qt = std::min(qt, std::max(1, atoi(env)));
    // queues of 16 where they fit beside at least nkb + 1 ring stages
    const int tqp = prune_fixed_smem(nkb, kk, qt, 16) + (size_t)nw * 8 + (size_t)std::max(2, nkb + 1) * KB_B <= GMD_SMEM_LIMIT ? 16 : 12;
    const size_t fixed = prune_fixed_smem(nkb, kk, qt, tqp) + (size_t)nw * 8;
    const int na = qt == 2 ? 2 : NACC;
    const int ns = (int)std::min<size_t>(std::min(NS_MAX, na * nkb), (GMD_SMEM_LIMIT - fixed) / KB_B);
    GMD_REQUIRE(ns >= 2, GMD_ERR_UNSUPPORTED, "scan_tc_prune: D=%d kk=%d needs more than 227 KB of shared memory", D, kk);
    const size_t smem = fixed + (size_t)ns * KB_B;
    PruneParams p;
    p.qtiles = (const unsigned char*)qtiles;
    p.ctiles = (const unsigned char*)ctiles;
    p.Nq_pos = Nq_pos;
    p.Nc_pos = Nc_pos;
    p.ntiles_c = ntiles_c;
    p.kk = kk;
    p.ns = ns;
    p.D = D;
    p.perm_q = perm_q;
    p.perm_c = perm_c;
    p.tgroup_q = tgroup_q;
    p.gstart_c = gstart_c;
    p.ngroups = ngroups;
    p.ng_alloc = ng_alloc;
    p.gmeta = gmeta;
    p.tmeta_c = tmeta_c;
    p.nt_alloc_c = nt_alloc_c;
    p.tmeta_q = tmeta_q;
    p.nt_alloc_q = nt_alloc_q;
    p.same = same;
    p.c_index_offset = c_index_offset;
    p.stats_c = stats_c;
    p.stats_q = stats_q;
    p.ntiles_q = ntiles_q;
    p.cand_idx = cand_idx;
    p.cand_val = cand_val;
    p.cand_tau = cand_tau;
    p.counters = (unsigned long long*)counters;
    auto launch = [&](auto kern, int q) -> int {
        GMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)((ntiles_q + q - 1) / q), prune_threads(q), smem, (cudaStream_t)stream>>>(p);
        return GMD_OK;
    };
    int rc;
    switch (nkb) {
        case 1:
            rc = qt == 2 ? (tqp == 16 ? launch(scan_tc_prune_kernel<1, 2, 16>, 2) : launch(scan_tc_prune_kernel<1, 2, 12>, 2))
                         : (tqp == 16 ? launch(scan_tc_prune_kernel<1, 1, 16>, 1) : launch(scan_tc_prune_kernel<1, 1, 12>, 1));
            break;
        case 2: rc = tqp == 16 ? launch(scan_tc_prune_kernel<2, 1, 16>, 1) : launch(scan_tc_prune_kernel<2, 1, 12>, 1); break;
        case 3: rc = tqp == 16 ? launch(scan_tc_prune_kernel<3, 1, 16>, 1) : launch(scan_tc_prune_kernel<3, 1, 12>, 1); break;
        case 4: rc = tqp == 16 ? launch(scan_tc_prune_kernel<4, 1, 16>, 1) : launch(scan_tc_prune_kernel<4, 1, 12>, 1); break;
        default: rc = tqp == 16 ? launch(scan_tc_prune_kernel<5, 1, 16>, 1) : launch(scan_tc_prune_kernel<5, 1, 12>, 1); break;
    }
    if (rc != GMD_OK) return rc;
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}
