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
// whose tile holds values under a threshold appends (index, value) to its OWN cell of the
// query's pending set in shared memory -- a cell per (lane, register-tile row), its fill
// count kept in a register nibble -- so a hit costs predicated stores only: no atomics, no
// cross-lane traffic.  The cells are folded into the heaps (lane L owns query L) only when
// one of them is full, so heap maintenance is amortised over many hits.
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
    uint32_t* best1 = nullptr;   // BMU scan only, optional: [Nq] the winning cell's best candidate (FP32 arithmetic)
};

// Per-query running top-kk set: a 4-ary MAX-heap of u64 keys in shared memory (root = current
// kk-th best = the threshold).  Four children per node: the sift-down of a replaced root is
// a chain of ceil(log4 kk) dependent shared-memory round trips instead of log2 kk (3 instead of
// 6 at kk = 40), each one loading its four children back to back.  The owning lane / thread
// maintains the heap by itself; the float64 re-rank orders the survivors anyway.  Empty slots
// hold GMD_EMPTY_KEY (largest key), so a freshly initialised array is already a valid heap.
__device__ __forceinline__ void heap_replace_root(uint64_t* h, int kk, uint64_t key) {
    int pos = 0;
    while (true) {
        const int c = 4 * pos + 1;
        if (c >= kk) break;
        const uint64_t k0 = h[c];
        const uint64_t k1 = (c + 1 < kk) ? h[c + 1] : 0ull;
        const uint64_t k2 = (c + 2 < kk) ? h[c + 2] : 0ull;
        const uint64_t k3 = (c + 3 < kk) ? h[c + 3] : 0ull;
        const bool b01 = k1 > k0, b23 = k3 > k2;
        const uint64_t m01 = b01 ? k1 : k0, m23 = b23 ? k3 : k2;
        const bool hi = m23 > m01;
        const uint64_t best = hi ? m23 : m01;
        if (best <= key) break;
        h[pos] = best;
        pos = c + (hi ? 2 + (int)b23 : (int)b01);
    }
    h[pos] = key;
}

// Two ways of parking hits until a heap update is worth it, chosen per dimension by
// measurement:
//   HP_CELLS  a cell per (lane, register-tile row), fill counts in a register nibble: no
//             atomics, all flagged rows handled in one sweep -- best when hits are dense
//             relative to the tile time (expanded form, D <= 16)
//   HP_QUEUE  one queue per query, slots reserved with a shared-memory atomic, one row per
//             lane and pass through a scratch row -- least work when hits are sparse
constexpr int HP_CELLS = 0, HP_QUEUE = 1;
constexpr int QL = 4;      // HP_CELLS: pending-hit slots per (lane, row) cell; 8 cells per query
constexpr int QCAP = 16;   // HP_QUEUE: pending-hit slots per query
constexpr int QHIGH = 8;   // HP_QUEUE: a queue this full asks for a drain of the warp's queues
__host__ __device__ constexpr int pick_hitpath(int D, int form) {
    return (form == GMD_FORM_EXPANDED && D <= 16) ? HP_CELLS : HP_QUEUE;
}

// ---- cold-path helper, deliberately out of line: the hot loop (distance tile + min-tree)
// must stay resident in the ~6 KB L0 / 32 KB L1.5 instruction caches.

// Lane L folds the 8 cells of query L into its heap and publishes the new threshold.
// cells: this warp's [8 rows][QL][32 lanes] array of (index, value bits); the cell of
// (row i, lane s) holds nib(counts of lane s, i) entries.  The query of lane L is row
// iL = ((L >> 4) << 2) + (L & 3) of the lanes 8*((L >> 2) & 3) + c, c = 0..7.
static __device__ __noinline__ void drain_cells(uint64_t* h, int kk, const uint2* cells, unsigned cnts, int lane,
                                                float* thr_out) {
    const int iL = ((lane >> 4) << 2) + (lane & 3);
    const int s0 = ((lane >> 2) & 3) << 3;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
        const int n = (int)((__shfl_sync(0xFFFFFFFFu, cnts, s0 + c) >> (4 * iL)) & 15u);
        for (int r = 0; r < n; ++r) {
            const uint2 e = cells[(iL * QL + r) * 32 + s0 + c];
            const uint64_t key = gmd_key(__uint_as_float(e.y), e.x);
            if (key < h[0]) heap_replace_root(h, kk, key);
        }
    }
    const uint64_t root = h[0];
    *thr_out = (root == GMD_EMPTY_KEY) ? __int_as_float(0x7f800000) : gmd_unflip((uint32_t)(root >> 32));
}

// HP_QUEUE: lane L folds query L's queue of raw (index, value bits) pairs into its heap.
static __device__ __noinline__ void drain_queue(uint64_t* h, int kk, const uint2* q, int* cnt, float* thr_out) {
    const int n = min(*cnt, QCAP);
    for (int r = 0; r < n; ++r) {
        const uint2 e = q[r];
        const uint64_t key = gmd_key(__uint_as_float(e.y), e.x);
        if (key < h[0]) heap_replace_root(h, kk, key);
    }
    const uint64_t root = h[0];
    *thr_out = (root == GMD_EMPTY_KEY) ? __int_as_float(0x7f800000) : gmd_unflip((uint32_t)(root >> 32));
    *cnt = 0;
}

// predicated 8-byte shared-memory store (kept as a predicated STS, never a branch)
__device__ __forceinline__ void sts64_if(unsigned ok, uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.u32 p, %0, 0;\n\t"
        "@p st.shared.v2.u32 [%1], {%2, %3};\n\t"
        "}\n" ::"r"(ok),
        "r"(addr), "r"(a), "r"(b)
        : "memory");
}

// Dimension chunk per stage: whole rows up to 64 dimensions, else an even split
// (96 = 3 x 32, 136 = 4 x 34) so the accumulators persist across the chunks of one tile pair.
__host__ __device__ constexpr int dim_chunk(int D) {
    if (D <= 64) return D;
    for (int c = 34; c >= 16; --c)
        if (D % c == 0) return c;
    return D;
}

// Ring depth: about 32 KB of candidate stages in flight, so that small-D tiles (a few hundred
// cycles of work each) still hide the ~1 us latency of a bulk copy.
__host__ __device__ constexpr int pick_nst(int D) {
    if (D > 64) return 3;
    const int n = 32768 / ((dim_chunk(D) + 1) * GMD_TW * 4);
    return n < 4 ? 4 : (n > 12 ? 12 : n);
}

// d-loop shape inside a chunk: a rolled loop of U steps per iteration, U a divisor of DC.
// The accumulators are initialised before the loop (from the norm row / zero) and every
// step has the same shape; with a peeled first step ptxas failed to coalesce the
// loop-carried accumulators and left up to 37 register moves per iteration in the loop.
__host__ __device__ constexpr int pick_unroll(int DC) {
    if (DC <= 12) return DC;  // whole chunk straight-line
    if (DC % 10 == 0) return 10;
    for (int u = 17; u >= 2; --u)
        if (DC % u == 0) return u;
    return DC;
}

// Operands of one d-step: 8 query values (broadcast operands) and 8 candidate values
// (4 packed pairs).
struct Frag {
    float4 qa, qb;  // q0..q3, q4..q7
    float4 c0, c1;  // candidates 0..3 | 4..7
};

// FIRST: the accumulators are written, not updated: acc = q*c + cn (expanded form, cn = the
// norm row) or the first term alone (direct forms).  Only used where the d-steps of a tile
// are straight-line code; in a rolled loop a peeled first step makes ptxas copy the
// loop-carried accumulators around.
template <int FORM, bool FIRST = false>
__device__ __forceinline__ void tile_step(float2 (&acc)[8][4], const Frag& f, const float2* cn = nullptr) {
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
    static constexpr int NQT = BQ / GMD_TW;             // query tiles of the CTA
    static constexpr int DC = dim_chunk(D);             // dimensions per stage
    static constexpr int NCH = D / DC;                  // stages per tile pair
    static constexpr int TILE_F = (D + 1) * GMD_TW;     // floats of one HBM tile
    static constexpr int STAGE_F = (DC + 1) * GMD_TW;   // floats of one stage (chunk rows + norm row)
    static size_t smem_bytes(int kk, int form) {
        // pending hits: HP_CELLS 8 cells x QL x 8 B per query; HP_QUEUE QCAP x 8 B + count + scratch row
        const size_t pend = pick_hitpath(D, form) == HP_CELLS ? (size_t)8 * QL * 8 : (size_t)QCAP * 8 + 4 + 32;
        return (size_t)(NST * STAGE_F + NQT * TILE_F + BQ) * 4 + (2 * NST + 2) * 8 + (size_t)BQ * pend +
               (size_t)BQ * (kk | 1) * 8;
    }
};

template <int D, int FORM, int NW, int NST>
__global__ void __launch_bounds__((NW + 1) * 32, 1) scan_topk_kernel(const ScanParams p) {
    using S = ScanShape<D, NW, NST>;
    constexpr int BQ = S::BQ, NQT = S::NQT, DC = S::DC, NCH = S::NCH, TILE_F = S::TILE_F, STAGE_F = S::STAGE_F;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr int HP = pick_hitpath(D, FORM);
    constexpr int PEND_B = HP == HP_CELLS ? 8 * QL * 8 : QCAP * 8 + 4 + 32;  // bytes per query
    constexpr int U = pick_unroll(DC);
    static_assert(DC % U == 0, "d-loop shape");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* Cs = reinterpret_cast<float*>(smem_raw);                // NST x [DC+1][64]
    float* Qs = Cs + NST * STAGE_F;                                // NQT x [D+1][64]  (HBM tile layout)
    float* thr_s = Qs + NQT * TILE_F;                              // [BQ]  heap-root value (+inf until full)
    uint64_t* bars = reinterpret_cast<uint64_t*>(thr_s + BQ);      // full[NST], empty[NST], queries, pad
    unsigned char* pend_raw = reinterpret_cast<unsigned char*>(bars + 2 * NST + 2);
    // HP_CELLS: NW x [8 rows][QL][32 lanes] (index, value bits)
    uint2* cells = reinterpret_cast<uint2*>(pend_raw);
    // HP_QUEUE: [BQ][QCAP] (index, value bits) | [BQ][8] scratch rows | [BQ] counts
    uint2* queue = reinterpret_cast<uint2*>(pend_raw);
    float* rowbuf = reinterpret_cast<float*>(pend_raw + (size_t)BQ * QCAP * 8);
    int* qcnt = reinterpret_cast<int*>(pend_raw + (size_t)BQ * (QCAP * 8 + 32));
    uint64_t* lists = reinterpret_cast<uint64_t*>(pend_raw + (size_t)BQ * PEND_B);  // [BQ][kkp]  max-heaps

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

    // ---- stage the CTA's query tiles with bulk copies (they arrive while the heaps are
    // initialised), then scale them in place (x -2 for the expanded form)
    const int qt0 = (int)(q0 / GMD_TW);
    const int nqt = min(NQT, p.ntiles_q - qt0);
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&bars[s], 1);
            mbar_init(&bars[NST + s], NW);
        }
        mbar_init(&bars[2 * NST], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bars[2 * NST], (uint32_t)nqt * TILE_F * 4);
        for (int j = 0; j < nqt; ++j)
            bulk_g2s(Qs + j * TILE_F, p.qtiles + (size_t)(qt0 + j) * TILE_F, TILE_F * 4, &bars[2 * NST]);
    }
    for (int e = tid; e < BQ * kkp; e += (NW + 1) * 32) lists[e] = GMD_EMPTY_KEY;
    for (int e = tid; e < BQ; e += (NW + 1) * 32) {
        thr_s[e] = __int_as_float(0x7f800000);
        if (HP == HP_QUEUE) qcnt[e] = 0;
    }
    for (int e = nqt * TILE_F + tid; e < NQT * TILE_F; e += (NW + 1) * 32) Qs[e] = 0.0f;  // tiles past the end
    __syncthreads();
    mbar_wait(&bars[2 * NST], 0);
    if (FORM == GMD_FORM_EXPANDED) {
        for (int e = tid; e < nqt * TILE_F; e += (NW + 1) * 32)
            if (e % TILE_F < D * GMD_TW) Qs[e] *= -2.0f;
    }
    __syncthreads();
    // no CTA-wide barrier below this line: the producer warp retires early

    if (warp == NW) {
        // ---- producer: keep the ring full.  Stage `it` = (tile it / NCH, dimension chunk it % NCH).
        if (lane == 0) {
            const int nit = (te - tb) * NCH;
            for (int it = 0; it < nit; ++it) {
                const int s = it % NST;
                if (it >= NST) mbar_wait_backoff(&bars[NST + s], ((it / NST) - 1) & 1, 16 * DC);
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

    // ---- compute warps: warp w owns queries [32w, 32w+32) = half (w & 1) of query tile w >> 1
    const int lq = lane >> 3, lc = lane & 7;
    const float* Qw = Qs + (warp >> 1) * TILE_F + (warp & 1) * 32 + lq * 4;
    const float* qnw = Qs + (warp >> 1) * TILE_F + D * GMD_TW + (warp & 1) * 32;  // ||q||^2 of the warp's queries
    float* wthr = thr_s + warp * 32;
    uint2* wcells = cells + warp * (8 * QL * 32);
    const uint32_t mycell = smem_u32(wcells + lane);  // + (row * QL + slot) * 256 bytes
    unsigned cnts = 0;                                // fill of this lane's 8 cells, a nibble per row
    int* wcnt = qcnt + warp * 32;
    uint2* wqueue = queue + warp * 32 * QCAP;
    float4* myrow = reinterpret_cast<float4*>(rowbuf + (size_t)(warp * 32 + lane) * 8);
    uint64_t* wheap = lists + (size_t)warp * 32 * kkp;
    // register copy of the thresholds of this lane's 8 rows (re-reading them from shared
    // memory every tile put an LDS on the selection's critical path: -2 % at D = 30)
    float thr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) thr[i] = __int_as_float(0x7f800000);
    int s = 0;             // ring stage and its phase parity
    uint32_t ph = 0;
    for (int t = tb; t < te; ++t) {
        float2 acc2[8][4];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            mbar_wait(&bars[s], ph);
            const float* cp = Cs + s * STAGE_F + lc * 4;   // running operand pointers
            const float* qp = Qw + ch * DC * GMD_TW;
            auto load = [&](int d) {
                Frag f;
                f.qa = *reinterpret_cast<const float4*>(qp + d * GMD_TW);
                f.qb = *reinterpret_cast<const float4*>(qp + d * GMD_TW + 16);
                f.c0 = *reinterpret_cast<const float4*>(cp + d * GMD_TW);
                f.c1 = *reinterpret_cast<const float4*>(cp + d * GMD_TW + 32);
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
            // software pipeline one step ahead in registers; the read one row past the chunk
            // at its end is harmless (norm row of the stage / of the query tile)
            Frag cur = load(0);
            if constexpr (U == DC && NCH == 1) {
                // short profiles: straight-line steps, the first one writes the accumulators
#pragma unroll
                for (int u = 0; u < DC; ++u) {
                    const Frag nxt = load(u + 1);
                    if (u == 0)
                        tile_step<FORM, true>(acc2, cur, cn);
                    else
                        tile_step<FORM>(acc2, cur);
                    cur = nxt;
                }
            } else {
                if (ch == 0) {
                    // accumulators start from ||c||^2 or zero, then every step has the same shape
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int x = 0; x < 4; ++x) acc2[i][x] = cn[x];
                }
#pragma unroll 1
                for (int d0 = 0; d0 < DC; d0 += U) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const Frag nxt = load(u + 1);
                        tile_step<FORM>(acc2, cur);
                        cur = nxt;
                    }
                    qp += U * GMD_TW;
                    cp += U * GMD_TW;
                }
            }
            // this stage's shared memory is no longer needed by this warp
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[NST + s]);
            if (++s == NST) {
                s = 0;
                ph ^= 1;
            }
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
          if constexpr (HP == HP_CELLS) {
            // ---- hit path.  Rows with a hit in ANY lane are visited (warp-uniform branches);
            // a lane appends its values under the threshold to its own cell of that row.
            // Thresholds stay as they are until a cell fills up: then lane L folds query L's
            // cells into its heap and refreshes the threshold.  A stale threshold is an upper
            // bound of the true one, so nothing is lost, and the heap orders by the full
            // (value, index) key, so the result does not depend on queueing order.
            const uint32_t cbase = (uint32_t)t * GMD_TW + (lc << 2);
            unsigned rows = __reduce_or_sync(FULL, rf);
            unsigned left_lo = 0xFFFFFFFFu, left_hi = 0xFFFFFFFFu;  // values still to be queued, a byte per row
            bool again;
            do {
                bool full = false;
                unsigned over = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (rows & (1u << i)) {
                        const float v[8] = {acc2[i][0].x, acc2[i][0].y, acc2[i][1].x, acc2[i][1].y,
                                            acc2[i][2].x, acc2[i][2].y, acc2[i][3].x, acc2[i][3].y};
                        unsigned mask = 0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) mask |= (v[j] < thr[i]) ? (1u << j) : 0u;
                        mask &= ((i < 4 ? left_lo : left_hi) >> ((i & 3) * 8)) & 0xFFu;
                        const unsigned cnt = (cnts >> (4 * i)) & 15u;
                        unsigned slot = cnt, lost = 0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const unsigned b = (mask >> j) & 1u;
                            const unsigned ok = b & (slot < QL ? 1u : 0u);
                            // value j: candidate group j>>2 (32 rows apart), element j&3
                            sts64_if(ok, mycell + (i * QL + slot) * 256u, cbase + ((j >> 2) << 5) + (j & 3),
                                     __float_as_uint(v[j]));
                            lost |= (b & (ok ^ 1u)) << j;
                            slot += b;
                        }
                        const unsigned fill = min(slot, (unsigned)QL);
                        cnts = (cnts & ~(15u << (4 * i))) | (fill << (4 * i));
                        full |= fill == QL;
                        over |= lost ? (1u << i) : 0u;
                        if (i < 4)
                            left_lo = (left_lo & ~(0xFFu << (i * 8))) | (lost << (i * 8));
                        else
                            left_hi = (left_hi & ~(0xFFu << ((i - 4) * 8))) | (lost << ((i - 4) * 8));
                    }
                }
                again = false;
                if (__any_sync(FULL, full)) {
                    __syncwarp();
                    drain_cells(wheap + (size_t)lane * kkp, kk, wcells, cnts, lane, &wthr[lane]);
                    cnts = 0;
                    __syncwarp();
#pragma unroll
                    for (int ii = 0; ii < 8; ++ii) thr[ii] = wthr[((ii >> 2) << 4) + (lq << 2) + (ii & 3)];
                    // values that found their cell full (heap warm-up, in practice): once more
                    rows = __reduce_or_sync(FULL, over);
                    again = rows != 0;
                }
            } while (again);
          } else {
            // ---- hit path, HP_QUEUE: a warp-uniform loop in which every lane with hits handles
            // ONE of its rows per pass.  The row's 8 values go through a 32-byte shared-memory
            // scratch (stores from the statically indexed registers, one load back), so the
            // body exists once.  A lane appends (index, value) pairs to the owning query's
            // queue (one shared-memory atomic reserves the slots); thresholds stay as they
            // are.  Only when a queue reaches QHIGH entries or overflows does the warp drain.
            const uint32_t cbase = (uint32_t)t * GMD_TW + (lc << 2);
            unsigned left = 0;  // values of the current row that found the queue full
            do {
                const int i = __ffs(rf) - 1;  // this lane's row for the pass (-1: none)
#pragma unroll
                for (int ii = 0; ii < 8; ++ii) {
                    if (i == ii) {
                        myrow[0] = make_float4(acc2[ii][0].x, acc2[ii][0].y, acc2[ii][1].x, acc2[ii][1].y);
                        myrow[1] = make_float4(acc2[ii][2].x, acc2[ii][2].y, acc2[ii][3].x, acc2[ii][3].y);
                    }
                }
                bool drain = false;
                if (i >= 0) {
                    const int ql = ((i >> 2) << 4) + (lq << 2) + (i & 3);
                    const float th = wthr[ql];
                    const float4 r0 = myrow[0], r1 = myrow[1];
                    const float v[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                    unsigned mask = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) mask |= (v[j] < th) ? (1u << j) : 0u;
                    if (left) mask &= left;
                    left = 0;
                    if (mask) {
                        int slot = atomicAdd(&wcnt[ql], __popc(mask));
                        uint2* qd = wqueue + ql * QCAP;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (mask & (1u << j)) {
                                // value j: candidate group j>>2 (32 rows apart), element j&3
                                if (slot < QCAP)
                                    qd[slot] = make_uint2(cbase + ((j >> 2) << 5) + (j & 3), __float_as_uint(v[j]));
                                else
                                    left |= 1u << j;
                                ++slot;
                            }
                        }
                        drain = slot >= QHIGH;
                    }
                    if (!left) rf &= rf - 1;  // row done
                }
                if (__any_sync(FULL, drain)) {
                    __syncwarp();
                    drain_queue(wheap + (size_t)lane * kkp, kk, wqueue + lane * QCAP, &wcnt[lane], &wthr[lane]);
                    __syncwarp();
#pragma unroll
                    for (int ii = 0; ii < 8; ++ii) thr[ii] = wthr[((ii >> 2) << 4) + (lq << 2) + (ii & 3)];
                }
            } while (__any_sync(FULL, rf != 0));
          }
        }
    }

    // ---- fold what is still queued into the heaps
    __syncwarp();
    if constexpr (HP == HP_CELLS)
        drain_cells(wheap + (size_t)lane * kkp, kk, wcells, cnts, lane, &wthr[lane]);
    else
        drain_queue(wheap + (size_t)lane * kkp, kk, wqueue + lane * QCAP, &wcnt[lane], &wthr[lane]);

    // ---- write this warp's proposals: cand_*[q][split][kk], cand_tau[q][split]
    __syncwarp();
    for (int ql = 0; ql < 32; ++ql) {
        const long long q = q0 + warp * 32 + ql;
        if (q >= p.Nq) break;
        const uint64_t* h = wheap + (size_t)ql * kkp;
        const float qnv = qnw[ql];
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
    const size_t smem = S::smem_bytes(p.kk, FORM);
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
    constexpr int NST = pick_nst(D);
    if constexpr (D <= 64) {
        if (ScanShape<D, 8, NST>::smem_bytes(p.kk, FORM) <= GMD_SMEM_LIMIT) return scan_launch<D, FORM, 8, NST>(p, st);
        return scan_launch<D, FORM, 4, NST>(p, st);
    } else {
        return scan_launch<D, FORM, 4, NST>(p, st);
    }
}

// ------------------------------------------------------------------------------------------
// k = 1 (best-matching unit, torusMesh.py:127): no heaps, no thresholds, no hit path.
//
// Same CTA shape, ring and register tile as scan_topk_kernel (expanded form only, D <= 32,
// straight-line d-steps).  A "cell" is the 8 candidates one lane holds of one tile.  Every
// (lane, row) keeps, branch-free, the smallest cell minimum it has seen, the tile it came
// from (strict <, so the earliest tile wins ties) and the second smallest cell minimum.
// After the scan the 8 lanes that share a query row merge their records with shuffles: the
// winning cell's 8 candidates are proposed (cand_* with kk = 8, nsplit = 1) and
// tau = min(second smallest cell minimum) bounds every candidate that was not.  The float64
// re-rank and certificate of refine.cu then work unchanged.
constexpr int BMU_KK = 8;

template <int D, int NW, int NST>
struct BmuShape {
    static constexpr int BQ = NW * 32;
    static constexpr int NQT = BQ / GMD_TW;
    static constexpr int TILE_F = (D + 1) * GMD_TW;
    static size_t smem_bytes() { return (size_t)(NST * TILE_F + NQT * TILE_F) * 4 + (2 * NST + 2) * 8; }
};

template <int D, int NW, int NST>
__global__ void __launch_bounds__((NW + 1) * 32, 1) scan_bmu_kernel(const ScanParams p) {
    using S = BmuShape<D, NW, NST>;
    constexpr int BQ = S::BQ, NQT = S::NQT, TILE_F = S::TILE_F;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    constexpr int FORM = GMD_FORM_EXPANDED;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* Cs = reinterpret_cast<float*>(smem_raw);            // NST x [D+1][64]
    float* Qs = Cs + NST * TILE_F;                             // NQT x [D+1][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(Qs + NQT * TILE_F);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long q0 = (long long)blockIdx.x * BQ;
    const int tb = 0, te = p.ntiles_c;
    const int qt0 = (int)(q0 / GMD_TW);
    const int nqt = min(NQT, p.ntiles_q - qt0);
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&bars[s], 1);
            mbar_init(&bars[NST + s], NW);
        }
        mbar_init(&bars[2 * NST], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bars[2 * NST], (uint32_t)nqt * TILE_F * 4);
        for (int j = 0; j < nqt; ++j)
            bulk_g2s(Qs + j * TILE_F, p.qtiles + (size_t)(qt0 + j) * TILE_F, TILE_F * 4, &bars[2 * NST]);
    }
    for (int e = nqt * TILE_F + tid; e < NQT * TILE_F; e += (NW + 1) * 32) Qs[e] = 0.0f;
    __syncthreads();
    mbar_wait(&bars[2 * NST], 0);
    for (int e = tid; e < nqt * TILE_F; e += (NW + 1) * 32)
        if (e % TILE_F < D * GMD_TW) Qs[e] *= -2.0f;
    __syncthreads();

    if (warp == NW) {
        if (lane == 0) {
            for (int it = 0; it < te - tb; ++it) {
                const int s = it % NST;
                if (it >= NST) mbar_wait_backoff(&bars[NST + s], ((it / NST) - 1) & 1, 16 * D);
                mbar_expect_tx(&bars[s], TILE_F * 4);
                bulk_g2s(Cs + s * TILE_F, p.ctiles + (size_t)(tb + it) * TILE_F, TILE_F * 4, &bars[s]);
            }
        }
        return;
    }

    const int lq = lane >> 3, lc = lane & 7;
    const float* Qw = Qs + (warp >> 1) * TILE_F + (warp & 1) * 32 + lq * 4;
    const float* qnw = Qs + (warp >> 1) * TILE_F + D * GMD_TW + (warp & 1) * 32;
    const float INF = __int_as_float(0x7f800000);
    float best[8], second[8];
    int btile[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) best[i] = INF, second[i] = INF, btile[i] = 0;

    int s = 0;
    uint32_t ph = 0;
    for (int t = tb; t < te; ++t) {
        float2 acc2[8][4];
        mbar_wait(&bars[s], ph);
        const float* cp = Cs + s * TILE_F + lc * 4;
        auto load = [&](int d) {
            Frag f;
            f.qa = *reinterpret_cast<const float4*>(Qw + d * GMD_TW);
            f.qb = *reinterpret_cast<const float4*>(Qw + d * GMD_TW + 16);
            f.c0 = *reinterpret_cast<const float4*>(cp + d * GMD_TW);
            f.c1 = *reinterpret_cast<const float4*>(cp + d * GMD_TW + 32);
            return f;
        };
        float2 cn[4];
        {
            const Frag n = load(D);
            cn[0] = make_float2(n.c0.x, n.c0.y), cn[1] = make_float2(n.c0.z, n.c0.w);
            cn[2] = make_float2(n.c1.x, n.c1.y), cn[3] = make_float2(n.c1.z, n.c1.w);
        }
        Frag cur = load(0);
#pragma unroll
        for (int u = 0; u < D; ++u) {
            const Frag nxt = load(u + 1);
            if (u == 0)
                tile_step<FORM, true>(acc2, cur, cn);
            else
                tile_step<FORM>(acc2, cur);
            cur = nxt;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[NST + s]);
        if (++s == NST) {
            s = 0;
            ph ^= 1;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float m = fminf(acc2[i][0].x, acc2[i][0].y);
#pragma unroll
            for (int x = 1; x < 4; ++x) m = fminf(m, fminf(acc2[i][x].x, acc2[i][x].y));
            second[i] = fminf(second[i], fmaxf(best[i], m));
            btile[i] = (m < best[i]) ? t : btile[i];
            best[i] = fminf(best[i], m);
        }
    }

    // ---- merge the records of the 8 lanes (lc) that share each query row
    uint32_t win[8];                                      // winning cell per row i (the same in the 8 lanes of a row)
    float wb1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float b1 = best[i], b2 = second[i];
        uint32_t k1 = ((uint32_t)btile[i] << 3) | (uint32_t)lc;  // cell id: ascending with the candidate index
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const float ob1 = __shfl_xor_sync(FULL, b1, o), ob2 = __shfl_xor_sync(FULL, b2, o);
            const uint32_t ok1 = __shfl_xor_sync(FULL, k1, o);
            const bool take = ob1 < b1 || (ob1 == b1 && ok1 < k1);
            b2 = fminf(fminf(b2, ob2), take ? b1 : ob1);
            k1 = take ? ok1 : k1;
            b1 = take ? ob1 : b1;
        }
        win[i] = k1;
        wb1[i] = b1;
        // lane lc proposes candidate j = lc of the winning cell
        const int ql = ((i >> 2) << 4) + (lq << 2) + (i & 3);
        const long long q = q0 + warp * 32 + ql;
        if (q < p.Nq) {
            const float qnv = qnw[ql];
            const uint32_t wt = k1 >> 3, wl = k1 & 7u;
            const uint32_t li = wt * GMD_TW + (wl << 2) + (((uint32_t)lc >> 2) << 5) + ((uint32_t)lc & 3u);
            const bool real = (long long)li < p.Nc && b1 < INF;
            const size_t ob = (size_t)q * BMU_KK + lc;
            p.cand_idx[ob] = real ? li + (uint32_t)p.c_index_offset : GMD_EMPTY_IDX;
            p.cand_val[ob] = real ? fmaxf(b1 + qnv, 0.0f) : INF;
            if (lc == 0) p.cand_tau[q] = (b2 < INF) ? fmaxf(b2 + qnv, 0.0f) : INF;
        }
    }
    if (p.best1) {
        // which of the winning cell's 8 candidates it was: lane lc recomputes candidate lc of every row's winning
        // cell from the tile in global memory (L1/L2 hits: callers with a few hundred candidates) -- the 8 rows'
        // chains side by side -- and 3 shuffles per row pick the minimum (rows past Nq compute on tile 0 and
        // write nothing; every lane of the warp takes part)
        float v[8];
        const float* ct[8];
        const float* qc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t wt = win[i] >> 3, wl = win[i] & 7u;
            ct[i] = p.ctiles + (size_t)wt * TILE_F + (wl << 2) + (((uint32_t)lc >> 2) << 5) + ((uint32_t)lc & 3u);
            qc[i] = Qs + (warp >> 1) * TILE_F + (warp & 1) * 32 + ((i >> 2) << 4) + (lq << 2) + (i & 3);   // scaled by -2
            v[i] = ct[i][D * GMD_TW];
        }
#pragma unroll
        for (int d = 0; d < D; ++d)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(qc[i][d * GMD_TW], ct[i][d * GMD_TW], v[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t wt = win[i] >> 3, wl = win[i] & 7u;
            const uint32_t li = wt * GMD_TW + (wl << 2) + (((uint32_t)lc >> 2) << 5) + ((uint32_t)lc & 3u);
            float vv = ((long long)li < p.Nc && wb1[i] < INF) ? v[i] : INF;
            uint32_t bi = li;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const float ov = __shfl_xor_sync(FULL, vv, o);
                const uint32_t oi = __shfl_xor_sync(FULL, bi, o);
                if (ov < vv || (ov == vv && oi < bi)) vv = ov, bi = oi;
            }
            const long long q = q0 + warp * 32 + ((i >> 2) << 4) + (lq << 2) + (i & 3);
            if (lc == 0 && q < p.Nq) p.best1[q] = bi + (uint32_t)p.c_index_offset;
        }
    }
}

template <int D>
static int bmu_launch(const ScanParams& p, cudaStream_t st) {
    constexpr int NW = 8, NST = pick_nst(D);
    using S = BmuShape<D, NW, NST>;
    const size_t smem = S::smem_bytes();
    auto kern = scan_bmu_kernel<D, NW, NST>;
    GMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)((p.Nq + S::BQ - 1) / S::BQ), (NW + 1) * 32, smem, st>>>(p);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

}  // namespace gmd_scan

// padded dimensions the scan is instantiated for; gmd_scan_pad_dim(D) is the smallest >= D
#define GMD_SCAN_DIMS(X) \
    X(2) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(20) X(24) X(28) X(30) X(32) X(40) X(48) X(64) X(96) X(136)

// padded dimensions of the k = 1 kernel
#define GMD_BMU_DIMS(X) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(20) X(24) X(28) X(30) X(32)
#define GMD_BMU_MAX_D 32

int gmd_scan_pad_dim(int D);
int gmd_bmu_launch(int DP, const gmd_scan::ScanParams& p, cudaStream_t st);
int gmd_scan_launch_form0(int DP, const gmd_scan::ScanParams& p, cudaStream_t st);
int gmd_scan_launch_form1(int DP, const gmd_scan::ScanParams& p, cudaStream_t st);
int gmd_scan_launch_form2(int DP, const gmd_scan::ScanParams& p, cudaStream_t st);
int gmd_scan_block_queries(int DP);
