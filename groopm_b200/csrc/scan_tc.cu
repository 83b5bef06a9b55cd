// Tensor-core variant of the fused distance + top-k scan (expanded form, Euclidean, 4 <= D <= 136).
//
// The dot products q.c of a 128-query x 128-candidate tile are a contraction on the
// 5th-generation tensor cores: tcgen05.mma kind::f16 (FP16 operands, FP32 accumulation in
// tensor memory), operands K-major SWIZZLE_64B images in shared memory, the D + 1 columns cut
// into NKB = ceil((D + 1) / 32) K blocks of 32 that accumulate into the same tensor-memory
// stage.  One FP16 (11-bit significand) is far too coarse for a neighbour search, so every
// centred FP32 value, scaled by a power of two that brings the largest row norm into (128, 256],
// is split v = hi + lo (both FP16) and the contraction is issued three times: hi.hi + hi.lo +
// lo.hi.  Measured against float64 (tools/tc_probe16.cu): |error| <= 2^-21.9 |q||c|, i.e.
// FP32-class, inside the certificate's bound (refine.cu, GMD_FORM_TC).  The same split with
// TF32 operands (tools/tc_probe.cu, 2^-20.9) needs twice the tensor time and twice the bytes;
// it was the first version of this kernel.
//
// The norm rides in the contraction: column D of a QUERY image holds 8, column D of a
// CANDIDATE image holds -||c||^2 / 16, so the accumulator is  u = q.c - ||c||^2/2 = -s/2  with
// s = ||c||^2 - 2 q.c the value the FP32 scan ranks by; selection maximises u directly (no
// per-pair FMA, no norm loads) and converts to s = -2u (exact) only for the few values it keeps.
//
// Roles (224 threads).  Warps 0..3 are the selection warps: thread r owns query r of the CTA =
// TMEM lane r, reads its 128 accumulators with tcgen05.ld, and keeps the query's top-kk heap in
// shared memory all by itself (no atomics: a thread sees every candidate of its query, in
// ascending index order).  Warp 4 lane 0 streams candidate K blocks HBM -> shared memory with
// bulk async copies (mbarrier ring, one 16 KB K block per stage).  Warps 5 and 6, lane 0 each,
// issue the MMAs (6 per K block) of the even and of the odd candidate tiles into one of the four
// 128-column accumulator stages and commit completion to mbarriers.  Two issuers because the
// barrier probes, fence and commit of a tile cost the issuing thread ~290 cycles in which it
// feeds the tensor pipe nothing (tools/tc_smem.cu: 672 instead of 384 cycles per tile of 6
// MMAs); with two threads one is always issuing (MMA + copy pipeline alone: 159 -> 134 ms on
// 1M x 30).  The single-thread roles have the highest warp ids: a sub-partition's scheduler
// picks the eligible warp with the highest id first.
//
// HBM tile format ("TC tile", 128 profiles): NKB x [hi image 8 KB][lo image 8 KB], then
// [scaled ||x||^2 128 f32][pad to 1 KB]; images already in the swizzled shared-memory layout
// (8-row x 64 B atoms, 16-byte chunk c of row r at c ^ ((r >> 1) & 3)) so that one bulk copy
// per K block is all the producer does.  Query and candidate tiles differ in column D only.
#include <stdlib.h>

#include <algorithm>

#include "tc_common.cuh"

// Developer probes of the scan (GMD_TC_DBG: pipeline without selection, no candidate traffic, one issuer ...)
// exist only in builds made with -DGMD_TC_PROBES; the shipped kernel has none of their branches.
#ifdef GMD_TC_PROBES
#define TC_DBG(x) (x)
#else
#define TC_DBG(x) 0
#endif

namespace gmd_tc {

// ------------------------------------------------------------------ staging: float64 rows -> TC tiles
// pass 1: largest centred row norm of a profile set (fixes the power-of-two scale)
__global__ void __launch_bounds__(256) tc_norms_kernel(const double* __restrict__ X, long long N, int D,
                                                       const double* __restrict__ center, float* __restrict__ stats,
                                                       int* __restrict__ flags) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float n2f = 0.0f, n1f = 0.0f;
    bool bad = false;
    if (row < N) {
        double n2 = 0.0, n1 = 0.0;
        for (int k = 0; k < D; ++k) {
            const double x = X[(size_t)row * D + k];
            bad |= !isfinite(x);
            const double v = (double)(float)(x - (center ? center[k] : 0.0));
            n2 += v * v;
            n1 += fabs(v);
        }
        n2f = __double2float_ru(n2);
        n1f = __double2float_ru(n1);
    }
    for (int o = 16; o > 0; o >>= 1) {
        n2f = fmaxf(n2f, __shfl_xor_sync(0xFFFFFFFFu, n2f, o));
        n1f = fmaxf(n1f, __shfl_xor_sync(0xFFFFFFFFu, n1f, o));
    }
    if ((threadIdx.x & 31) == 0 && isfinite(n2f)) {
        atomicMax(reinterpret_cast<int*>(stats), __float_as_int(n2f));
        atomicMax(reinterpret_cast<int*>(stats) + 1, __float_as_int(n1f));
    }
    if (bad) atomicExch(flags, 1);
}

// pass 2: scaled FP16 hi / lo images, nkb K blocks of 32 columns.  role 0: query tile (column D =
// QCOL), role 1: candidate tile (column D = -||x||^2 / (2 QCOL), in scaled units).
// stats_c / stats_q: max ||x||^2 of the two sets.
// perm (nullable): tile position -> row of X, for profiles staged in a locality-preserving order;
// GMD_EMPTY_IDX marks a padding position (grouped order of order.cu), Npos counts positions.
// tmeta (nullable): [D + 1][nt_alloc] ball of the tile in scaled units -- centroid of its real rows and
// the largest distance of one of them to it, rounded up (-1 for a tile of padding only).  Any centre
// gives a valid ball; the radius is taken against the centre exactly as stored.
// The rows of a tile are read from global memory ONCE (a warp per row: coalesced 8 D bytes), kept in shared
// memory as the scaled FP32 values the images are made of, and everything else -- norms, hi / lo images written
// as 16-byte chunks (a warp stores 512 contiguous bytes), centroid, radius -- works from there.  role 2 writes
// both roles' tiles (tiles: queries, tiles2: candidates) for profiles that serve as both.
__global__ void __launch_bounds__(256) prep_tc_kernel(const double* __restrict__ X, long long Npos, int D, int nkb,
                                                      const uint32_t* __restrict__ perm,
                                                      const double* __restrict__ center, int role,
                                                      const float* __restrict__ stats_c, const float* __restrict__ stats_q,
                                                      unsigned char* __restrict__ tiles, unsigned char* __restrict__ tiles2,
                                                      float* __restrict__ tmeta, int nt_alloc) {
    extern __shared__ __align__(16) float xs[];          // [TM][P], P = D | 1: conflict-free by row and by column
    __shared__ float s_n2[TM];
    __shared__ long long s_src[TM];                      // row of X, -1 for padding
    __shared__ float s_m[GMD_MAX_D];
    __shared__ int s_cnt, s_rad;
    const int P = D | 1;
    const float scale = ldexpf(1.0f, gmd_tc_exponent(stats_c[0], stats_q[0]));
    const long long t = blockIdx.x, row0 = t * TM;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_cnt = 0, s_rad = 0;
    if (tid < TM) {
        const long long pos = row0 + tid;
        long long src = -1;
        if (pos < Npos) {
            if (perm) {
                const uint32_t r = perm[pos];
                if (r != GMD_EMPTY_IDX) src = (long long)r;
            } else {
                src = pos;
            }
        }
        s_src[tid] = src;
    }
    __syncthreads();
    for (int r = warp; r < TM; r += 8) {
        const long long src = s_src[r];
        for (int k = lane; k < D; k += 32)
            xs[r * P + k] = src >= 0 ? (float)(X[(size_t)src * D + k] - (center ? center[k] : 0.0)) * scale : 0.0f;   // exact: power of two
    }
    __syncthreads();
    if (tid < TM) {
        double n2 = 0.0;
#pragma unroll 6
        for (int k = 0; k < D; ++k) {
            const double v = (double)xs[tid * P + k];
            n2 += v * v;
        }
        s_n2[tid] = (float)n2;
        if (s_src[tid] >= 0) atomicAdd(&s_cnt, 1);
    }
    __syncthreads();
    // images: item (row r, 16-byte chunk c of K block kb) = 8 consecutive columns; thread = 4 r + c
    const size_t tb = (size_t)t * tile_bytes(nkb);
    for (int kb = 0; kb < nkb; ++kb) {
        for (int it = tid; it < TM * 4; it += 256) {
            const int r = it >> 2, c = it & 3, k0 = kb * TK + c * 8;
            const bool real = s_src[r] >= 0;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = k0 + j < D ? xs[r * P + k0 + j] : 0.0f;
            const int off = (r >> 3) * 512 + (r & 7) * 64 + (((c ^ ((r >> 1) & 3)) & 3) << 4);
            const bool has_d = D >= k0 && D < k0 + 8;
            auto put = [&](unsigned char* base, float cold) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const float a = (has_d && k0 + j == D) ? cold : v[j];
                    const float b = (has_d && k0 + j + 1 == D) ? cold : v[j + 1];
                    const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
                    const __half la = __float2half_rn(a - __half2float(ha)), lb = __float2half_rn(b - __half2float(hb));
                    hi[j >> 1] = (uint32_t)__half_as_ushort(ha) | ((uint32_t)__half_as_ushort(hb) << 16);
                    lo[j >> 1] = (uint32_t)__half_as_ushort(la) | ((uint32_t)__half_as_ushort(lb) << 16);
                }
                unsigned char* img = base + tb + (size_t)kb * KB_B + off;
                *reinterpret_cast<uint4*>(img) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(img + IMG_B) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            };
            const float qcol = QCOL, ccol = real ? -s_n2[r] / (2.0f * QCOL) : PAD_COL;
            if (role == 0) {
                put(tiles, qcol);
            } else if (role == 1) {
                put(tiles, ccol);
            } else {
                put(tiles, qcol);
                put(tiles2, ccol);
            }
        }
    }
    if (tid < TM) {
        const float nv = s_src[tid] >= 0 ? s_n2[tid] : 0.0f;
        reinterpret_cast<float*>(tiles + tb + (size_t)nkb * KB_B)[tid] = nv;
        if (role == 2) reinterpret_cast<float*>(tiles2 + tb + (size_t)nkb * KB_B)[tid] = nv;
    }
    if (!tmeta) return;
    // ---- the tile's ball
    const int cnt = s_cnt;
    for (int k = tid; k < D; k += blockDim.x) {
        double sum = 0.0;
        for (int r = 0; r < TM; ++r)
            if (s_src[r] >= 0) sum += (double)xs[r * P + k];
        const float m = cnt ? (float)(sum / cnt) : 0.0f;
        s_m[k] = m;
        tmeta[(size_t)k * nt_alloc + t] = m;
    }
    __syncthreads();
    if (tid < TM && s_src[tid] >= 0) {
        double d2 = 0.0;
#pragma unroll 6
        for (int k = 0; k < D; ++k) {
            const double u = (double)xs[tid * P + k] - (double)s_m[k];
            d2 += u * u;
        }
        atomicMax(&s_rad, __float_as_int(__double2float_ru(sqrt(d2) * (1.0 + 1e-9))));
    }
    __syncthreads();
    if (tid == 0) tmeta[(size_t)D * nt_alloc + t] = cnt ? __int_as_float(s_rad) : -1.0f;
}

// ------------------------------------------------------------------ the scan
// nkb: K blocks per tile (1 for D <= 31 ... 5 for D = 136); ns: ring stages of one K block each.
// NKB is a template parameter: with a run-time K-block loop the issuing thread falls behind the tensor
// pipe (1M x 30: 196 ms instead of 159 ms with selection off).
// SWEEP: the profiles were staged in a locality-preserving order (perm[position] = original row, the
// same for queries and candidates) and every CTA starts its sweep over the candidate tiles at its own
// query tile, wrapping around: a query meets its neighbourhood first, its threshold is near-final after a
// few tiles, and hits become rare (tools/sim_hits.py: a vote fires in ~20 instead of ~2960 of 7813
// tiles).  The selection warps then use the variant that is fastest when nothing hits: all 16 group
// maxima and ONE vote per tile.  Proposals are written under the ORIGINAL row numbers.
template <int NKB, bool SWEEP>
__global__ void __launch_bounds__(NTHREADS, 1) scan_tc_kernel(const unsigned char* __restrict__ qtiles,
                                                              const unsigned char* __restrict__ ctiles,
                                                              long long Nq, long long Nc, int ntiles_c, int kk,
                                                              int ns, const uint32_t* __restrict__ perm,
                                                              long long c_index_offset, int dbg,
                                                              const float* __restrict__ stats_c,
                                                              const float* __restrict__ stats_q,
                                                              uint32_t* __restrict__ cand_idx,
                                                              float* __restrict__ cand_val,
                                                              float* __restrict__ cand_tau) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // (an offset from the array, not an integer round trip: the compiler must keep seeing SHARED memory here, or
    // every access below becomes a generic LD.E / ST.E)
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int nkb = NKB;
    constexpr int TILE_B = tile_bytes(nkb);
    unsigned char* Aimg = base;                          // the CTA's query tile: nkb x (hi, lo) + norms
    unsigned char* Bst = base + TILE_B;                  // ns ring stages: one K block (hi, lo) of a candidate tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(Bst + (size_t)ns * KB_B);
    uint64_t* full = bars;                               // [NS_MAX] K block landed in shared memory
    uint64_t* sfree = bars + NS_MAX;                     // [NS_MAX] the MMAs that read the stage are complete
    uint64_t* tdone = bars + 2 * NS_MAX;                 // [NACC] all MMAs of a tile complete: accumulator stage written
    uint64_t* tempty = bars + 2 * NS_MAX + NACC;         // [NACC] accumulator stage read back by the selection warps
    uint64_t* qbar = bars + 2 * NS_MAX + 2 * NACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS_MAX + 2 * NACC + 2);
    uint2* queues = reinterpret_cast<uint2*>(bars + 2 * NS_MAX + 2 * NACC + 4);  // [TQ][128] parked (index, value bits)
    uint64_t* lists = reinterpret_cast<uint64_t*>(queues + TQ * TM);      // [128][kkp] heaps

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kkp = kk | 1;
    const long long qt = blockIdx.x;
    const int t_first = SWEEP ? (int)(qt % ntiles_c) : 0;   // sweep step t visits tile (t_first + t) mod ntiles_c

    if (tid == 0) {
        for (int s = 0; s < ns; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&sfree[s], 1);
        }
        for (int a = 0; a < NACC; ++a) {
            mbar_init(&tdone[a], 1);
            mbar_init(&tempty[a], 4);
        }
        mbar_init(qbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(qbar, TILE_B);
        bulk_g2s(Aimg, qtiles + (size_t)qt * TILE_B, TILE_B, qbar);
    }
    for (int e = tid; e < TM * kkp; e += NTHREADS) lists[e] = GMD_EMPTY_KEY;
    if (warp == W_ISSUE) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;

    if (warp == W_PROD) {
        // ---- producer: ring item `it` = K block it % nkb of candidate tile it / nkb
        if (lane == 0) {
            // A stage that held the last K block of a tile is free when the tile's accumulator is
            // complete (tdone), any other when the MMAs of its K block are (sfree): one commit per block.
            // ns <= NACC * nkb keeps the tdone parity unambiguous (tile pt + NACC is not loaded yet).
            const long long nit = (long long)ntiles_c * nkb;
            int s = 0, kb = 0, t = 0, pkb = 0, pt = 0;
            uint32_t sph = 0;                              // bit s: parity of the next sfree[s] phase
            for (long long it = 0; it < nit; ++it) {
                if (it >= ns) {
                    if (pkb == nkb - 1) {
                        mbar_wait_backoff(&tdone[pt % NACC], (uint32_t)(pt / NACC) & 1u, 64);
                    } else {
                        mbar_wait_backoff(&sfree[s], (sph >> s) & 1u, 64);
                        sph ^= 1u << s;
                    }
                    if (++pkb == nkb) pkb = 0, ++pt;
                }
                const uint32_t bytes = (TC_DBG(dbg) & 15) == 2 ? 1024u : (uint32_t)KB_B;  // (2: developer probe, no candidate traffic)
                mbar_expect_tx(&full[s], bytes);
                const int ta = t + t_first < ntiles_c ? t + t_first : t + t_first - ntiles_c;
                bulk_g2s(Bst + (size_t)s * KB_B, ctiles + (size_t)ta * TILE_B + (size_t)kb * KB_B, bytes, &full[s]);
                if (++kb == nkb) kb = 0, ++t;
                if (++s == ns) s = 0;
            }
        }
    } else if (warp >= W_ISSUE) {
        // ---- MMA issuers: warp W_ISSUE takes the even candidate tiles, the next warp the odd ones.
        // Tiles are independent (own accumulator stage, own ring stages) and tcgen05.commit tracks the
        // committing thread's own MMAs.
        if (lane == 0) {
            mbar_wait(qbar, 0);
            // FP32 accumulator, FP16 A and B (format 0), both K-major, N = M = 128
            const uint32_t idesc = (1u << 4) | ((uint32_t)(TM >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            // descriptors differ only in the 16-byte-granular start address: build them once and add
            const uint64_t a_base = umma_desc(smem_u32(Aimg));
            const uint64_t b_base = umma_desc(smem_u32(Bst));
            // Two issuers need ns > nkb: an issuer only sees every other fill of a ring stage, and a parity
            // wait is ambiguous if the fill before the awaited one may still be in flight.
            const int nissue = (ns > nkb && !(TC_DBG(dbg) & 256)) ? 2 : 1;  // (dbg & 256: developer probe, one issuer)
            const int first = warp - W_ISSUE;
            int s = 0;                                     // ring stage and phase of item t * nkb
            uint32_t ph = 0;
            auto skip_tile = [&]() {
                s += nkb;
                while (s >= ns) s -= ns, ph ^= 1;
            };
            if (first) skip_tile();
            if (first >= nissue) s = 0, ph = 0;
            for (int t = first; t < (first < nissue ? ntiles_c : 0); t += nissue) {
                const int a = t % NACC;
                if (t >= NACC) mbar_wait(&tempty[a], ((t / NACC) - 1) & 1);
                const uint32_t acc = tmem + (uint32_t)(a * TM);
#pragma unroll
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint64_t a_hi = a_base + (uint64_t)kb * (KB_B >> 4), a_lo = a_hi + (IMG_B >> 4);
                    const uint64_t b_hi = b_base + (uint64_t)s * (KB_B >> 4), b_lo = b_hi + (IMG_B >> 4);
                    // hi.hi + hi.lo + lo.hi, K = 16 halves (32 bytes) per MMA
#pragma unroll
                    for (int p = 0; p < 3; ++p)
#pragma unroll
                        for (int k = 0; k < TK / 16; ++k)
                            umma_f16(acc, (p == 2 ? a_lo : a_hi) + 2 * k, (p == 1 ? b_lo : b_hi) + 2 * k, idesc,
                                     (kb | p | k) ? 1u : 0u);
                    // one commit per K block: the last one of a tile signals the accumulator stage (and
                    // frees its ring stage with it), the others only their ring stage
                    umma_commit(kb + 1 < nkb ? &sfree[s] : &tdone[a]);
                    if (++s == ns) s = 0, ph ^= 1;
                }
                if (nissue == 2) skip_tile();              // the other issuer's tile
            }
        }
    } else {
        // ---- selection warps: thread <-> query row <-> TMEM lane
        const int quarter = warp & 3;                      // TMEM lanes this warp may read
        const int row = quarter * 32 + lane;
        uint64_t* h = lists + (size_t)row * kkp;
        uint2* myq = queues + row;                         // slot r of this thread at myq[r * TM]: conflict-free
        int cnt = 0;
        float thr = __int_as_float(0x7f800000);
        constexpr unsigned FULLW = 0xFFFFFFFFu;
        float uthr = __int_as_float(0xff800000);            // -thr / 2: the accumulators are u = -s/2
        if ((TC_DBG(dbg) & 15) == 4) uthr = __int_as_float(0x7f800000);    // developer probe: nothing ever hits
        for (int t = 0; t < ntiles_c; ++t) {
            const int a = t % NACC;
            mbar_wait(&tdone[a], (t / NACC) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t cbase = (uint32_t)(t + t_first < ntiles_c ? t + t_first : t + t_first - ntiles_c) * TM;
            const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * TM);
            // all 128 columns of the row in one load, then the accumulator stage goes back
            uint32_t r0[32], r1[32], r2[32], r3[32];
            tmem_ld128_sync4(trow, r0, r1, r2, r3);   // load and wait in ONE statement (see tc_common.cuh)
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[a]);
            if constexpr (SWEEP) {
                if ((TC_DBG(dbg) & 15) >= 1 && (TC_DBG(dbg) & 15) != 4) continue;  // developer probe: MMA + copy pipeline only
                // Maxima of the 16 groups of 8 columns first, as shallow trees with all 128 values in
                // flight (a dependent chain or a vote-and-branch per 32 columns is pure latency for the one
                // selection warp of a scheduler: 430 instead of ~200 cycles per tile), then ONE vote.
                float gm[16];
                auto tree = [&](const uint32_t (&r)[32], int b) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float m0 = fmaxf(fmaxf(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1])), __uint_as_float(r[8 * g + 2]));
                        const float m1 = fmaxf(fmaxf(__uint_as_float(r[8 * g + 3]), __uint_as_float(r[8 * g + 4])), __uint_as_float(r[8 * g + 5]));
                        const float m2 = fmaxf(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
                        gm[4 * b + g] = fmaxf(fmaxf(m0, m1), m2);
                    }
                };
                tree(r0, 0);
                tree(r1, 1);
                tree(r2, 2);
                tree(r3, 3);
                float cm;
                {
                    const float a0 = fmaxf(fmaxf(gm[0], gm[1]), gm[2]), a1 = fmaxf(fmaxf(gm[3], gm[4]), gm[5]);
                    const float a2 = fmaxf(fmaxf(gm[6], gm[7]), gm[8]), a3 = fmaxf(fmaxf(gm[9], gm[10]), gm[11]);
                    const float a4 = fmaxf(fmaxf(gm[12], gm[13]), gm[14]);
                    cm = fmaxf(fmaxf(fmaxf(a0, a1), a2), fmaxf(fmaxf(a3, a4), gm[15]));
                }
                if (!__any_sync(FULLW, cm > uthr)) continue;
                // which groups hold a hit for any lane: one warp-wide OR instead of 16 votes
                unsigned mine = 0;
#pragma unroll
                for (int g = 0; g < 16; ++g) mine |= (gm[g] > uthr) ? (1u << g) : 0u;
                const unsigned hot = __reduce_or_sync(FULLW, mine);
                // 8 values at a time: park what beats the (possibly stale) threshold, drain all the warp's
                // queues side by side when one of them is half full
                auto park = [&](const uint32_t (&r)[32], int b) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (hot & (1u << (4 * b + g))) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float u = __uint_as_float(r[8 * g + j]);
                                if (u > uthr) {
                                    myq[cnt * TM] = make_uint2(cbase + 32 * b + 8 * g + j, __float_as_uint(-2.0f * u));
                                    ++cnt;
                                }
                            }
                            if (__any_sync(FULLW, cnt > TQ - 8)) {
                                thr = drain_own(h, kk, myq, cnt);
                                uthr = -0.5f * thr;
                                cnt = 0;
                            }
                        }
                    }
                };
                park(r0, 0);
                park(r1, 1);
                park(r2, 2);
                park(r3, 3);
            } else {
                auto process = [&](const uint32_t (&r)[32], int c0) {
                    if ((TC_DBG(dbg) & 15) >= 1 && (TC_DBG(dbg) & 15) != 4) return;  // developer probe: MMA + copy pipeline only
                    // group maxima as trees (a chain of dependent ops is pure latency for the one
                    // selection warp of a scheduler), one vote per 32 columns
                    float gm[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float m01 = fmaxf(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1]));
                        const float m23 = fmaxf(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
                        const float m45 = fmaxf(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
                        const float m67 = fmaxf(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
                        gm[g] = fmaxf(fmaxf(m01, m23), fmaxf(m45, m67));
                    }
                    const float cm = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
                    if (__any_sync(FULLW, cm > uthr)) {
                        // 8 values at a time: park what beats the (possibly stale) threshold, drain all
                        // the warp's queues side by side when one of them is half full
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (__any_sync(FULLW, gm[g] > uthr)) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float u = __uint_as_float(r[8 * g + j]);
                                    if (u > uthr) {
                                        myq[cnt * TM] = make_uint2(cbase + c0 + 8 * g + j, __float_as_uint(-2.0f * u));
                                        ++cnt;
                                    }
                                }
                                if (__any_sync(FULLW, cnt > TQ - 8)) {
                                    thr = drain_own(h, kk, myq, cnt);
                                    uthr = -0.5f * thr;
                                    cnt = 0;
                                }
                            }
                        }
                    }
                };
                process(r0, 0);
                process(r1, 32);
                process(r2, 64);
                process(r3, 96);
            }
        }
        thr = drain_own(h, kk, myq, cnt);
        // ---- write this query's proposals
        const long long qpos = qt * TM + row;
        if (qpos < Nq) {
            const long long q = (SWEEP && perm) ? (long long)perm[qpos] : qpos;
            // back to the caller's units: everything above ran on rows scaled by 2^e
            const float unscale = ldexpf(1.0f, -2 * gmd_tc_exponent(stats_c[0], stats_q[0]));
            const float qnv = reinterpret_cast<const float*>(Aimg + (size_t)nkb * KB_B)[row];
            const size_t ob = (size_t)q * kk;
            for (int x = 0; x < kk; ++x) {
                const uint64_t key = h[x];
                const uint32_t li = (uint32_t)key;
                float vv = gmd_unflip((uint32_t)(key >> 32));
                uint32_t oi = GMD_EMPTY_IDX;
                if (key != GMD_EMPTY_KEY && (long long)li < Nc) {
                    oi = ((SWEEP && perm) ? perm[li] : li) + (uint32_t)c_index_offset;
                    vv = fmaxf(vv + qnv, 0.0f) * unscale;
                } else {
                    vv = __int_as_float(0x7f800000);
                }
                cand_idx[ob + x] = oi;
                cand_val[ob + x] = vv;
            }
            float tau = thr;
            if (h[0] != GMD_EMPTY_KEY) tau = fmaxf(tau + qnv, 0.0f) * unscale;
            cand_tau[q] = tau;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == W_ISSUE) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

}  // namespace gmd_tc

extern "C" int gmd_tc_supported(int k, int metric, int D) {
    // one column of the contraction carries the norm; up to MAX_KB blocks of 32 columns
    return metric == GMD_L2 && k >= 2 && k <= GMD_MAX_K && D >= 4 && D <= GMD_MAX_D && tc_nkb(D) <= gmd_tc::MAX_KB;
}

extern "C" size_t gmd_tc_tiles_bytes(int64_t N, int D) {
    if (N < 0 || D < 1 || tc_nkb(D) > gmd_tc::MAX_KB) return 0;
    return (size_t)((N + gmd_tc::TM - 1) / gmd_tc::TM) * gmd_tc::tile_bytes(tc_nkb(D));
}

extern "C" int gmd_tc_norms_f64(const double* X, int64_t N, int D, const double* center, float* stats, int32_t* flags,
                                void* stream) {
    GMD_REQUIRE(X && stats && flags && N > 0 && D >= 1 && tc_nkb(D) <= gmd_tc::MAX_KB, GMD_ERR_BAD_ARG,
                "tc_norms: bad arguments");
    gmd_tc::tc_norms_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, N, D, center, stats, flags);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

static int prep_tc_launch(const double* X, int64_t Npos, int D, const double* center, int role, const float* stats_c,
                          const float* stats_q, const uint32_t* perm, void* tiles, void* tiles2, float* tmeta,
                          int nt_alloc, void* stream) {
    GMD_REQUIRE(X && tiles && stats_c && stats_q && Npos > 0 && D >= 1 && tc_nkb(D) <= gmd_tc::MAX_KB &&
                    (role == 0 || role == 1 || (role == 2 && tiles2)),
                GMD_ERR_BAD_ARG, "prep_tc: bad arguments");
    const unsigned grid = (unsigned)((Npos + gmd_tc::TM - 1) / gmd_tc::TM);
    GMD_REQUIRE(!tmeta || (int64_t)grid <= nt_alloc, GMD_ERR_BAD_ARG, "prep_tc: tile meta too small");
    const size_t smem = (size_t)gmd_tc::TM * (D | 1) * sizeof(float);
    GMD_CUDA_CHECK(cudaFuncSetAttribute(gmd_tc::prep_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gmd_tc::prep_tc_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(X, Npos, D, tc_nkb(D), perm, center, role, stats_c,
                                                                     stats_q, (unsigned char*)tiles, (unsigned char*)tiles2,
                                                                     tmeta, nt_alloc);
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

extern "C" int gmd_prep_tc_grouped_f64(const double* X, int64_t Npos, int D, const double* center, int role,
                                       const float* stats_c, const float* stats_q, const uint32_t* perm, void* tiles,
                                       float* tmeta, int nt_alloc, void* stream) {
    GMD_REQUIRE(role == 0 || role == 1, GMD_ERR_BAD_ARG, "prep_tc: bad role %d", role);
    return prep_tc_launch(X, Npos, D, center, role, stats_c, stats_q, perm, tiles, nullptr, tmeta, nt_alloc, stream);
}

// one pass for profiles that are queries AND candidates (same rows, same order): both roles' tiles, one set of balls
extern "C" int gmd_prep_tc_grouped_both_f64(const double* X, int64_t Npos, int D, const double* center,
                                            const float* stats_c, const float* stats_q, const uint32_t* perm,
                                            void* qtiles, void* ctiles, float* tmeta, int nt_alloc, void* stream) {
    return prep_tc_launch(X, Npos, D, center, 2, stats_c, stats_q, perm, qtiles, ctiles, tmeta, nt_alloc, stream);
}

extern "C" int gmd_prep_tc_perm_f64(const double* X, int64_t N, int D, const double* center, int role,
                                    const float* stats_c, const float* stats_q, const uint32_t* perm, void* tiles,
                                    void* stream) {
    return gmd_prep_tc_grouped_f64(X, N, D, center, role, stats_c, stats_q, perm, tiles, nullptr, 0, stream);
}

extern "C" int gmd_prep_tc_f64(const double* X, int64_t N, int D, const double* center, int role,
                               const float* stats_c, const float* stats_q, void* tiles, void* stream) {
    return gmd_prep_tc_perm_f64(X, N, D, center, role, stats_c, stats_q, nullptr, tiles, stream);
}

extern "C" int gmd_scan_tc_sweep_f32(const void* qtiles, int64_t Nq, const void* ctiles, int64_t Nc, int D, int kk,
                                     int64_t c_index_offset, const float* stats_c, const float* stats_q,
                                     const uint32_t* perm, int sweep, uint32_t* cand_idx, float* cand_val,
                                     float* cand_tau, void* stream) {
    GMD_REQUIRE(qtiles && ctiles && stats_c && stats_q && cand_idx && cand_val && cand_tau, GMD_ERR_BAD_ARG,
                "scan_tc: null pointer");
    GMD_REQUIRE(Nq > 0 && Nc > 0 && Nc + c_index_offset < 0xFFFFFFFFll, GMD_ERR_BAD_ARG, "scan_tc: bad sizes");
    GMD_REQUIRE(!sweep || !perm || Nq == Nc, GMD_ERR_BAD_ARG, "scan_tc: one permutation serves queries and candidates");
    using namespace gmd_tc;
    const int nkb = tc_nkb(D);
    GMD_REQUIRE(D >= 1 && nkb <= MAX_KB && kk >= 1 && kk <= 96, GMD_ERR_UNSUPPORTED, "scan_tc: D=%d kk=%d", D, kk);
    const size_t fixed = 1024 + (size_t)tile_bytes(nkb) + 512 + (size_t)TQ * TM * 8 + (size_t)TM * (kk | 1) * 8;
    GMD_REQUIRE(fixed + 2 * (size_t)KB_B <= GMD_SMEM_LIMIT, GMD_ERR_UNSUPPORTED,
                "scan_tc: D=%d kk=%d needs more than 227 KB of shared memory", D, kk);
    const int ns = (int)std::min<size_t>(std::min(NS_MAX, NACC * nkb), (GMD_SMEM_LIMIT - fixed) / KB_B);
    const size_t smem = fixed + (size_t)ns * KB_B;
    const unsigned grid = (unsigned)((Nq + TM - 1) / TM);
#ifdef GMD_TC_PROBES
    const int dbg = getenv("GMD_TC_DBG") ? atoi(getenv("GMD_TC_DBG")) : 0;  // developer probe (tools/tc_time.py)
#else
    const int dbg = 0;
#endif
    const int ntiles_c = (int)((Nc + TM - 1) / TM);
    auto launch = [&](auto kern) -> int {
        GMD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>((const unsigned char*)qtiles, (const unsigned char*)ctiles,
                                                           Nq, Nc, ntiles_c, kk, ns, perm, c_index_offset, dbg,
                                                           stats_c, stats_q, cand_idx, cand_val, cand_tau);
        return GMD_OK;
    };
    int rc;
    switch (nkb * 2 + (sweep ? 1 : 0)) {
        case 2: rc = launch(scan_tc_kernel<1, false>); break;
        case 3: rc = launch(scan_tc_kernel<1, true>); break;
        case 4: rc = launch(scan_tc_kernel<2, false>); break;
        case 5: rc = launch(scan_tc_kernel<2, true>); break;
        case 6: rc = launch(scan_tc_kernel<3, false>); break;
        case 7: rc = launch(scan_tc_kernel<3, true>); break;
        case 8: rc = launch(scan_tc_kernel<4, false>); break;
        case 9: rc = launch(scan_tc_kernel<4, true>); break;
        case 10: rc = launch(scan_tc_kernel<5, false>); break;
        default: rc = launch(scan_tc_kernel<5, true>); break;
    }
    if (rc != GMD_OK) return rc;
    GMD_CUDA_CHECK(cudaGetLastError());
    return GMD_OK;
}

extern "C" int gmd_scan_tc_f32(const void* qtiles, int64_t Nq, const void* ctiles, int64_t Nc, int D, int kk,
                               int64_t c_index_offset, const float* stats_c, const float* stats_q, uint32_t* cand_idx,
                               float* cand_val, float* cand_tau, void* stream) {
    return gmd_scan_tc_sweep_f32(qtiles, Nq, ctiles, Nc, D, kk, c_index_offset, stats_c, stats_q, nullptr, 0, cand_idx,
                                 cand_val, cand_tau, stream);
}
