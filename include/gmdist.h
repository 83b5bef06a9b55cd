/*
 * gmdist.h -- C ABI of libgmdist.so, the B200 (sm_100a) contig-profile distance
 * and neighbour-search engine.
 *
 * The reference (GroopM, /root/reference/groopm) has no FFI of its own: its hot
 * path is ~20 inline scipy/numpy calls.  Each entry point below names the
 * reference call site (file:line) whose arithmetic it replaces; INTEGRATION.md
 * shows the ctypes binding a GroopM maintainer would add at that site.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - every function returns a gmd_status (0 == OK, < 0 == error);
 *     gmd_last_error() returns a per-thread message for the last failure.
 *   - "*_host_*" functions take HOST pointers in the reference's own dtype
 *     (float64, C-contiguous, row-major [rows, D]) and do allocation, H2D, the
 *     kernels and D2H themselves on the current device -- the drop-in layer.
 *   - all other functions take DEVICE pointers, never allocate or free device
 *     memory, are stream-ordered and asynchronous (stream is a cudaStream_t
 *     passed as void*; NULL == the legacy default stream).
 *   - metric: 0 = Euclidean (scipy 'euclidean'), 1 = cityblock (scipy 'cityblock').
 *   - results are exact in the reference's float64 sense: the FP32 scan only
 *     proposes candidates; a float64 re-rank from the original float64 rows
 *     emits the distances and a per-row certificate sends any row whose FP32
 *     candidate set could be incomplete to an exact float64 scan.  Ties are
 *     broken by lowest candidate index (np.argmin / first-occurrence rule,
 *     torusMesh.py:127).
 */
#ifndef GMDIST_H
#define GMDIST_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define GMD_API __attribute__((visibility("default")))
#else
#define GMD_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    GMD_OK = 0,
    GMD_ERR_BAD_ARG = -1,      /* shapes / k / metric out of range  -> ValueError      */
    GMD_ERR_NONFINITE = -2,    /* NaN/Inf in an input (np.seterr(all='raise'),
                                  torusMesh.py:57)                  -> FloatingPointError */
    GMD_ERR_CUDA = -3,         /* CUDA runtime error                 -> RuntimeError    */
    GMD_ERR_WORKSPACE = -4,    /* caller workspace too small         -> ValueError      */
    GMD_ERR_UNSUPPORTED = -5   /* D or k beyond what the kernels are built for          */
} gmd_status;

enum { GMD_L2 = 0, GMD_L1 = 1 };

/* FP32 scan forms (chosen by gmd_pick_form unless forced):
 *   0 expanded  ||c||^2 - 2 q.c   (L2 only; D FFMA per pair)
 *   1 direct L2 sum (q-c)^2       (2D instr per pair; for ill-conditioned / tiny D)
 *   2 direct L1 sum |q-c|         (2D instr per pair)
 *   3 tensor cores: ||c||^2 - 2 q.c with q.c as a 3 x FP16 (hi/lo split) tcgen05 contraction
 *     (L2 only, D <= 32; its own tile format, see gmd_scan_tc_f32)                    */
enum { GMD_FORM_EXPANDED = 0, GMD_FORM_DIRECT_L2 = 1, GMD_FORM_DIRECT_L1 = 2, GMD_FORM_TC = 3, GMD_FORM_AUTO = -1 };

#define GMD_TILE_ROWS 64        /* profile rows per HBM tile                        */
#define GMD_MAX_K 64            /* largest k of the fused top-k                     */
#define GMD_MAX_D 136           /* largest profile dimension (raw tetranucleotides) */

GMD_API const char* gmd_last_error(void);
GMD_API int gmd_version(void);

/* ------------------------------------------------------------------ host-level drop-ins
 * (inputs/outputs are HOST float64/int32 arrays; each call is synchronous)            */

/* Best-matching unit of every row of X on a flattened torus mesh.
 * Replaces the per-contig  argmin(cdist(flatNodes, [v]))  of TorusMesh.bestMatch
 * (torusMesh.py:123-130), the classify loops refine.py:1219-1222 / 2137-2138 and the
 * inner BMU of SOM.train (som.py:348-353).  out_loc[i] is the flat node index
 * (first minimum); out_dist (nullable) the float64 Euclidean distance.               */
GMD_API int gmd_bmu_host_f64(const double* nodes, int64_t M, const double* X, int64_t N, int D,
                     int32_t* out_loc, double* out_dist);

/* k nearest candidates of every query row, ascending (distance, index), self included
 * when Q and C are the same rows.  Replaces KDTree(data).query(x, k)
 * (refine.py:653-665, metric 0) and sort(squareform(pdist(X,'cityblock')))[:, :k]
 * (cluster.py:495-498, metric 1) without materialising the n x n matrix.
 * out_idx [Nq,k] int32, out_dist [Nq,k] float64.                                      */
GMD_API int gmd_knn_host_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D,
                     int k, int metric, int32_t* out_idx, double* out_dist);

/* ------------------------------------------------------------------ device-level stages
 * (device pointers; stream-ordered; caller owns all memory)                           */

/* bytes of the FP32 tile array for N rows of dimension D */
GMD_API size_t gmd_tiles_bytes(int64_t N, int D);

/* column means of X[N,D] (float64) -> center[D]; the centring that keeps FP32
 * norm expansion well conditioned (SURVEY.md section 7 "hard parts").                 */
GMD_API int gmd_col_mean_f64(const double* X, int64_t N, int D, double* center, void* stream);

/* float64 row-major rows -> centred FP32 tiles [tile][D+1][64] (row D: ||x||^2 for
 * form 0).  stats[0] = max ||x||_2^2, stats[1] = max ||x||_1 over the rows (FP32 bit
 * patterns, updated with atomicMax: zero it once per candidate set);
 * flags[0] is set to 1 if any input is NaN/Inf.                                       */
GMD_API int gmd_prep_f64(const double* X, int64_t N, int D, const double* center, int form,
                 float* tiles, float* stats, int32_t* flags, void* stream);

/* The hot kernel: FP32 distance tiles + fused warp-level top-kk selection; the
 * Nq x Nc matrix is never materialised.  Candidates are cut into nsplit contiguous
 * ranges; for each (query, split) the kk best (value, index) pairs are written, in no
 * particular order, to cand_idx / cand_val [Nq][nsplit][kk] (index 0xFFFFFFFF == empty
 * slot) and cand_tau [Nq][nsplit] receives the FP32 value that every candidate of the
 * split NOT in the list is known to reach or exceed (+inf if all were listed).
 * c_index_offset is added to every candidate index (row-sharded candidates).          */
GMD_API int gmd_scan_topk_f32(const float* qtiles, int64_t Nq, const float* ctiles, int64_t Nc,
                      int D, int form, int kk, int nsplit, int64_t c_index_offset,
                      uint32_t* cand_idx, float* cand_val, float* cand_tau, void* stream);

/* The k = 1 case (best-matching unit: argmin(cdist(flatNodes, [v])), torusMesh.py:127,
 * som.py:351) with a short profile: same tiles in, no heaps -- every (lane, row) tracks its
 * best 8-candidate cell branch-free; the winning cell is proposed with kk = gmd_bmu_kk() = 8,
 * nsplit = 1, and cand_tau bounds all other candidates, so gmd_refine_topk_f64 /
 * gmd_exact_rows_f64 finish it exactly as for the general scan.  Valid when
 * gmd_use_bmu_scan(k, metric, D) is non-zero (k = 1, Euclidean, expanded form, D <= 16). */
GMD_API int gmd_use_bmu_scan(int k, int metric, int D);
GMD_API int gmd_bmu_kk(void);
GMD_API int gmd_scan_bmu_f32(const float* qtiles, int64_t Nq, const float* ctiles, int64_t Nc, int D,
                     int64_t c_index_offset, uint32_t* cand_idx, float* cand_val, float* cand_tau,
                     void* stream);

/* Tensor-core scan (form GMD_FORM_TC): the dot products of a 128-query x 128-candidate tile
 * are one tcgen05.mma kind::f16 contraction (FP32 accumulation in tensor memory), every centred
 * FP32 value -- scaled by a power of two that brings the largest row norm into (128, 256] --
 * split hi + lo (both FP16) and the contraction issued three times (hi.hi + hi.lo + lo.hi:
 * FP32-class accuracy, measured 2^-21.9 |q||c|); the norm rides in the contraction (column D);
 * selection threads read the accumulators from tensor memory and keep the per-query heaps.
 * Operands are "TC tiles" (128 profiles: hi image, lo image, norms), staged per ROLE
 * (0 queries, 1 candidates) in two passes: gmd_tc_norms_f64 on each set (stats[0] <- max
 * ||x - center||^2, zero the slot first; flags[0] <- 1 on NaN/Inf), then gmd_prep_tc_f64 on each
 * set with BOTH sets' stats slots (the same slot twice when queries == candidates).
 * Outputs as gmd_scan_topk_f32 with nsplit = 1, in the caller's units, to be finished by
 * gmd_refine_topk_f64(form = GMD_FORM_TC, stats = candidate slot, stats + 2 = query slot) /
 * gmd_exact_rows_f64.  A tile holds ceil((D + 1) / 32) K blocks of 32 columns, streamed through the
 * ring one block at a time and accumulated in tensor memory.
 * gmd_tc_supported(k, metric, D): Euclidean, 2 <= k <= 64, 4 <= D <= 136.                  */
GMD_API int gmd_tc_supported(int k, int metric, int D);
GMD_API int gmd_use_tc_scan(int64_t Nq, int64_t Nc, int k, int metric, int D);  /* the size policy of gmd_knn_f64 */
GMD_API size_t gmd_tc_tiles_bytes(int64_t N, int D);
GMD_API int gmd_tc_norms_f64(const double* X, int64_t N, int D, const double* center, float* stats, int32_t* flags,
                     void* stream);
GMD_API int gmd_prep_tc_f64(const double* X, int64_t N, int D, const double* center, int role,
                    const float* stats_c, const float* stats_q, void* tiles, void* stream);
GMD_API int gmd_scan_tc_f32(const void* qtiles, int64_t Nq, const void* ctiles, int64_t Nc, int D, int kk,
                    int64_t c_index_offset, const float* stats_c, const float* stats_q, uint32_t* cand_idx,
                    float* cand_val, float* cand_tau, void* stream);

/* Locality-ordered sweep of the tensor-core scan (all-pairs: queries == candidates).
 * gmd_order_by_key: counting sort of the rows by a small integer key (0 <= key < nkeys; the
 * engine uses the nearest of ~N/1024 sampled profiles): perm[position] = row.
 * gmd_prep_tc_perm_f64 stages tiles in that order (perm NULL: as gmd_prep_tc_f64).
 * gmd_scan_tc_sweep_f32 with sweep != 0: the CTA of query tile t visits the candidate tiles
 * t, t+1, ... wrapping around, so a query meets its neighbourhood first and few later
 * candidates beat its threshold; proposals (cand_idx rows and values) are written under the
 * ORIGINAL row numbers (perm NULL: positions are row numbers).  Everything downstream
 * (gmd_refine_topk_f64 / gmd_exact_rows_f64) is unchanged, and so are the results.        */
GMD_API size_t gmd_order_workspace_bytes(int nkeys);
GMD_API int gmd_order_by_key(const int32_t* key, int64_t N, int nkeys, uint32_t* perm, void* workspace,
                     size_t workspace_bytes, void* stream);
GMD_API int gmd_prep_tc_perm_f64(const double* X, int64_t N, int D, const double* center, int role,
                     const float* stats_c, const float* stats_q, const uint32_t* perm, void* tiles,
                     void* stream);
GMD_API int gmd_scan_tc_sweep_f32(const void* qtiles, int64_t Nq, const void* ctiles, int64_t Nc, int D, int kk,
                     int64_t c_index_offset, const float* stats_c, const float* stats_q,
                     const uint32_t* perm, int sweep, uint32_t* cand_idx, float* cand_val,
                     float* cand_tau, void* stream);

/* Pruned tensor-core scan (the default for large Euclidean problems; grouped.cu, order.cu,
 * scan_tc_prune.cu).  The reference has no counterpart: KDTree.query (refine.py:653-665) is its only
 * space-partitioning search; this is the same idea -- exact neighbours without looking at every pair --
 * for the all-pairs / many-queries shapes.
 *   Grouped order.  gmd_group_leaders_f64 picks "leaders" from a strided sample of ns rows by leader
 * clustering (a sample becomes a leader unless an earlier leader lies within rfac x the median
 * nearest-sample distance; at most cap); SYNCHRONISES the stream to return their number.  Outputs:
 * L [cap][D] centred FP32 leader rows, lead_row [cap] their row numbers,
 * lead_rows [cap][D] float64 copies (the candidate set of the nearest-leader search, which is this
 * library's own k = 1 pipeline), info [2] = {leaders, float bits of the radius}.
 * gmd_group_order: given every row's nearest leader and float64 distance to it, keys the rows by
 * (leader, distance shell of width radius / 16; gmd_group_shells() shells), counting-sorts them with every
 * leader's group padded to whole `tile`-row tiles:  perm[position] = row, 0xFFFFFFFF = padding
 * (npos >= gmd_group_positions(N, ngroups, tile) positions), gstart [ngroups + 1] first tile of a group,
 * tgroup [npos / tile] group of a tile (-1: padding only), grad [ngroups] (nullable) largest member
 * distance of a group, rounded up; key [N] scratch.
 *   gmd_prep_tc_grouped_f64 stages TC tiles through such a perm and records a ball per tile in
 * tmeta [D + 1][nt_alloc] (scaled units: centroid, radius rounded up, -1 for a tile of padding).
 *   gmd_tc_stage_grouped_f64 does all of it for a query and a candidate set (Q == C: all-pairs; Q a run
 * of C's rows: the row-sharded path, keys are shared; otherwise Q is keyed by its own search) and fills
 * the descriptor the scan takes; workspace from gmd_tc_grouped_workspace_bytes, tile buffers sized by
 * gmd_tc_grouped_tiles_bytes(N rows, Nc, D).  stats_c / stats_q as for gmd_prep_tc_f64.
 *   gmd_scan_tc_grouped_f32: the tcgen05 scan of gmd_scan_tc_f32, except that the CTA of a query tile
 * first visits a window of its own group, takes T = its queries' largest k-th distance so far, and then
 * loads only the candidate tiles whose ball bound |m_q - m_c| - r_q - r_c does not exceed T (bounds
 * are evaluated with every rounding taken against them).  With one K block (D <= 31) a CTA takes two
 * query tiles at once (twice the selection warps per SM: the selection's dependent shared-memory round
 * trips are what bounds a scan that visits few tiles).  Outputs and the meaning of cand_tau are unchanged
 * (nsplit = 1), so gmd_refine_topk_f64 / gmd_exact_rows_f64 finish it; results are identical to the full
 * scan.  counters (device, in the descriptor): [0] tiles visited, [1] query tiles, [2] groups surviving. */
typedef struct gmd_grouped {
    int64_t npos_q, npos_c;               /* staged positions (rows + padding) */
    int32_t ngroups, ng_alloc, nt_alloc_q, nt_alloc_c, same, reserved;
    const uint32_t* perm_q;
    const uint32_t* perm_c;
    const int32_t* tgroup_q;
    const int32_t* gstart_c;
    const float* gmeta;                   /* [D + 1][ng_alloc] leaders + group radii, scaled units */
    const float* tmeta_q;                 /* [D + 1][nt_alloc_q] */
    const float* tmeta_c;                 /* [D + 1][nt_alloc_c] */
    uint64_t* counters;                   /* [8]: tiles visited, query tiles, groups surviving, -, then SM cycles of
                                             thread 0 summed over the CTAs: set-up, pipeline rounds, bounds, write-out */
} gmd_grouped;
GMD_API int gmd_group_shells(void);
GMD_API void gmd_group_params(int64_t N, int* ns, int* cap);
GMD_API int64_t gmd_group_positions(int64_t N, int ngroups, int tile);
GMD_API size_t gmd_group_leaders_workspace_bytes(int ns, int D);
GMD_API int gmd_group_leaders_f64(const double* X, int64_t N, int D, const double* center, int ns, int cap, float rfac,
                     float* L, int32_t* lead_row, double* lead_rows, int32_t* info, void* workspace,
                     size_t workspace_bytes, int32_t* nlead_host, void* stream);
GMD_API size_t gmd_group_order_workspace_bytes(int ngroups);
GMD_API int gmd_group_order(const int32_t* lead, const double* dist, int64_t N, int ngroups, const int32_t* info,
                     int tile, int32_t* key, float* grad, uint32_t* perm, int64_t npos, int32_t* gstart,
                     int32_t* tgroup, void* workspace, size_t workspace_bytes, void* stream);
/* the same with every group padded to a multiple of `pad` tiles (gstart / tgroup stay in units of one tile; positions:
 * gmd_group_positions(N, ngroups, tile * pad)): for scans whose CTA owns `pad` neighbouring query tiles */
GMD_API int gmd_group_order_pad(const int32_t* lead, const double* dist, int64_t N, int ngroups, const int32_t* info,
                     int tile, int pad, int32_t* key, float* grad, uint32_t* perm, int64_t npos, int32_t* gstart,
                     int32_t* tgroup, void* workspace, size_t workspace_bytes, void* stream);
GMD_API int gmd_prep_tc_grouped_f64(const double* X, int64_t Npos, int D, const double* center, int role,
                     const float* stats_c, const float* stats_q, const uint32_t* perm, void* tiles,
                     float* tmeta, int nt_alloc, void* stream);
/* one pass for profiles that serve as queries AND candidates (same rows, same order): both roles' tiles, one set of balls */
GMD_API int gmd_prep_tc_grouped_both_f64(const double* X, int64_t Npos, int D, const double* center,
                     const float* stats_c, const float* stats_q, const uint32_t* perm, void* qtiles, void* ctiles,
                     float* tmeta, int nt_alloc, void* stream);
GMD_API int gmd_tc_group_meta(const float* L, const float* grad, int ngroups, int D, const float* stats_c,
                     const float* stats_q, float* gmeta, int ng_alloc, void* stream);
GMD_API int gmd_use_tc_prune(int64_t Nq, int64_t Nc, int k, int metric, int D);   /* the size policy of gmd_knn_f64 */
GMD_API size_t gmd_tc_prune_max_tiles(int D, int kk);
GMD_API size_t gmd_tc_grouped_tiles_bytes(int64_t N, int64_t Nc, int D);
GMD_API size_t gmd_tc_grouped_workspace_bytes(int64_t Nq, int64_t Nc, int D);
GMD_API int gmd_tc_stage_grouped_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D,
                     const double* center, const float* stats_c, const float* stats_q, void* qtiles, void* ctiles,
                     void* workspace, size_t workspace_bytes, gmd_grouped* out, void* stream);
/* The staging in two halves, for row-sharded callers (queries = a run of the candidates' rows on every
 * GPU): gmd_tc_grouped_leaders_f64 picks the leaders into the workspace (same result on every rank) and
 * returns their number and device rows; the caller finds the nearest leader of ITS rows, all-gathers
 * (index, float64 distance) over the ranks and hands the full arrays to gmd_tc_stage_grouped_ext_f64, which
 * does the rest.  Same workspace (gmd_tc_grouped_workspace_bytes) for both calls.                    */
GMD_API int gmd_tc_grouped_leaders_f64(const double* C, int64_t Nq, int64_t Nc, int D, const double* center,
                     void* workspace, size_t workspace_bytes, int32_t* nlead_host, const double** lead_rows_dev,
                     void* stream);
GMD_API int gmd_tc_stage_grouped_ext_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D,
                     const double* center, const float* stats_c, const float* stats_q, void* qtiles, void* ctiles,
                     void* workspace, size_t workspace_bytes, int32_t nlead, const int32_t* nearest_idx,
                     const double* nearest_dist, gmd_grouped* out, void* stream);
/* A leader for every row: best proposal of the FP32 best-matching-unit scan plus the exact float64 distance to it
 * (no re-rank or certificate: the grouped staging needs a good deterministic assignment and the true distance to the
 * ASSIGNED leader, not the proven nearest).  Workspace: gmd_knn_workspace_bytes(N, nlead, D, 1).            */
GMD_API int gmd_assign_leaders_f64(const double* X, int64_t N, const double* L, int nlead, int D, int32_t* out_idx,
                     double* out_dist, void* workspace, size_t workspace_bytes, void* stream);
GMD_API int gmd_pick_proposal_f64(const double* X, int64_t N, const double* C, int D, int m, const uint32_t* cand_idx,
                     const float* cand_val, int32_t* out_idx, double* out_dist, void* stream);
/* Re-stages only the QUERY side of a staging whose queries were a run of the candidates' rows, for another run of
 * the same length (Q = C + r0 * D): the candidates' order, tiles and balls stay.  For callers that walk the rows
 * in runs so that the copy-out of one run overlaps the scan of the next.                                     */
GMD_API int gmd_tc_restage_queries_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D,
                     const double* center, const float* stats_c, const float* stats_q, void* qtiles,
                     void* workspace, size_t workspace_bytes, gmd_grouped* g, void* stream);
GMD_API int gmd_scan_tc_prune_f32(const void* qtiles, int64_t Nq_pos, const void* ctiles, int64_t Nc_pos, int D, int kk,
                     int64_t c_index_offset, const float* stats_c, const float* stats_q,
                     const uint32_t* perm_q, const uint32_t* perm_c, const int32_t* tgroup_q,
                     const int32_t* gstart_c, int ngroups, const float* gmeta, int ng_alloc,
                     const float* tmeta_q, int nt_alloc_q, const float* tmeta_c, int nt_alloc_c, int same,
                     uint32_t* cand_idx, float* cand_val, float* cand_tau, uint64_t* counters, void* stream);
GMD_API int gmd_scan_tc_grouped_f32(const void* qtiles, const void* ctiles, int D, int kk, int64_t c_index_offset,
                     const float* stats_c, const float* stats_q, const gmd_grouped* g,
                     uint32_t* cand_idx, float* cand_val, float* cand_tau, void* stream);

/* float64 re-rank of the proposed candidates from the ORIGINAL rows, lowest-index tie
 * rule, plus the per-row certificate; uncertified rows are appended to fail_rows
 * (count in fail_count[0]) and finished by gmd_exact_rows_f64.                        */
GMD_API int gmd_refine_topk_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D,
                        const double* center, int k, int kk, int nsplit, int metric,
                        int form, const uint32_t* cand_idx, const float* cand_val,
                        const float* cand_tau, const float* stats, int32_t* out_idx,
                        double* out_dist,
                        int32_t* fail_count, int32_t* fail_rows, void* stream);
/* The same with the queries visited in a staging order (qorder[position] = query row, 0xFFFFFFFF =
 * padding, nqpos positions; e.g. gmd_grouped.perm_q): consecutive blocks then re-rank against the same
 * neighbourhood of candidate rows and find them in L2.  Results are identical.               */
GMD_API int gmd_refine_topk_order_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D,
                     const double* center, int k, int kk, int nsplit, int metric, int form,
                     const uint32_t* cand_idx, const float* cand_val, const float* cand_tau,
                     const float* stats, int32_t* out_idx, double* out_dist, int32_t* fail_count,
                     int32_t* fail_rows, const uint32_t* qorder, int64_t nqpos, void* stream);

/* exact float64 scan of all candidates for the listed rows (reads the count on device).
 * With a workspace of gmd_exact_workspace_bytes(Nq) the listed rows (up to 4096) share one
 * pass over the candidates: each block stages 128 candidates once and every listed row
 * visits them, keeping what lies within its re-ranked k-th distance; rows with more than
 * 256 such candidates, rows beyond the capacity, and all rows when workspace is NULL are
 * scanned one row per block.                                                              */
GMD_API size_t gmd_exact_workspace_bytes(int64_t Nq);
GMD_API int gmd_exact_rows_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D,
                       int k, int metric, const int32_t* fail_count,
                       const int32_t* fail_rows, int32_t* out_idx, double* out_dist,
                       void* workspace, size_t workspace_bytes, void* stream);

/* whole pipeline on device pointers with a caller workspace (see
 * gmd_knn_workspace_bytes); form = GMD_FORM_AUTO picks by metric and D.
 * info[0] (nullable, device int32) receives the number of rows that needed the exact
 * float64 scan.                                                                       */
GMD_API size_t gmd_knn_workspace_bytes(int64_t Nq, int64_t Nc, int D, int k);
GMD_API int gmd_knn_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D, int k,
                int metric, int form, int32_t* out_idx, double* out_dist,
                void* workspace, size_t workspace_bytes, int32_t* info, void* stream);

/* ------------------------------------------------------------------ float64 pairwise stages
 * (device pointers; values bit-identical to scipy's cdist/pdist: direct differences,
 * sequential float64 accumulation)                                                    */

/* cdist(XA, XB, metric) -> out [mA, mB] row-major (torusMesh.py:127, som.py:351,
 * refine.py:927, 945, 1006, 1503; mstore.py:1915).                                    */
GMD_API int gmd_cdist_f64(const double* A, int64_t mA, const double* B, int64_t mB, int D, int metric,
                  double* out, void* stream);

/* Angles between coverage profiles for every pair of a small block (RefineEngine.cDist /
 * cDistBetweenBins, refine.py:1528-1583):
 *     arccos(dot(A[i], B[j]) / (normA[i] * normB[j]))  -> out [mA, mB]
 * normA / normB are the caller's normCoverages of the rows (refine.py:1546, 1578 divide by
 * self.PM.normCoverages, loaded at profileManager.py:205, not by norms recomputed from the rows);
 * null = the rows' own Euclidean norms.  Pairs the reference skips (FloatingPointError: argument
 * outside [-1, 1], or a zero norm) are NaN.                                              */
GMD_API int gmd_angle_block_f64(const double* A, int64_t mA, const double* B, int64_t mB, int D,
                        const double* normA, const double* normB, double* out, void* stream);

/* pdist(X, metric) -> condensed out [n(n-1)/2], scipy order (cluster.py:495, 497;
 * refine.py:1492; index map of mstore.py:1880-1888).                                  */
GMD_API int gmd_pdist_f64(const double* X, int64_t n, int D, int metric, double* out, void* stream);

/* where(D[i, :] <= radius)[0] for every row i without the n x n matrix
 * (cluster.py:533-536, 570-573): count pass -> caller prefix-sums into indptr [n+1] ->
 * fill pass writes ascending candidate indices (and, if dists != NULL, distances).    */
GMD_API int gmd_radius_count_f64(const double* X, int64_t n, int D, int metric, double radius,
                         int64_t* counts, void* stream);
GMD_API int gmd_radius_fill_f64(const double* X, int64_t n, int D, int metric, double radius,
                        const int64_t* indptr, int32_t* indices, double* dists, void* stream);

/* distances of the CSR pairs (i, indices[p]) in ANOTHER profile space: the
 * k_dist_matrix[index][neighbours] / c_dist_matrix[index][neighbours] gathers of
 * cluster.py:547, 583.                                                                */
GMD_API int gmd_gather_dist_f64(const double* X, int64_t n, int D, int metric, const int64_t* indptr,
                        const int32_t* indices, double* out, void* stream);

/* np.median of a never-materialised distance block by radix select:
 * B != NULL: median(cdist(A, B, metric))  (kDistBetweenBins, refine.py:1502-1503);
 * B == NULL: median(pdist(A, metric))     (kDist, refine.py:1490-1496).
 * out_median: device double[1]; workspace: gmd_median_workspace_bytes() device bytes. */
GMD_API size_t gmd_median_workspace_bytes(void);
GMD_API int gmd_median_dist_f64(const double* A, int64_t mA, const double* B, int64_t mB, int D, int metric,
                        double* out_median, void* workspace, void* stream);

/* Many medians in ONE launch (one block per job, the radix select entirely inside the block):
 * the per-bin loops of getKCut (refine.py:1479-1488 -> kDist) and of findMergeGroups /
 * combineMergers (refine.py:718, 876 -> kDistBetweenBins).  X [N, D] holds all profiles;
 * rows is a concatenation of int32 row-index lists; job j is the int32 quadruple
 * jobs[4j..4j+3] = (a_off, a_len, b_off, b_len) into rows, b_len < 0 meaning
 * median(pdist(X[a])) and otherwise median(cdist(X[a], X[b])).  out[j] is NaN when job j has
 * no pairs.  Meant for bins up to a few thousand rows; larger blocks go through
 * gmd_median_dist_f64.                                                                   */
GMD_API int gmd_median_batch_f64(const double* X, int64_t N, int D, int metric, const int32_t* rows,
                         const int32_t* jobs, int n_jobs, double* out, void* stream);

/* Host-layer helpers: flag[0] <- 1 if any of the n values is NaN / Inf (the reference runs under
 * np.seterr(all='raise'), torusMesh.py:57: a non-finite profile is a FloatingPointError there);
 * indptr[0] = 0, indptr[i + 1] = counts[0] + ... + counts[i] (the CSR row pointer between
 * gmd_radius_count_f64 and gmd_radius_fill_f64).                                        */
GMD_API int gmd_nonfinite_f64(const double* X, int64_t n, int32_t* flag, void* stream);
GMD_API int gmd_zero_async(void* p, size_t bytes, void* stream);
/* argmin(cdist(nodes, [X_i])) for a FEW vectors (N <= 65535), straight in float64 with scipy's arithmetic,
 * one block per vector, first minimum: the per-contig bestMatch call (torusMesh.py:123-130) and the BMUs
 * of one SOM.train iteration (som.py:348-353) without tile staging -- one launch.  flag (nullable) <- 1 if
 * a distance is NaN / Inf.                                                                */
GMD_API int gmd_bmu_small_f64(const double* nodes, int64_t M, const double* X, int64_t N, int D, int32_t* out_idx,
                      double* out_dist, int32_t* flag, void* stream);
/* The k nearest candidates of a FEW query rows against at most gmd_knn_small_max_candidates() rows, straight
 * in float64 (one block per query, distances in shared memory, k rounds of selection): the per-bin
 * KDTree(data).query(x, k=6) of findMergeGroups (refine.py:653-665) in one launch.  Ascending (distance,
 * index); out_idx [Nq, k], out_dist [Nq, k].                                               */
GMD_API int gmd_knn_small_max_candidates(void);
GMD_API int gmd_knn_small_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D, int k, int metric,
                      int32_t* out_idx, double* out_dist, int32_t* flag, void* stream);
GMD_API int gmd_indptr_from_counts_i64(const int64_t* counts, int64_t n, int64_t* indptr, void* stream);

/* One iteration of SOM.train's weight update (som.py:328-384) for the T training vectors
 * X[index[t]] whose best-matching units on the frozen weights are loc[t] (flat node index,
 * row = loc / cols): every node gathers  -(w - x) * stamp  over the 3 x 3 torus tiling in the
 * reference's summation order and Wout = clip(W + folded delta, 0, 1); nodes with mask[node]
 * == 0 keep their weight (mask NULL: all nodes move).  stamp is the culled
 * (2*stamp_radius+1)^2 influence array of the iteration.  Bit-identical to the reference's
 * float64 result, deterministic.  W and Wout may be the same buffer.                      */
GMD_API int gmd_som_update_f64(const double* W, double* Wout, int rows, int cols, int D, const double* X,
                       int64_t n_train, const int32_t* index, const int32_t* loc, int T,
                       const double* stamp, int stamp_radius, const unsigned char* mask, void* stream);

/* split / slack policy shared with the host layer */
GMD_API int gmd_pick_form(int metric, int D);
GMD_API int gmd_pick_kk(int k);
GMD_API int gmd_pick_nsplit(int64_t Nq, int64_t Nc, int D, int kk);

/* FP32 FFMA peak probe (roofline denominator): runs
 * iters x 4096 dependent-free FFMAs per thread on every SM and returns the flop
 * count in *flops; time it with events around the call.                               */
GMD_API int gmd_fp32_peak_probe(int iters, float* sink, double* flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GMDIST_H */
