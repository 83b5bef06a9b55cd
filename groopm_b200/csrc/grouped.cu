// Host side of the pruned tensor-core scan: stages both profile sets in the grouped order of order.cu
// (leaders from a sample of the candidates, every row keyed by its nearest leader with this library's own
// k = 1 pipeline, padded counting sort), writes the TC tiles with their balls, and hands the scan a
// descriptor of everything it needs (gmd_grouped).  One call, one host synchronisation (the number of
// leaders decides the shape of the nearest-leader search).
#include <algorithm>

#include "tc_common.cuh"

namespace {

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Carve {
    char* p;
    size_t off = 0;
    explicit Carve(void* base) : p((char*)base) {}
    template <class T>
    T* take(size_t n) {
        T* r = p ? (T*)(p + off) : nullptr;
        off += align_up(n * sizeof(T));
        return r;
    }
};

struct Layout {
    // leaders
    void* lead_ws;
    size_t lead_ws_bytes;
    float* Lt;
    int32_t* lead_row;
    double* lead_rows;
    int32_t* info;
    float* grad;
    // nearest-leader search
    int32_t* idx_c;
    double* dist_c;
    int32_t* idx_q;
    double* dist_q;
    void* knn_ws;
    size_t knn_ws_bytes;
    // order
    int32_t* key_c;
    int32_t* key_q;
    uint32_t* perm_c;
    uint32_t* perm_q;
    int32_t* gstart_c;
    int32_t* gstart_q;
    int32_t* tgroup_c;
    int32_t* tgroup_q;
    void* order_ws;
    size_t order_ws_bytes;
    // balls
    float* tmeta_c;
    float* tmeta_q;
    float* gmeta;
    uint64_t* counters;
    int ns, cap;
    int64_t cap_pos_c, cap_pos_q;
    size_t bytes;
};

// relation 0: unrelated sets, 1: queries are a slice of the candidates, 2: the same rows
Layout carve(void* base, int64_t Nq, int64_t Nc, int D, int relation) {
    Layout L;
    Carve c(base);
    gmd_group_params(Nc, &L.ns, &L.cap);
    const int TM = gmd_tc::TM;
    L.cap_pos_c = gmd_group_positions(Nc, L.cap, TM);
    L.cap_pos_q = gmd_group_positions(Nq, L.cap, 2 * TM);   // (query groups may be padded to pairs of tiles)
    L.lead_ws_bytes = gmd_group_leaders_workspace_bytes(L.ns, D);
    L.lead_ws = c.take<char>(L.lead_ws_bytes);
    L.Lt = c.take<float>((size_t)D * L.cap);
    L.lead_row = c.take<int32_t>(L.cap);
    L.lead_rows = c.take<double>((size_t)L.cap * D);
    L.info = c.take<int32_t>(4);
    L.grad = c.take<float>(L.cap);
    L.idx_c = c.take<int32_t>(Nc);
    L.dist_c = c.take<double>(Nc);
    L.idx_q = relation == 0 ? c.take<int32_t>(Nq) : nullptr;
    L.dist_q = relation == 0 ? c.take<double>(Nq) : nullptr;
    L.knn_ws_bytes = gmd_knn_workspace_bytes(std::max(Nq, Nc), L.cap, D, 1);
    L.knn_ws = c.take<char>(L.knn_ws_bytes);
    L.key_c = c.take<int32_t>(Nc);
    L.key_q = relation == 2 ? nullptr : c.take<int32_t>(Nq);
    L.perm_c = c.take<uint32_t>(L.cap_pos_c);
    L.perm_q = relation == 2 ? nullptr : c.take<uint32_t>(L.cap_pos_q);
    L.gstart_c = c.take<int32_t>(L.cap + 1);
    L.gstart_q = relation == 2 ? nullptr : c.take<int32_t>(L.cap + 1);
    L.tgroup_c = c.take<int32_t>(L.cap_pos_c / TM);
    L.tgroup_q = relation == 2 ? nullptr : c.take<int32_t>(L.cap_pos_q / TM);
    L.order_ws_bytes = gmd_group_order_workspace_bytes(L.cap);
    L.order_ws = c.take<char>(L.order_ws_bytes);
    L.tmeta_c = c.take<float>((size_t)(D + 1) * (L.cap_pos_c / TM));
    L.tmeta_q = c.take<float>((size_t)(D + 1) * (L.cap_pos_q / TM));
    L.gmeta = c.take<float>((size_t)(D + 1) * L.cap);
    L.counters = c.take<uint64_t>(16);
    L.bytes = c.off;
    return L;
}

// tiles a CTA of the pruned scan owns (gmd_scan_tc_prune_f32: two with one K block)
int query_pad(int D) { return tc_nkb(D) == 1 ? 2 : 1; }

int relation_of(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D, int64_t* q_row0) {
    *q_row0 = -1;
    if (Q == C && Nq == Nc) {
        *q_row0 = 0;
        return 2;
    }
    // queries that are a contiguous run of the candidates' rows (the row-sharded path: q = c[r0:r1])
    if (Q >= C && Q + (size_t)Nq * D <= C + (size_t)Nc * D && (Q - C) % D == 0) {
        *q_row0 = (Q - C) / D;
        return 1;
    }
    return 0;
}

}  // namespace

// The pruned scan pays when both sets are large (its staging is a nearest-leader pass over every row)
// and the candidate bitmasks fit the CTA's shared memory.  GMD_TC_PRUNE=0 / 1 overrides (developer switch).
extern "C" int gmd_use_tc_prune(int64_t Nq, int64_t Nc, int k, int metric, int D) {
    if (!gmd_tc_supported(k, metric, D)) return 0;
    int ns, cap;
    gmd_group_params(Nc, &ns, &cap);
    const size_t tiles = (size_t)(gmd_group_positions(Nc, cap, gmd_tc::TM) / gmd_tc::TM);
    if (tiles > gmd_tc_prune_max_tiles(D, gmd_pick_kk(k))) return 0;
    if (const char* env = getenv("GMD_TC_PRUNE")) return atoi(env) != 0;
    return Nq >= 32768 && Nc >= 32768;
}

extern "C" size_t gmd_tc_grouped_tiles_bytes(int64_t N, int64_t Nc, int D) {
    int ns, cap;
    gmd_group_params(Nc, &ns, &cap);
    return gmd_tc_tiles_bytes(gmd_group_positions(N, cap, 2 * gmd_tc::TM), D);   // enough for pairs of tiles per group
}

extern "C" size_t gmd_tc_grouped_workspace_bytes(int64_t Nq, int64_t Nc, int D) {
    return carve(nullptr, Nq, Nc, D, 0).bytes;
}

// Row-sharded callers split the staging in two: gmd_tc_grouped_leaders_f64 picks the leaders (identical on
// every rank: the kernels are deterministic) and hands back their rows; the caller finds the nearest leader of
// ITS rows only (this library's k = 1 pipeline), all-gathers (index, distance) of all rows and passes them to
// gmd_tc_stage_grouped_f64 as nearest_idx / nearest_dist -- the nearest-leader pass over every candidate row is
// then the one stage of the staging that is not repeated on every GPU.
extern "C" int gmd_tc_grouped_leaders_f64(const double* C, int64_t Nq, int64_t Nc, int D, const double* center,
                                          void* workspace, size_t workspace_bytes, int32_t* nlead_host,
                                          const double** lead_rows_dev, void* stream) {
    GMD_REQUIRE(C && center && workspace && nlead_host && lead_rows_dev && Nq > 0 && Nc > 0 && D >= 1 && D <= GMD_MAX_D,
                GMD_ERR_BAD_ARG, "grouped_leaders: bad arguments");
    Layout L = carve(workspace, Nq, Nc, D, 1);
    GMD_REQUIRE(L.bytes <= workspace_bytes, GMD_ERR_WORKSPACE, "grouped_leaders: workspace %zu B < required %zu B",
                workspace_bytes, L.bytes);
    *lead_rows_dev = L.lead_rows;
    return gmd_group_leaders_f64(C, Nc, D, center, L.ns, L.cap, 2.0f, L.Lt, L.lead_row, L.lead_rows, L.info, L.lead_ws,
                                 L.lead_ws_bytes, nlead_host, stream);
}

static int stage_grouped(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D, const double* center,
                         const float* stats_c, const float* stats_q, void* qtiles, void* ctiles, void* workspace,
                         size_t workspace_bytes, int32_t nlead_given, const int32_t* nearest_idx,
                         const double* nearest_dist, gmd_grouped* out, void* stream) {
    GMD_REQUIRE(Q && C && center && stats_c && stats_q && qtiles && ctiles && workspace && out, GMD_ERR_BAD_ARG,
                "stage_grouped: null pointer");
    GMD_REQUIRE(Nq > 0 && Nc > 0 && D >= 1 && D <= GMD_MAX_D, GMD_ERR_BAD_ARG, "stage_grouped: bad sizes");
    int64_t q_row0;
    int relation = relation_of(Q, Nq, C, Nc, D, &q_row0);
    const bool external = nearest_idx != nullptr;
    // (the split staging lays the workspace out for "queries are a run of the candidates' rows")
    GMD_REQUIRE(!external || (relation == 1 && nearest_dist && nlead_given >= 1), GMD_ERR_BAD_ARG,
                "stage_grouped: an external nearest-leader result needs queries that are a run of the candidates' rows");
    Layout L = carve(workspace, Nq, Nc, D, relation);
    GMD_REQUIRE(L.bytes <= workspace_bytes, GMD_ERR_WORKSPACE, "stage_grouped: workspace %zu B < required %zu B",
                workspace_bytes, L.bytes);
    cudaStream_t st = (cudaStream_t)stream;
    const int TM = gmd_tc::TM;
    int rc;
    int32_t nlead = nlead_given;
    if (external) {
        GMD_REQUIRE(nlead <= L.cap, GMD_ERR_BAD_ARG, "stage_grouped: %d leaders exceed the capacity %d", nlead, L.cap);
        GMD_CUDA_CHECK(cudaMemcpyAsync(L.idx_c, nearest_idx, (size_t)Nc * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        GMD_CUDA_CHECK(cudaMemcpyAsync(L.dist_c, nearest_dist, (size_t)Nc * sizeof(double), cudaMemcpyDeviceToDevice, st));
    } else {
        if ((rc = gmd_group_leaders_f64(C, Nc, D, center, L.ns, L.cap, 2.0f, L.Lt, L.lead_row, L.lead_rows, L.info,
                                        L.lead_ws, L.lead_ws_bytes, &nlead, stream)) != GMD_OK)
            return rc;
        // a leader (+ exact float64 distance to it) for every candidate row, and for the queries if they
        // are rows of their own
        if ((rc = gmd_assign_leaders_f64(C, Nc, L.lead_rows, nlead, D, L.idx_c, L.dist_c, L.knn_ws, L.knn_ws_bytes,
                                         stream)) != GMD_OK)
            return rc;
    }
    const int64_t npos_c = gmd_group_positions(Nc, nlead, TM);
    if ((rc = gmd_group_order(L.idx_c, L.dist_c, Nc, nlead, L.info, TM, L.key_c, L.grad, L.perm_c, npos_c, L.gstart_c,
                              L.tgroup_c, L.order_ws, L.order_ws_bytes, stream)) != GMD_OK)
        return rc;
    int64_t npos_q = npos_c;
    const uint32_t* perm_q = L.perm_c;
    const int32_t* tgroup_q = L.tgroup_c;
    if (relation != 2) {
        const int32_t* idx_q = L.idx_c + q_row0;
        const double* dist_q = L.dist_c + q_row0;
        if (relation == 0) {
            if ((rc = gmd_assign_leaders_f64(Q, Nq, L.lead_rows, nlead, D, L.idx_q, L.dist_q, L.knn_ws, L.knn_ws_bytes,
                                             stream)) != GMD_OK)
                return rc;
            idx_q = L.idx_q;
            dist_q = L.dist_q;
        }
        // a query order of its own: groups padded to PAIRS of tiles where the scan gives a CTA two query tiles, so that
        // no CTA straddles two groups (and walks two windows) -- what happens to half of them once a group holds two
        // or three query tiles (row-sharded callers, results streamed in runs)
        const int pad_q = query_pad(D);
        npos_q = gmd_group_positions(Nq, nlead, TM * pad_q);
        if ((rc = gmd_group_order_pad(idx_q, dist_q, Nq, nlead, L.info, TM, pad_q, L.key_q, nullptr, L.perm_q, npos_q,
                                      L.gstart_q, L.tgroup_q, L.order_ws, L.order_ws_bytes, stream)) != GMD_OK)
            return rc;
        perm_q = L.perm_q;
        tgroup_q = L.tgroup_q;
    }
    const int nt_c = (int)(L.cap_pos_c / TM), nt_q = (int)(L.cap_pos_q / TM);
    const float* tmeta_q = L.tmeta_q;
    int nt_q_out = nt_q;
    if (relation == 2) {
        // the same rows in the same order on both sides: one pass writes both roles' tiles and one set of balls
        if ((rc = gmd_prep_tc_grouped_both_f64(C, npos_c, D, center, stats_c, stats_q, L.perm_c, qtiles, ctiles, L.tmeta_c,
                                               nt_c, stream)) != GMD_OK)
            return rc;
        tmeta_q = L.tmeta_c;
        nt_q_out = nt_c;
    } else {
        if ((rc = gmd_prep_tc_grouped_f64(C, npos_c, D, center, 1, stats_c, stats_q, L.perm_c, ctiles, L.tmeta_c, nt_c,
                                          stream)) != GMD_OK)
            return rc;
        if ((rc = gmd_prep_tc_grouped_f64(Q, npos_q, D, center, 0, stats_c, stats_q, perm_q, qtiles, L.tmeta_q, nt_q,
                                          stream)) != GMD_OK)
            return rc;
    }
    if ((rc = gmd_tc_group_meta(L.Lt, L.grad, nlead, D, stats_c, stats_q, L.gmeta, L.cap, stream)) != GMD_OK)
        return rc;
    GMD_CUDA_CHECK(cudaMemsetAsync(L.counters, 0, 16 * sizeof(uint64_t), st));
    out->npos_q = npos_q;
    out->npos_c = npos_c;
    out->ngroups = nlead;
    out->ng_alloc = L.cap;
    out->nt_alloc_q = nt_q_out;
    out->nt_alloc_c = nt_c;
    out->same = relation == 2;
    out->reserved = 0;
    out->perm_q = perm_q;
    out->perm_c = L.perm_c;
    out->tgroup_q = tgroup_q;
    out->gstart_c = L.gstart_c;
    out->gmeta = L.gmeta;
    out->tmeta_q = tmeta_q;
    out->tmeta_c = L.tmeta_c;
    out->counters = L.counters;
    return GMD_OK;
}

extern "C" int gmd_tc_stage_grouped_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D,
                                        const double* center, const float* stats_c, const float* stats_q, void* qtiles,
                                        void* ctiles, void* workspace, size_t workspace_bytes, gmd_grouped* out,
                                        void* stream) {
    return stage_grouped(Q, Nq, C, Nc, D, center, stats_c, stats_q, qtiles, ctiles, workspace, workspace_bytes, 0, nullptr,
                         nullptr, out, stream);
}

extern "C" int gmd_tc_stage_grouped_ext_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D,
                                            const double* center, const float* stats_c, const float* stats_q,
                                            void* qtiles, void* ctiles, void* workspace, size_t workspace_bytes,
                                            int32_t nlead, const int32_t* nearest_idx, const double* nearest_dist,
                                            gmd_grouped* out, void* stream) {
    GMD_REQUIRE(nearest_idx && nearest_dist, GMD_ERR_BAD_ARG, "stage_grouped_ext: null nearest-leader arrays");
    return stage_grouped(Q, Nq, C, Nc, D, center, stats_c, stats_q, qtiles, ctiles, workspace, workspace_bytes, nlead,
                         nearest_idx, nearest_dist, out, stream);
}

// A staging (gmd_tc_stage_grouped_f64 / _ext_) whose queries were a run of the candidates' rows keeps the
// candidates' side in the workspace: leaders, nearest-leader keys, order, tiles, balls.  This call re-stages only
// the QUERY side for another run of the same length (Q = C + r0 * D): keys of the run -> padded order -> query
// tiles and their balls.  A caller that streams results back to the host walks the rows in runs this way, so
// that the copy of one run's results overlaps the scan of the next (engine.knn on host arrays).
extern "C" int gmd_tc_restage_queries_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D,
                                          const double* center, const float* stats_c, const float* stats_q,
                                          void* qtiles, void* workspace, size_t workspace_bytes, gmd_grouped* g,
                                          void* stream) {
    GMD_REQUIRE(Q && C && center && stats_c && stats_q && qtiles && workspace && g, GMD_ERR_BAD_ARG,
                "restage_queries: null pointer");
    GMD_REQUIRE(Nq > 0 && Nc > 0 && D >= 1 && D <= GMD_MAX_D, GMD_ERR_BAD_ARG, "restage_queries: bad sizes");
    int64_t q_row0;
    GMD_REQUIRE(relation_of(Q, Nq, C, Nc, D, &q_row0) == 1 && !g->same, GMD_ERR_BAD_ARG,
                "restage_queries: the queries must be a run of the candidates' rows, as in the staging before");
    Layout L = carve(workspace, Nq, Nc, D, 1);
    GMD_REQUIRE(L.bytes <= workspace_bytes, GMD_ERR_WORKSPACE, "restage_queries: workspace %zu B < required %zu B",
                workspace_bytes, L.bytes);
    GMD_REQUIRE(g->perm_c == L.perm_c && g->perm_q == L.perm_q && g->ngroups >= 1 && g->ngroups <= L.cap,
                GMD_ERR_BAD_ARG, "restage_queries: the descriptor does not belong to this workspace and shape");
    const int TM = gmd_tc::TM;
    const int nlead = g->ngroups;
    const int pad_q = query_pad(D);
    const int64_t npos_q = gmd_group_positions(Nq, nlead, TM * pad_q);
    const int nt_q = (int)(L.cap_pos_q / TM);
    int rc;
    if ((rc = gmd_group_order_pad(L.idx_c + q_row0, L.dist_c + q_row0, Nq, nlead, L.info, TM, pad_q, L.key_q, nullptr,
                                  L.perm_q, npos_q, L.gstart_q, L.tgroup_q, L.order_ws, L.order_ws_bytes, stream)) != GMD_OK)
        return rc;
    if ((rc = gmd_prep_tc_grouped_f64(Q, npos_q, D, center, 0, stats_c, stats_q, L.perm_q, qtiles, L.tmeta_q, nt_q,
                                      stream)) != GMD_OK)
        return rc;
    g->npos_q = npos_q;
    return GMD_OK;
}

extern "C" int gmd_scan_tc_grouped_f32(const void* qtiles, const void* ctiles, int D, int kk, int64_t c_index_offset,
                                       const float* stats_c, const float* stats_q, const gmd_grouped* g,
                                       uint32_t* cand_idx, float* cand_val, float* cand_tau, void* stream) {
    GMD_REQUIRE(g, GMD_ERR_BAD_ARG, "scan_tc_grouped: null descriptor");
    return gmd_scan_tc_prune_f32(qtiles, g->npos_q, ctiles, g->npos_c, D, kk, c_index_offset, stats_c, stats_q, g->perm_q,
                                 g->perm_c, g->tgroup_q, g->gstart_c, g->ngroups, g->gmeta, g->ng_alloc, g->tmeta_q,
                                 g->nt_alloc_q, g->tmeta_c, g->nt_alloc_c, g->same, cand_idx, cand_val, cand_tau,
                                 g->counters, stream);
}
