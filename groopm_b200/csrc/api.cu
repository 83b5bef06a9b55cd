// C ABI glue: error state, policy helpers, the device-level pipeline and the
// host-level drop-ins (see include/gmdist.h for the reference call sites).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

#include "scan_impl.cuh"

static thread_local char g_err[512] = "";

void gmd_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* gmd_last_error(void) { return g_err; }
extern "C" int gmd_version(void) { return 100; }

extern "C" int gmd_pick_form(int metric, int D) {
    if (metric == GMD_L1) return GMD_FORM_DIRECT_L1;
    // tiny D: the direct form costs <= 6 instructions per pair and is immune to the
    // cancellation of ||a||^2 + ||b||^2 - 2a.b on [0, 999]^3 transformedCP data
    return D <= 3 ? GMD_FORM_DIRECT_L2 : GMD_FORM_EXPANDED;
}

// k = 1 with the expanded form and a short profile: the dedicated best-matching-unit scan
extern "C" int gmd_use_bmu_scan(int k, int metric, int D) {
    if (k != 1 || metric != GMD_L2 || gmd_pick_form(metric, D) != GMD_FORM_EXPANDED) return 0;
    const int DP = gmd_scan_pad_dim(D);
    return DP >= 4 && DP <= GMD_BMU_MAX_D;
}
extern "C" int gmd_bmu_kk(void) { return gmd_scan::BMU_KK; }

// The tensor-core scan (scan_tc.cu) wins once both sides are large enough to fill the SMs with
// its 128-query CTAs and to amortise its per-CTA set-up (measured cross-over ~50 k rows at
// D = 10; 2.4 x at 250 k x 30, 3.2 x at 1 M x 30).  GMD_TC=0 / 1 overrides (developer switch).
extern "C" int gmd_use_tc_scan(int64_t Nq, int64_t Nc, int k, int metric, int D) {
    if (!gmd_tc_supported(k, metric, D)) return 0;
    if (const char* env = getenv("GMD_TC")) return atoi(env) != 0;
    return Nq >= 32768 && Nc >= 32768;
}

extern "C" int gmd_pick_kk(int k) { return k + std::min(8, std::max(3, k / 4)); }

extern "C" int gmd_pick_nsplit(int64_t Nq, int64_t Nc, int D, int kk) {
    (void)kk;
    if (const char* env = getenv("GMD_NSPLIT")) {  // tuning / profiling override
        const int v = atoi(env);
        if (v >= 1) return (int)std::min<int64_t>(v, std::max<int64_t>(1, gmd_ntiles(Nc)));
    }
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int DP = gmd_scan_pad_dim(D);
    const int bq = gmd_scan_block_queries(DP < 0 ? 64 : DP);
    const int64_t nqb = (Nq + bq - 1) / bq;
    const int64_t ntiles = gmd_ntiles(Nc);
    const int smax = (int)std::max<int64_t>(1, std::min<int64_t>(16, ntiles / 16));
    int best = 1;
    double best_cost = 1e300;
    for (int s = 1; s <= smax; ++s) {
        const double ctas = (double)nqb * s;
        const double waves = std::ceil(ctas / sms);
        // per-CTA time ~ tiles scanned + the cost of a split's own heap warm-up, hit
        // handling and write-out, measured at 300-500 tile times on B200 (cfg2 / cfg5 sweeps,
        // DESIGN.md section 5): cutting the candidates only pays when the query blocks
        // alone leave most SMs idle
        const double cost = waves * ((double)ntiles / s + 300.0);
        if (cost < best_cost * 0.98) {
            best_cost = cost;
            best = s;
        }
    }
    return best;
}

extern "C" int gmd_scan_topk_f32(const float* qtiles, int64_t Nq, const float* ctiles, int64_t Nc, int D, int form,
                                 int kk, int nsplit, int64_t c_index_offset, uint32_t* cand_idx, float* cand_val,
                                 float* cand_tau, void* stream) {
    GMD_REQUIRE(qtiles && ctiles && cand_idx && cand_val && cand_tau, GMD_ERR_BAD_ARG, "scan: null pointer");
    GMD_REQUIRE(Nq > 0 && Nc > 0 && Nc + c_index_offset < 0xFFFFFFFFll, GMD_ERR_BAD_ARG, "scan: bad sizes");
    GMD_REQUIRE(kk >= 1 && kk <= 96 && nsplit >= 1, GMD_ERR_BAD_ARG, "scan: kk=%d nsplit=%d out of range", kk, nsplit);
    const int DP = gmd_scan_pad_dim(D);
    GMD_REQUIRE(DP > 0, GMD_ERR_UNSUPPORTED, "scan: D=%d not supported (max %d)", D, GMD_MAX_D);
    gmd_scan::ScanParams p;
    p.qtiles = qtiles;
    p.ctiles = ctiles;
    p.Nq = Nq;
    p.Nc = Nc;
    p.ntiles_q = (int)gmd_ntiles(Nq);
    p.ntiles_c = (int)gmd_ntiles(Nc);
    p.nsplit = nsplit;
    p.kk = kk;
    p.c_index_offset = c_index_offset;
    p.cand_idx = cand_idx;
    p.cand_val = cand_val;
    p.cand_tau = cand_tau;
    cudaStream_t st = (cudaStream_t)stream;
    switch (form) {
        case GMD_FORM_EXPANDED:
            return gmd_scan_launch_form0(DP, p, st);
        case GMD_FORM_DIRECT_L2:
            return gmd_scan_launch_form1(DP, p, st);
        case GMD_FORM_DIRECT_L1:
            return gmd_scan_launch_form2(DP, p, st);
        default:
            break;
    }
    gmd_set_error("scan: bad form %d", form);
    return GMD_ERR_BAD_ARG;
}

static int scan_bmu(const float* qtiles, int64_t Nq, const float* ctiles, int64_t Nc, int D, int64_t c_index_offset,
                    uint32_t* cand_idx, float* cand_val, float* cand_tau, uint32_t* best1, void* stream) {
    GMD_REQUIRE(qtiles && ctiles && cand_idx && cand_val && cand_tau, GMD_ERR_BAD_ARG, "bmu scan: null pointer");
    GMD_REQUIRE(Nq > 0 && Nc > 0 && Nc + c_index_offset < 0xFFFFFFFFll && Nc < (1ll << 29), GMD_ERR_BAD_ARG,
                "bmu scan: bad sizes");
    const int DP = gmd_scan_pad_dim(D);
    GMD_REQUIRE(DP >= 4 && DP <= GMD_BMU_MAX_D, GMD_ERR_UNSUPPORTED, "bmu scan: D=%d outside its range", D);
    gmd_scan::ScanParams p;
    p.qtiles = qtiles;
    p.ctiles = ctiles;
    p.Nq = Nq;
    p.Nc = Nc;
    p.ntiles_q = (int)gmd_ntiles(Nq);
    p.ntiles_c = (int)gmd_ntiles(Nc);
    p.nsplit = 1;
    p.kk = gmd_scan::BMU_KK;
    p.c_index_offset = c_index_offset;
    p.cand_idx = cand_idx;
    p.cand_val = cand_val;
    p.cand_tau = cand_tau;
    p.best1 = best1;
    return gmd_bmu_launch(DP, p, (cudaStream_t)stream);
}

extern "C" int gmd_scan_bmu_f32(const float* qtiles, int64_t Nq, const float* ctiles, int64_t Nc, int D,
                                int64_t c_index_offset, uint32_t* cand_idx, float* cand_val, float* cand_tau,
                                void* stream) {
    return scan_bmu(qtiles, Nq, ctiles, Nc, D, c_index_offset, cand_idx, cand_val, cand_tau, nullptr, stream);
}

// ------------------------------------------------------------------ workspace layout
namespace {
struct Workspace {
    double* center;     // [D]
    float* stats;       // [4]  (2 used by candidates, 2 scratch for queries)
    int32_t* flags;     // [4]  flags[0] non-finite, flags[1] fail_count
    float* ctiles;
    float* qtiles;
    uint32_t* cand_idx;
    float* cand_val;
    float* cand_tau;
    int32_t* fail_rows;
    void* exact_ws;     // hit lists of the batched exact pass
    size_t exact_bytes;
    void* grouped_ws;   // staging of the pruned tensor-core scan (grouped.cu)
    size_t grouped_bytes;
    size_t bytes;
};

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

Workspace carve(void* base, int64_t Nq, int64_t Nc, int D, int kk, int nsplit, bool same, bool tc = false,
                bool prune = false) {
    Workspace w;
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t n) {
        char* r = p ? p + off : nullptr;
        off += align_up(n);
        return r;
    };
    w.center = (double*)take(sizeof(double) * GMD_MAX_D);
    w.stats = (float*)take(sizeof(float) * 4);
    w.flags = (int32_t*)take(sizeof(int32_t) * 4);
    w.ctiles = (float*)take(prune ? gmd_tc_grouped_tiles_bytes(Nc, Nc, D) : tc ? gmd_tc_tiles_bytes(Nc, D) : gmd_tiles_bytes(Nc, D));
    w.qtiles = (same && !tc) ? w.ctiles
                             : (float*)take(prune ? gmd_tc_grouped_tiles_bytes(Nq, Nc, D)
                                            : tc  ? gmd_tc_tiles_bytes(Nq, D)
                                                  : gmd_tiles_bytes(Nq, D));
    w.cand_idx = (uint32_t*)take((size_t)Nq * nsplit * kk * sizeof(uint32_t));
    w.cand_val = (float*)take((size_t)Nq * nsplit * kk * sizeof(float));
    w.cand_tau = (float*)take((size_t)Nq * nsplit * sizeof(float));
    w.fail_rows = (int32_t*)take((size_t)Nq * sizeof(int32_t));
    w.exact_bytes = gmd_exact_workspace_bytes(Nq);
    w.exact_ws = take(w.exact_bytes);
    w.grouped_bytes = prune ? gmd_tc_grouped_workspace_bytes(Nq, Nc, D) : 0;
    w.grouped_ws = take(w.grouped_bytes);
    w.bytes = off;
    return w;
}
}  // namespace

extern "C" size_t gmd_knn_workspace_bytes(int64_t Nq, int64_t Nc, int D, int k) {
    // same policy as gmd_knn_f64 (the larger of the two scans' needs); sized for separate query tiles
    const int kk = std::max(gmd_pick_kk(k), k == 1 ? gmd_scan::BMU_KK : 0);
    const size_t plain = carve(nullptr, Nq, Nc, D, kk, gmd_pick_nsplit(Nq, Nc, D, kk), false).bytes;
    const size_t tc = gmd_tc_supported(k, GMD_L2, D) ? carve(nullptr, Nq, Nc, D, kk, 1, false, true).bytes : 0;
    const size_t pr = gmd_use_tc_prune(Nq, Nc, k, GMD_L2, D) ? carve(nullptr, Nq, Nc, D, kk, 1, false, true, true).bytes : 0;
    return std::max(plain, std::max(tc, pr));
}

extern "C" int gmd_knn_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D, int k, int metric,
                           int form, int32_t* out_idx, double* out_dist, void* workspace, size_t workspace_bytes,
                           int32_t* info, void* stream) {
    GMD_REQUIRE(Q && C && out_idx && out_dist && workspace, GMD_ERR_BAD_ARG, "knn: null pointer");
    GMD_REQUIRE(Nq > 0 && Nc > 0, GMD_ERR_BAD_ARG, "knn: empty input (Nq=%lld Nc=%lld)", (long long)Nq, (long long)Nc);
    GMD_REQUIRE(D >= 1 && D <= GMD_MAX_D, GMD_ERR_UNSUPPORTED, "knn: D=%d outside [1, %d]", D, GMD_MAX_D);
    GMD_REQUIRE(k >= 1 && k <= GMD_MAX_K, GMD_ERR_UNSUPPORTED, "knn: k=%d outside [1, %d]", k, GMD_MAX_K);
    GMD_REQUIRE(k <= Nc, GMD_ERR_BAD_ARG, "knn: k=%d exceeds the %lld candidates", k, (long long)Nc);
    GMD_REQUIRE(metric == GMD_L2 || metric == GMD_L1, GMD_ERR_BAD_ARG, "knn: bad metric %d", metric);
    if (form == GMD_FORM_AUTO) form = gmd_use_tc_scan(Nq, Nc, k, metric, D) ? GMD_FORM_TC : gmd_pick_form(metric, D);
    GMD_REQUIRE((metric == GMD_L1) == (form == GMD_FORM_DIRECT_L1), GMD_ERR_BAD_ARG,
                "knn: form %d does not compute metric %d", form, metric);
    const bool tc = form == GMD_FORM_TC;
    GMD_REQUIRE(!tc || gmd_tc_supported(k, metric, D), GMD_ERR_UNSUPPORTED,
                "knn: the tensor-core scan needs the Euclidean metric, 2 <= k <= 64 and 4 <= D <= 136");
    const bool bmu = form == GMD_FORM_EXPANDED && gmd_use_bmu_scan(k, metric, D);
    const int kk = bmu ? gmd_scan::BMU_KK : gmd_pick_kk(k);
    // large tensor-core problems go through the grouped order and the pruned scan 
    const bool prune = tc && gmd_use_tc_prune(Nq, Nc, k, metric, D);
    const int nsplit = (bmu || tc) ? 1 : gmd_pick_nsplit(Nq, Nc, D, kk);
    const bool same = (Q == C && Nq == Nc);
    Workspace w = carve(workspace, Nq, Nc, D, kk, nsplit, same, tc, prune);
    GMD_REQUIRE(w.bytes <= workspace_bytes, GMD_ERR_WORKSPACE, "knn: workspace %zu B < required %zu B",
                workspace_bytes, w.bytes);
    gmd_grouped grouped;
    cudaStream_t st = (cudaStream_t)stream;
    GMD_CUDA_CHECK(cudaMemsetAsync(w.stats, 0, sizeof(float) * 4, st));
    GMD_CUDA_CHECK(cudaMemsetAsync(w.flags, 0, sizeof(int32_t) * 4, st));
    int rc;
    if ((rc = gmd_col_mean_f64(C, Nc, D, w.center, stream)) != GMD_OK) return rc;
    if (tc) {
        if ((rc = gmd_tc_norms_f64(C, Nc, D, w.center, w.stats, w.flags, stream)) != GMD_OK) return rc;
        if (!same && (rc = gmd_tc_norms_f64(Q, Nq, D, w.center, w.stats + 2, w.flags, stream)) != GMD_OK) return rc;
        if (prune) {
            if ((rc = gmd_tc_stage_grouped_f64(Q, Nq, C, Nc, D, w.center, w.stats, w.stats + 2, w.qtiles, w.ctiles,
                                               w.grouped_ws, w.grouped_bytes, &grouped, stream)) != GMD_OK)
                return rc;
        } else {
            if ((rc = gmd_prep_tc_f64(C, Nc, D, w.center, 1, w.stats, w.stats + 2, w.ctiles, stream)) != GMD_OK) return rc;
            if ((rc = gmd_prep_tc_f64(Q, Nq, D, w.center, 0, w.stats, w.stats + 2, w.qtiles, stream)) != GMD_OK) return rc;
        }
    } else {
        if ((rc = gmd_prep_f64(C, Nc, D, w.center, form, w.ctiles, w.stats, w.flags, stream)) != GMD_OK) return rc;
        if (!same && (rc = gmd_prep_f64(Q, Nq, D, w.center, form, w.qtiles, w.stats + 2, w.flags, stream)) != GMD_OK)
            return rc;
    }
    rc = prune ? gmd_scan_tc_grouped_f32(w.qtiles, w.ctiles, D, kk, 0, w.stats, w.stats + 2, &grouped, w.cand_idx,
                                         w.cand_val, w.cand_tau, stream) :
         tc ? gmd_scan_tc_f32(w.qtiles, Nq, w.ctiles, Nc, D, kk, 0, w.stats, w.stats + 2, w.cand_idx, w.cand_val,
                              w.cand_tau, stream) :
         bmu ? gmd_scan_bmu_f32(w.qtiles, Nq, w.ctiles, Nc, D, 0, w.cand_idx, w.cand_val, w.cand_tau, stream)
             : gmd_scan_topk_f32(w.qtiles, Nq, w.ctiles, Nc, D, form, kk, nsplit, 0, w.cand_idx, w.cand_val,
                                 w.cand_tau, stream);
    if (rc != GMD_OK) return rc;
    if ((rc = gmd_refine_topk_order_f64(Q, Nq, C, Nc, D, w.center, k, kk, nsplit, metric, form, w.cand_idx, w.cand_val,
                                        w.cand_tau, w.stats, out_idx, out_dist, w.flags + 1, w.fail_rows,
                                        prune ? grouped.perm_q : nullptr, prune ? grouped.npos_q : 0, stream)) != GMD_OK)
        return rc;
    if ((rc = gmd_exact_rows_f64(Q, Nq, C, Nc, D, k, metric, w.flags + 1, w.fail_rows, out_idx, out_dist, w.exact_ws,
                                 w.exact_bytes, stream)) != GMD_OK)
        return rc;
    if (info) {
        // info[0] = rows finished by the exact scan, info[1] = non-finite input seen
        GMD_CUDA_CHECK(cudaMemcpyAsync(info, w.flags + 1, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        GMD_CUDA_CHECK(cudaMemcpyAsync(info + 1, w.flags, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    }
    return GMD_OK;
}

// A leader for every row (the grouped staging of the pruned scan): the FP32 best-matching-unit scan's best
// proposal and the exact float64 distance to it -- no re-rank, no certificate: the staging needs a good,
// deterministic assignment with the true distance to the ASSIGNED leader (its balls are built from that), not the
// proven nearest one.  Outside the BMU scan's range of D: the exact k = 1 pipeline.  Workspace:
// gmd_knn_workspace_bytes(N, nlead, D, 1).
extern "C" int gmd_assign_leaders_f64(const double* X, int64_t N, const double* L, int nlead, int D, int32_t* out_idx,
                                      double* out_dist, void* workspace, size_t workspace_bytes, void* stream) {
    GMD_REQUIRE(X && L && out_idx && out_dist && workspace, GMD_ERR_BAD_ARG, "assign_leaders: null pointer");
    GMD_REQUIRE(N > 0 && nlead >= 1 && D >= 1 && D <= GMD_MAX_D, GMD_ERR_BAD_ARG, "assign_leaders: bad sizes");
    if (!gmd_use_bmu_scan(1, GMD_L2, D))
        return gmd_knn_f64(X, N, L, nlead, D, 1, GMD_L2, GMD_FORM_AUTO, out_idx, out_dist, workspace, workspace_bytes,
                           nullptr, stream);
    const int kk = gmd_scan::BMU_KK;
    Workspace w = carve(workspace, N, nlead, D, kk, 1, false);
    GMD_REQUIRE(w.bytes <= workspace_bytes, GMD_ERR_WORKSPACE, "assign_leaders: workspace %zu B < required %zu B",
                workspace_bytes, w.bytes);
    cudaStream_t st = (cudaStream_t)stream;
    GMD_CUDA_CHECK(cudaMemsetAsync(w.stats, 0, sizeof(float) * 4, st));
    GMD_CUDA_CHECK(cudaMemsetAsync(w.flags, 0, sizeof(int32_t) * 4, st));
    int rc;
    if ((rc = gmd_col_mean_f64(L, nlead, D, w.center, stream)) != GMD_OK) return rc;
    if ((rc = gmd_prep_f64(L, nlead, D, w.center, GMD_FORM_EXPANDED, w.ctiles, w.stats, w.flags, stream)) != GMD_OK) return rc;
    if ((rc = gmd_prep_f64(X, N, D, w.center, GMD_FORM_EXPANDED, w.qtiles, w.stats + 2, w.flags, stream)) != GMD_OK) return rc;
    // (the scan also names the best candidate of the winning cell, by its own FP32 arithmetic: fail_rows as scratch)
    uint32_t* best1 = (uint32_t*)w.fail_rows;
    if ((rc = scan_bmu(w.qtiles, N, w.ctiles, nlead, D, 0, w.cand_idx, w.cand_val, w.cand_tau, best1, stream)) != GMD_OK)
        return rc;
    return gmd_pick_proposal_f64(X, N, L, D, 1, best1, w.cand_val, out_idx, out_dist, stream);
}

// ------------------------------------------------------------------ host-level drop-ins
namespace {
struct DevBuf {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    ~DevBuf() {
        if (p) cudaFreeAsync(p, st);
    }
    int alloc(size_t n, cudaStream_t s) {
        st = s;
        GMD_CUDA_CHECK(cudaMallocAsync(&p, std::max<size_t>(n, 16), s));
        return GMD_OK;
    }
};
struct StreamGuard {
    cudaStream_t st = nullptr;
    ~StreamGuard() {
        if (st) cudaStreamDestroy(st);
    }
};
}  // namespace

// Pageable host memory <-> device through two pinned staging halves: the host-side copy of a piece is cut over a
// few threads (one thread moves ~9 GB/s, the link ~55) and overlaps the DMA of the piece before it.  The halves are
// allocated once per process (32 MB pinned) and guarded by a mutex: host-level calls copy one at a time.
namespace {
constexpr size_t BOUNCE_BYTES = (size_t)16 << 20;
struct Bounce {
    std::mutex mu;
    void* half[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int ensure() {
        for (int b = 0; b < 2; ++b) {
            if (!half[b]) GMD_CUDA_CHECK(cudaHostAlloc(&half[b], BOUNCE_BYTES, cudaHostAllocDefault));
            if (!ev[b]) GMD_CUDA_CHECK(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming));
        }
        return GMD_OK;
    }
};
Bounce g_bounce;

void host_copy(void* dst, const void* src, size_t n) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nt = std::min<size_t>(std::min<unsigned>(8u, hw), std::max<size_t>(1, n >> 20));   // >= 1 MB per thread
    if (nt <= 1) {
        memcpy(dst, src, n);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((n + nt - 1) / nt + 4095) & ~(size_t)4095;
    for (size_t t = 1; t < nt; ++t) {
        const size_t o = t * per;
        if (o >= n) break;
        th.emplace_back([=] { memcpy((char*)dst + o, (const char*)src + o, std::min(per, n - o)); });
    }
    memcpy(dst, src, std::min(per, n));
    for (auto& x : th) x.join();
}

// caller holds g_bounce.mu
int upload(void* dev, const void* host, size_t n, cudaStream_t st) {
    Bounce& B = g_bounce;
    int i = 0;
    for (size_t o = 0; o < n; o += BOUNCE_BYTES, ++i) {
        const int b = i & 1;
        const size_t len = std::min(BOUNCE_BYTES, n - o);
        if (i >= 2) GMD_CUDA_CHECK(cudaEventSynchronize(B.ev[b]));     // the half's previous piece has left
        host_copy(B.half[b], (const char*)host + o, len);
        GMD_CUDA_CHECK(cudaMemcpyAsync((char*)dev + o, B.half[b], len, cudaMemcpyHostToDevice, st));
        GMD_CUDA_CHECK(cudaEventRecord(B.ev[b], st));
    }
    return GMD_OK;
}

int download(void* host, const void* dev, size_t n, cudaStream_t st) {
    Bounce& B = g_bounce;
    const size_t pieces = (n + BOUNCE_BYTES - 1) / BOUNCE_BYTES;
    for (size_t i = 0; i <= pieces; ++i) {
        if (i < pieces) {
            const size_t o = i * BOUNCE_BYTES;
            GMD_CUDA_CHECK(cudaMemcpyAsync(B.half[i & 1], (const char*)dev + o, std::min(BOUNCE_BYTES, n - o),
                                           cudaMemcpyDeviceToHost, st));
            GMD_CUDA_CHECK(cudaEventRecord(B.ev[i & 1], st));
        }
        if (i >= 1) {                                                  // the piece before: out of its half
            const size_t o = (i - 1) * BOUNCE_BYTES;
            GMD_CUDA_CHECK(cudaEventSynchronize(B.ev[(i - 1) & 1]));
            host_copy((char*)host + o, B.half[(i - 1) & 1], std::min(BOUNCE_BYTES, n - o));
        }
    }
    return GMD_OK;
}
}  // namespace

extern "C" int gmd_knn_host_f64(const double* Q, int64_t Nq, const double* C, int64_t Nc, int D, int k, int metric,
                                int32_t* out_idx, double* out_dist) {
    GMD_REQUIRE(Q && C && out_idx, GMD_ERR_BAD_ARG, "knn_host: null pointer");
    GMD_REQUIRE(Nq > 0 && Nc > 0 && D >= 1, GMD_ERR_BAD_ARG, "knn_host: empty input");
    GMD_REQUIRE(D <= GMD_MAX_D, GMD_ERR_UNSUPPORTED, "knn_host: D=%d > %d", D, GMD_MAX_D);
    GMD_REQUIRE(k >= 1 && k <= GMD_MAX_K, GMD_ERR_UNSUPPORTED, "knn_host: k=%d outside [1, %d]", k, GMD_MAX_K);
    GMD_REQUIRE(k <= Nc, GMD_ERR_BAD_ARG, "knn_host: k=%d exceeds the %lld candidates", k, (long long)Nc);
    StreamGuard sg;
    GMD_CUDA_CHECK(cudaStreamCreateWithFlags(&sg.st, cudaStreamNonBlocking));
    cudaStream_t st = sg.st;
    {
        // the device buffers of a call come from the stream-ordered pool; by default the pool hands freed memory
        // back to the driver at the next synchronisation, and a loop of host-level calls then re-creates ~2 GB of
        // mappings every time (measured: 65 to 900 ms per call).  Let the pool keep what it has.
        static std::once_flag once;
        std::call_once(once, [] {
            int dev = 0;
            cudaMemPool_t pool;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            cudaGetLastError();
        });
    }
    const bool same = (Q == C && Nq == Nc);
    DevBuf dC, dQ, dI, dD, dW, dInfo;
    int rc;
    const size_t cb = (size_t)Nc * D * sizeof(double), qb = (size_t)Nq * D * sizeof(double);
    // small problems: plain pageable copies (no staging, no lock)
    const bool staged = cb + (same ? 0 : qb) + (size_t)Nq * k * 12 >= ((size_t)8 << 20);
    std::unique_lock<std::mutex> lock(g_bounce.mu, std::defer_lock);
    if (staged) {
        lock.lock();
        if ((rc = g_bounce.ensure()) != GMD_OK) return rc;
    }
    auto up = [&](void* dev, const void* host, size_t n) -> int {
        if (staged) return upload(dev, host, n, st);
        GMD_CUDA_CHECK(cudaMemcpyAsync(dev, host, n, cudaMemcpyHostToDevice, st));
        return GMD_OK;
    };
    auto down = [&](void* host, const void* dev, size_t n) -> int {
        if (staged) return download(host, dev, n, st);
        GMD_CUDA_CHECK(cudaMemcpyAsync(host, dev, n, cudaMemcpyDeviceToHost, st));
        return GMD_OK;
    };
    if ((rc = dC.alloc(cb, st)) != GMD_OK) return rc;
    if ((rc = up(dC.p, C, cb)) != GMD_OK) return rc;
    const double* dq = (const double*)dC.p;
    if (!same) {
        if ((rc = dQ.alloc(qb, st)) != GMD_OK) return rc;
        if ((rc = up(dQ.p, Q, qb)) != GMD_OK) return rc;
        dq = (const double*)dQ.p;
    }
    const size_t wsb = gmd_knn_workspace_bytes(Nq, Nc, D, k);
    if ((rc = dI.alloc((size_t)Nq * k * sizeof(int32_t), st)) != GMD_OK) return rc;
    if ((rc = dD.alloc((size_t)Nq * k * sizeof(double), st)) != GMD_OK) return rc;
    if ((rc = dW.alloc(wsb, st)) != GMD_OK) return rc;
    if ((rc = dInfo.alloc(2 * sizeof(int32_t), st)) != GMD_OK) return rc;
    rc = gmd_knn_f64(dq, Nq, (const double*)dC.p, Nc, D, k, metric, GMD_FORM_AUTO, (int32_t*)dI.p, (double*)dD.p,
                     dW.p, wsb, (int32_t*)dInfo.p, st);
    if (rc != GMD_OK) return rc;
    int32_t info[2] = {0, 0};
    GMD_CUDA_CHECK(cudaMemcpyAsync(info, dInfo.p, sizeof(info), cudaMemcpyDeviceToHost, st));
    if ((rc = down(out_idx, dI.p, (size_t)Nq * k * sizeof(int32_t))) != GMD_OK) return rc;
    if (out_dist && (rc = down(out_dist, dD.p, (size_t)Nq * k * sizeof(double))) != GMD_OK) return rc;
    GMD_CUDA_CHECK(cudaStreamSynchronize(st));
    GMD_REQUIRE(info[1] == 0, GMD_ERR_NONFINITE, "non-finite value in an input profile");
    return GMD_OK;
}

extern "C" int gmd_bmu_host_f64(const double* nodes, int64_t M, const double* X, int64_t N, int D, int32_t* out_loc,
                                double* out_dist) {
    return gmd_knn_host_f64(X, N, nodes, M, D, 1, GMD_L2, out_loc, out_dist);
}
