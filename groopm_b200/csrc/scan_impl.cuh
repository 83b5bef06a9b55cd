// The hot kernel: FP32 distance tiles + fused warp-level top-kk selection.
//
// One CTA owns BQ = 32*NW query rows (resident in shared memory for the whole
// scan) and streams a contiguous range of candidate tiles through an NST-stage
// shared-memory ring filled by a dedicated producer warp with bulk async copies
// (cp.async.bulk -> SASS UBLKCP) that complete on mbarriers.  Each compute warp
// owns 32 of the CTA's queries outright -- its 32q x 64c warp tile is an 8x8
// register tile per lane -- so the running top-kk lists (shared memory, sorted
// u64 keys = order-preserving(value) << 32 | index) are warp-private and the
// warps never synchronise with one another.  The fast path per candidate tile is
// D x 64 FFMA + a min-tree against the per-query threshold; only when some
// value beats a threshold does the warp take the cooperative insert path.
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
};

__device__ __forceinline__ float sel8(const float (&a)[8], int j) {
    float v = a[0];
#pragma unroll
    for (int t = 1; t < 8; ++t) v = (j == t) ? a[t] : v;
    return v;
}

// Insert (v, idx) into one query's sorted list, all 32 lanes cooperating.
__device__ __forceinline__ void warp_insert(uint64_t* lst, int kk, float v, uint32_t idx, int lane,
                                            float* thr_slot) {
    const unsigned FULL = 0xFFFFFFFFu;
    const uint64_t key = gmd_key(v, idx);
    const int x0 = lane, x1 = lane + 32, x2 = lane + 64;
    const uint64_t e0 = x0 < kk ? lst[x0] : GMD_EMPTY_KEY;
    const uint64_t e1 = x1 < kk ? lst[x1] : GMD_EMPTY_KEY;
    const uint64_t e2 = x2 < kk ? lst[x2] : GMD_EMPTY_KEY;
    const int pos = __popc(__ballot_sync(FULL, e0 < key)) + __popc(__ballot_sync(FULL, e1 < key)) +
                    __popc(__ballot_sync(FULL, e2 < key));
    if (pos >= kk) return;  // warp-uniform
    __syncwarp();
    if (x0 >= pos && x0 + 1 < kk) lst[x0 + 1] = e0;
    if (x1 >= pos && x1 + 1 < kk) lst[x1 + 1] = e1;
    if (x2 >= pos && x2 + 1 < kk) lst[x2 + 1] = e2;
    if (lane == 0) lst[pos] = key;
    __syncwarp();
    if (lane == 0) {
        const uint64_t last = lst[kk - 1];
        *thr_slot = (last == GMD_EMPTY_KEY) ? __int_as_float(0x7f800000) : gmd_unflip((uint32_t)(last >> 32));
    }
}

template <int D, int FORM, int NW, int NST>
__global__ void __launch_bounds__((NW + 1) * 32, 1) scan_topk_kernel(const ScanParams p) {
    constexpr int BQ = NW * 32;
    constexpr int TILE_F = (D + 1) * GMD_TW;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* Cs = reinterpret_cast<float*>(smem_raw);          // NST x [D+1][64]
    float* Qs = Cs + NST * TILE_F;                           // [D][BQ]
    float* qn = Qs + D * BQ;                                 // [BQ]  ||q||^2 (form 0)
    float* thr_s = qn + BQ;                                  // [BQ]  current kk-th value
    uint64_t* bars = reinterpret_cast<uint64_t*>(thr_s + BQ);  // full[NST], empty[NST]
    uint64_t* lists = bars + 2 * NST;                        // [BQ][kk]

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int kk = p.kk;
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
        for (int e = tid; e < BQ * kk; e += (NW + 1) * 32) lists[e] = GMD_EMPTY_KEY;
        for (int e = tid; e < BQ; e += (NW + 1) * 32) thr_s[e] = __int_as_float(0x7f800000);
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
    // no CTA-wide barrier below this line: the producer warp may retire early

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
    float thr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) thr[i] = __int_as_float(0x7f800000);

    for (int t = tb; t < te; ++t) {
        const int it = t - tb, s = it % NST;
        mbar_wait(&bars[s], (it / NST) & 1);
        const float* Ct = Cs + s * TILE_F + lc * 4;

        float acc[8][8];
        {
            // d = 0 peeled: initialises the accumulators
            const float4 qa = *reinterpret_cast<const float4*>(Qw);
            const float4 qb = *reinterpret_cast<const float4*>(Qw + 16);
            const float4 ca = *reinterpret_cast<const float4*>(Ct);
            const float4 cb = *reinterpret_cast<const float4*>(Ct + 32);
            const float q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            const float c[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
            if (FORM == GMD_FORM_EXPANDED) {
                const float4 na = *reinterpret_cast<const float4*>(Ct + D * GMD_TW);
                const float4 nb = *reinterpret_cast<const float4*>(Ct + D * GMD_TW + 32);
                const float cn[8] = {na.x, na.y, na.z, na.w, nb.x, nb.y, nb.z, nb.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(q[i], c[j], cn[j]);
            } else if (FORM == GMD_FORM_DIRECT_L2) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float df = q[i] - c[j];
                        acc[i][j] = df * df;
                    }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fabsf(q[i] - c[j]);
            }
        }
#pragma unroll(D <= 32 ? D : 8)
        for (int d = 1; d < D; ++d) {
            const float4 qa = *reinterpret_cast<const float4*>(Qw + d * BQ);
            const float4 qb = *reinterpret_cast<const float4*>(Qw + d * BQ + 16);
            const float4 ca = *reinterpret_cast<const float4*>(Ct + d * GMD_TW);
            const float4 cb = *reinterpret_cast<const float4*>(Ct + d * GMD_TW + 32);
            const float q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            const float c[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (FORM == GMD_FORM_EXPANDED) {
                        acc[i][j] = fmaf(q[i], c[j], acc[i][j]);
                    } else if (FORM == GMD_FORM_DIRECT_L2) {
                        const float df = q[i] - c[j];
                        acc[i][j] = fmaf(df, df, acc[i][j]);
                    } else {
                        acc[i][j] += fabsf(q[i] - c[j]);
                    }
                }
        }
        // this stage's shared memory is no longer needed by this warp
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[NST + s]);

        // ---- selection fast path: one min-tree per query row against its threshold
        bool hit = false;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float m01 = fminf(acc[i][0], acc[i][1]), m23 = fminf(acc[i][2], acc[i][3]);
            const float m45 = fminf(acc[i][4], acc[i][5]), m67 = fminf(acc[i][6], acc[i][7]);
            const float m = fminf(fminf(m01, m23), fminf(m45, m67));
            hit |= (m < thr[i]);
        }
        if (__any_sync(FULL, hit)) {
            // ---- rare path: cooperative inserts, candidate by candidate
            const uint32_t cbase = (uint32_t)t * GMD_TW;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                unsigned hm = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) hm |= (acc[i][j] < thr[i]) ? (1u << j) : 0u;
                unsigned bal = __ballot_sync(FULL, hm != 0);
                while (bal) {
                    const int src = __ffs(bal) - 1;
                    bal &= bal - 1;
                    unsigned shm = __shfl_sync(FULL, hm, src);
                    const int ql = ((i >> 2) << 4) + ((src >> 3) << 2) + (i & 3);
                    uint64_t* lst = lists + (size_t)(warp * 32 + ql) * kk;
                    while (shm) {
                        const int j = __ffs(shm) - 1;
                        shm &= shm - 1;
                        const float v = __shfl_sync(FULL, sel8(acc[i], j), src);
                        const uint32_t cidx = cbase + ((j >> 2) << 5) + ((src & 7) << 2) + (j & 3);
                        warp_insert(lst, kk, v, cidx, lane, &thr_s[warp * 32 + ql]);
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) thr[i] = thr_s[warp * 32 + ((i >> 2) << 4) + (lq << 2) + (i & 3)];
        }
    }

    // ---- write this warp's lists: cand_*[q][split][kk]
    __syncwarp();
    for (int ql = 0; ql < 32; ++ql) {
        const long long q = q0 + warp * 32 + ql;
        if (q >= p.Nq) break;
        const uint64_t* lst = lists + (size_t)(warp * 32 + ql) * kk;
        const float qnv = qn[warp * 32 + ql];
        const size_t ob = ((size_t)q * p.nsplit + split) * kk;
        for (int x = lane; x < kk; x += 32) {
            const uint64_t key = lst[x];
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
    }
}

template <int D, int FORM, int NW, int NST>
static size_t scan_smem_bytes(int kk) {
    constexpr int BQ = NW * 32;
    constexpr int TILE_F = (D + 1) * GMD_TW;
    return (size_t)(NST * TILE_F + D * BQ + 2 * BQ) * 4 + 2 * NST * 8 + (size_t)BQ * kk * 8;
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
