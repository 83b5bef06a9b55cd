// Shared pieces of the tensor-core scans (scan_tc.cu, scan_tc_prune.cu): tile geometry, UMMA descriptor /
// tcgen05 wrappers.  See scan_tc.cu for the design.
#pragma once
#include <cuda_fp16.h>

#include "scan_impl.cuh"

namespace gmd_tc {
using namespace gmd_scan;

constexpr int TM = 128;                  // profiles per TC tile (UMMA M and N)
constexpr int TK = 32;                   // contraction length of one K block (64-byte rows of halves)
constexpr int IMG_B = TM * TK * 2;       // one FP16 image of one K block: 8 KB
constexpr int KB_B = 2 * IMG_B;          // hi + lo images of one K block: 16 KB = one ring stage
constexpr int MAX_KB = 5;                // K blocks per tile: D + 1 <= 160 covers the 136-d signatures
constexpr int NS_MAX = 8;                // most ring stages
constexpr int NACC = 4;                  // accumulator stages in tensor memory (4 x 128 of the 512 columns)
__host__ __device__ constexpr int tile_bytes(int nkb) { return nkb * KB_B + 1024; }  // K blocks, then norms
constexpr float QCOL = 8.0f;             // column D of a query row; a candidate row holds -||c||^2 / (2 QCOL)
constexpr float PAD_COL = -60000.0f;     // column D of a padding candidate: u = -480000, below any real value
// Warp roles.  The scheduler of an SM sub-partition picks the eligible warp with the HIGHEST id first, so the
// single-thread roles whose latency the whole pipeline waits on sit above the selection warps.
constexpr int W_PROD = 4;                // warps 0..3: selection (TMEM lane quarter = warp id), 4: producer
constexpr int W_ISSUE = 5;               // 5, 6: MMA issuers (even / odd candidate tiles)
constexpr int NTHREADS = 224;

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    // K-major SWIZZLE_64B: LBO 1, SBO 512 B (8 rows x 64 B), descriptor version 1 (sm_100), layout type 4
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem),
        "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// wait for the outstanding tcgen05.ld; the registers pass through the asm so that nothing that
// reads them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}


constexpr int TQ = 16;   // parked hits per selection thread (a thread owns its query's queue)

// Every lane folds its own queue into its own heap (the lanes work side by side: this is
// why hits are parked instead of being pushed the moment they are seen, which would
// serialise 32 divergent heap updates per warp).  Returns the new threshold value.
static __device__ __noinline__ float drain_own(uint64_t* h, int kk, const uint2* q, int n) {
    for (int r = 0; r < n; ++r) {
        const uint2 e = q[r * TM];
        const uint64_t key = gmd_key(__uint_as_float(e.y), e.x);
        if (key < h[0]) heap_replace_root(h, kk, key);
    }
    const uint64_t root = h[0];
    return (root == GMD_EMPTY_KEY) ? __int_as_float(0x7f800000) : gmd_unflip((uint32_t)(root >> 32));
}

}  // namespace gmd_tc

static inline int tc_nkb(int D) { return (D + 1 + gmd_tc::TK - 1) / gmd_tc::TK; }
