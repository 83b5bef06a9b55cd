// Shared definitions for libgmdist's kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/gmdist.h"

#define GMD_TW GMD_TILE_ROWS           // profile rows per tile
#define GMD_PAD_AUX 1.0e30f            // ||c||^2 of a padding row (expanded form)
#define GMD_PAD_COORD 1.0e15f          // coordinate of a padding row (direct forms)
#define GMD_EMPTY_IDX 0xFFFFFFFFu
#define GMD_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

void gmd_set_error(const char* fmt, ...);

#define GMD_CUDA_CHECK(expr)                                                        \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            gmd_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                          __FILE__, __LINE__);                                      \
            return GMD_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define GMD_REQUIRE(cond, code, ...)                                                \
    do {                                                                            \
        if (!(cond)) {                                                              \
            gmd_set_error(__VA_ARGS__);                                             \
            return (code);                                                          \
        }                                                                           \
    } while (0)

static inline int64_t gmd_ntiles(int64_t n) { return (n + GMD_TW - 1) / GMD_TW; }

// Tensor-core scan: the centred rows are multiplied by 2^e before the FP16 hi/lo split so that the
// largest row norm of the query and candidate sets lands in (128, 256].  stats[0] / stats[2] hold
// max ||x||^2 of the candidate / query set (0 when the sets are the same array).
__host__ __device__ __forceinline__ int gmd_tc_exponent(float max_n2_a, float max_n2_b) {
    const float m = fmaxf(max_n2_a, max_n2_b);
    if (!(m > 0.0f) || !(m < 3.0e38f)) return 0;
    int ex;
    frexpf(sqrtf(m), &ex);          // sqrt(m) = f * 2^ex, f in [0.5, 1)
    int e = 8 - ex;                 // sqrt(m) * 2^e in [128, 256)
    return e < -60 ? -60 : (e > 60 ? 60 : e);
}

// order-preserving float -> uint32 (so that (value, index) packs into one u64 key)
__device__ __forceinline__ uint32_t gmd_flip(float v) {
    uint32_t b = __float_as_uint(v);
    return b ^ ((b & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float gmd_unflip(uint32_t u) {
    uint32_t b = u ^ ((u & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu);
    return __uint_as_float(b);
}
__device__ __forceinline__ uint64_t gmd_key(float v, uint32_t idx) {
    return ((uint64_t)gmd_flip(v) << 32) | (uint64_t)idx;
}
