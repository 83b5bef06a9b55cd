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
