// Shared host/device helpers for libncde_b200.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "ncde_b200.h"

namespace ncde {

void set_error(const char* fmt, ...);

#define NCDE_CUDA_OK(expr)                                                                              \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            ncde::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return NCDE_ERR_CUDA;                                                                       \
        }                                                                                               \
    } while (0)

#define NCDE_REQUIRE(cond, code, ...)     \
    do {                                  \
        if (!(cond)) {                    \
            ncde::set_error(__VA_ARGS__); \
            return (code);                \
        }                                 \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// Number of knots strictly below t, minus one, clamped to a valid piece: torch.bucketize(t, knots) - 1 clamped to
// [0, K-2].  At an exact knot k >= 1 this selects the LEFT piece k-1.
template <typename T>
__device__ __forceinline__ int knot_index(const T* __restrict__ knots, int K, T t) {
    int lo = 0, hi = K;  // first position with knots[pos] >= t
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (knots[mid] < t) lo = mid + 1; else hi = mid;
    }
    int idx = lo - 1;
    idx = idx < 0 ? 0 : idx;
    idx = idx > K - 2 ? K - 2 : idx;
    return idx;
}

}  // namespace ncde
