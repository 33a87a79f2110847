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

// Every entry point runs on the device that owns its buffers, whatever the caller's current device is, and restores the
// caller's device on return (a model on cuda:1 called while cuda:0 is current must not launch on cuda:0).
struct DeviceGuard {
    int prev = -1, target = -1;
    explicit DeviceGuard(const void* device_ptr) {
        cudaPointerAttributes at;
        if (device_ptr && cudaPointerGetAttributes(&at, device_ptr) == cudaSuccess && at.type == cudaMemoryTypeDevice) {
            target = at.device;
            if (cudaGetDevice(&prev) == cudaSuccess && prev != target) cudaSetDevice(target);
            else prev = -1;
        } else {
            cudaGetLastError();   // a host pointer on an old driver reports an error: not ours to surface
        }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

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

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl may begin (prologue: shared-memory carve-up,
// TMEM allocation, staging of the packed weights) while its predecessor on the stream is still draining; it must
// execute pdl_wait() before touching anything a previous kernel produced.  pdl_trigger() lets the successor start
// its own prologue early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace ncde
