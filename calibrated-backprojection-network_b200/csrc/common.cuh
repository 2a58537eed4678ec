// Common helpers for the kbnet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/kbnet_b200.h"

namespace kbnet {

// ---- thread-local error message (no global mutable state shared between threads) ---------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);  // statistics for kbnet_launch_count()
void count_launch_v2();         // ... and for kbnet_tma_engine_launch_count()

#define KB_CHECK_ARG(cond, ...)                         \
    do {                                                \
        if (!(cond)) {                                  \
            ::kbnet::set_error(__VA_ARGS__);            \
            return KBNET_E_BADARG;                      \
        }                                               \
    } while (0)

#define KB_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::kbnet::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (int)_e;                                                                    \
        }                                                                                      \
    } while (0)

// ---- small device helpers ----------------------------------------------------------------------
__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : v * slope; }

// NaN-propagating min / max (ATen's max_pool2d keeps a NaN once it has seen one).
__device__ __forceinline__ float min_nan(float a, float b) {
    float d;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ float max_nan(float a, float b) {
    float d;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}

// torch 'nearest' source index: min(floor(dst * in/out), in - 1)  (ATen nearest_neighbor_compute_source_index)
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
    int s = (int)floorf((float)dst * scale);
    return s < in_size - 1 ? s : in_size - 1;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace kbnet
