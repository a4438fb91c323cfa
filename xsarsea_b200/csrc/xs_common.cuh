// Shared helpers for libxsarsea_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/xsarsea_b200.h"

namespace xs {

void set_error(const char *fmt, ...);
void keep_async_pool();
extern std::atomic<int64_t> g_launches;  // concurrent host threads launch through the same library

inline int check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return XS_E_CUDA;
    }
    return XS_OK;
}

#define XS_CUDA(call)                                 \
    do {                                              \
        int _rc = xs::check((call), #call);           \
        if (_rc != XS_OK) return _rc;                 \
    } while (0)

// Every kernel launch of the library goes through this so that xs_launch_count() is a true count.
#define XS_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
    do {                                                                  \
        kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__); \
        ++xs::g_launches;                                                 \
        XS_CUDA(cudaGetLastError());                                      \
    } while (0)

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// numpy.deg2rad: x * (pi/180)
__device__ __forceinline__ double deg2rad(double x) { return x * (3.14159265358979323846 / 180.0); }

}  // namespace xs
