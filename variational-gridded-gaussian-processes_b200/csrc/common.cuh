// Shared helpers for libvggp (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "../../include/vggp.h"

namespace vggp {

typedef int64_t i64;

extern thread_local char g_err[512];
extern unsigned long long g_launches;   // kernels launched by this library (host-side counter)

inline int fail(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

#define VGGP_CUDA(expr)                                                                            \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            snprintf(vggp::g_err, sizeof(vggp::g_err), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, \
                     cudaGetErrorString(_e));                                                      \
            return (int)_e;                                                                        \
        }                                                                                          \
    } while (0)

#define VGGP_LAUNCH_CHECK()            \
    do {                               \
        ++vggp::g_launches;            \
        VGGP_CUDA(cudaGetLastError()); \
    } while (0)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the whole block; result valid in thread 0.  `red` = shared scratch of >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();                 // protect `red` against a previous use
    if (lane == 0) red[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = (lane < nw) ? red[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

inline int ceil_div(i64 a, i64 b) { return (int)((a + b - 1) / b); }

}  // namespace vggp
