// common.cuh -- shared helpers for the b200id kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "../../include/b200id.h"

#ifndef B200ID_HD
#define B200ID_HD __host__ __device__ __forceinline__
#endif

namespace b200id {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
int  check_cuda(cudaError_t e, const char* what);
int  sm_count();

#define B200ID_REQUIRE(cond, code, ...)                         \
    do {                                                        \
        if (!(cond)) {                                          \
            b200id::set_error(__VA_ARGS__);                     \
            return (code);                                      \
        }                                                       \
    } while (0)

#define B200ID_LAUNCH_CHECK(what)                                                   \
    do {                                                                            \
        int _rc = b200id::check_cuda(cudaGetLastError(), what);                     \
        if (_rc != B200ID_OK) return _rc;                                           \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------- warp / block reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum: fixed shuffle tree inside each warp, then warp 0 folds the per-warp
// values in warp order.  `scratch` needs blockDim.x/32 doubles.  Result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        r = (lane < nw) ? scratch[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

}  // namespace b200id
