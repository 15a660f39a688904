// common.cuh -- shared helpers for libgravinv_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "gravinv_b200.h"

namespace gi {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
int sm_count();

#define GI_CUDA(call)                                                      \
    do {                                                                   \
        cudaError_t e__ = (call);                                          \
        if (e__ != cudaSuccess) return gi::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define GI_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            gi::set_error(__VA_ARGS__);  \
            return GI_ERR_INVALID;       \
        }                                \
    } while (0)

#define GI_LAUNCH_CHECK() GI_CUDA(cudaGetLastError())

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ------------------------------------------------------------------------
// 256-bit streaming load of 4 consecutive doubles (LDG.E.256 on sm_100a); G is read once per
// pass, so keep it out of L1 and mark it evict-first in L2 so the small reused vectors (x, r)
// stay resident.
__device__ __forceinline__ void ldg_stream4(const double *p, double &a, double &b, double &c,
                                            double &d) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(a), "=d"(b), "=d"(c), "=d"(d)
                 : "l"(p));
}
__device__ __forceinline__ void ldg4(const double *p, double &a, double &b, double &c, double &d) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(a), "=d"(b), "=d"(c), "=d"(d)
                 : "l"(p));
}
// cache-global variant for buffers that are rewritten in place by the same kernel
__device__ __forceinline__ void ldg4_cg(const double *p, double &a, double &b, double &c, double &d) {
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(a), "=d"(b), "=d"(c), "=d"(d)
                 : "l"(p));
}
__device__ __forceinline__ void stg4(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d)
                 : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block-wide sum (fixed shuffle tree + fixed-order cross-warp add).
// `scratch` must hold >= 32 doubles. Result valid in every thread.
__device__ __forceinline__ double block_sum(double v, double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nwarp; ++w) t += scratch[w];
    return t;
}

}  // namespace gi
