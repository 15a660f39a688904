// prism_math.cuh -- the guarded transcendental helpers of gravmag/_prism.pyx:16-34, shared by the gz
// assembly (assemble.cu) and the other prism fields (fields.cu).
#pragma once
#include <math.h>

namespace gi {

// _prism.pyx:21 spells pi with more digits than a double holds; this is the same binary64.
__device__ __constant__ const double kPi = 3.1415926535897931159979634685441851615906;

__device__ __forceinline__ double prism_safe_atan2(double y, double x) {
    // _prism.pyx:16-26
    if (y == 0.0) return 0.0;
    double a = atan2(y, x);
    if (x < 0.0) {
        if (y > 0.0) a = __dsub_rn(a, kPi);
        else if (y < 0.0) a = __dadd_rn(a, kPi);
    }
    return a;
}

__device__ __forceinline__ double prism_safe_log(double x) {
    // _prism.pyx:28-34
    return (x == 0.0) ? 0.0 : log(x);
}

}  // namespace gi
