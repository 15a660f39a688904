// fields.cu -- sensitivity matrices of the remaining right-rectangular-prism fields (SURVEY.md
// 8(f3)): potential, gx, gy, gz, the six gravity-gradient components, the total-field magnetic
// anomaly and the rows of the second-derivative tensor applied to a vector (bx / by / bz).
//
// Reference: gravmag/_prism.pyx -- corner kernels :36-70 (kernelpot, kernelx..kernelzz), the per-field
// drivers potential :484-509, gx :206-232, gy :235-261, gz :265-290, gxx :294-320, gxy :324-354,
// gxz :358-388, gyy :392-417, gyz :421-451, gzz :455-480, tf :72-112, bx/by/bz :116-202 -- called per
// prism by gravmag/prism.py:102-732 (one kernel2d column per prism, scaled after the 8-corner sum).
//
// Same layout and arithmetic discipline as prism_gz_kernel (assemble.cu): consecutive threads own
// consecutive columns of one observation row (256 B coalesced stores), the reference's operation
// order is kept with explicit round-to-nearest intrinsics (no FMA contraction), corners are visited
// k (z) outer, j (y), i (x) inner with x = [x2, x1] etc. and the sign (-1)^(i+j+k).
#include <math.h>

#include "common.cuh"
#include "prism_math.cuh"

namespace gi {

namespace {

struct Corner {
    double x, y, z, r;
};

__device__ __forceinline__ double sq3(double a, double b, double c) {
    // sqrt(a**2 + b**2 + c**2), left to right
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c)));
}

// second derivatives of 1/r integrated over the corner (_prism.pyx:52-70)
__device__ __forceinline__ double k_xx(const Corner &c) { return -prism_safe_atan2(__dmul_rn(c.z, c.y), __dmul_rn(c.x, c.r)); }
__device__ __forceinline__ double k_xy(const Corner &c) { return prism_safe_log(__dadd_rn(c.z, c.r)); }
__device__ __forceinline__ double k_xz(const Corner &c) { return prism_safe_log(__dadd_rn(c.y, c.r)); }
__device__ __forceinline__ double k_yy(const Corner &c) { return -prism_safe_atan2(__dmul_rn(c.z, c.x), __dmul_rn(c.y, c.r)); }
__device__ __forceinline__ double k_yz(const Corner &c) { return prism_safe_log(__dadd_rn(c.x, c.r)); }
__device__ __forceinline__ double k_zz(const Corner &c) { return -prism_safe_atan2(__dmul_rn(c.x, c.y), __dmul_rn(c.z, c.r)); }

__device__ __forceinline__ double dot3(double a0, double b0, double a1, double b1, double a2, double b2) {
    // (a0*b0 + a1*b1 + a2*b2), left to right
    return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2));
}

template <int FIELD>
__device__ __forceinline__ double corner_value(double dx, double dy, double dz, double ex, double ey,
                                               double ez, double v0, double v1, double v2) {
    Corner c{dx, dy, dz, 0.0};
    // the three mixed components move the computation point off the singular edge
    // (_prism.pyx:345-350, 380-385, 442-447); ex/ey/ez = 0.00001 * (cell size)
    if (FIELD == GI_FIELD_GXY && dx == 0.0 && dy == 0.0 && dz < 0.0) c.r = sq3(ex, ey, dz);
    else if (FIELD == GI_FIELD_GXZ && dx == 0.0 && dz == 0.0 && dy < 0.0) c.r = sq3(ex, ez, dy);
    else if (FIELD == GI_FIELD_GYZ && dy == 0.0 && dz == 0.0 && dx < 0.0) c.r = sq3(ey, ez, dx);
    else c.r = sq3(dx, dy, dz);
    const double x = c.x, y = c.y, z = c.z, r = c.r;
    switch (FIELD) {
        case GI_FIELD_POTENTIAL: {  // kernelpot, _prism.pyx:36-39
            const double t1 = __dmul_rn(__dmul_rn(x, y), prism_safe_log(__dadd_rn(z, r)));
            const double t2 = __dmul_rn(__dmul_rn(y, z), prism_safe_log(__dadd_rn(x, r)));
            const double t3 = __dmul_rn(__dmul_rn(x, z), prism_safe_log(__dadd_rn(y, r)));
            const double t4 = __dmul_rn(__dmul_rn(0.5, __dmul_rn(x, x)),
                                        prism_safe_atan2(__dmul_rn(z, y), __dmul_rn(x, r)));
            const double t5 = __dmul_rn(__dmul_rn(0.5, __dmul_rn(y, y)),
                                        prism_safe_atan2(__dmul_rn(z, x), __dmul_rn(y, r)));
            const double t6 = __dmul_rn(__dmul_rn(0.5, __dmul_rn(z, z)),
                                        prism_safe_atan2(__dmul_rn(x, y), __dmul_rn(z, r)));
            return __dsub_rn(__dsub_rn(__dsub_rn(__dadd_rn(__dadd_rn(t1, t2), t3), t4), t5), t6);
        }
        case GI_FIELD_GX: {  // kernelx, :43-44
            const double t1 = __dmul_rn(y, prism_safe_log(__dadd_rn(z, r)));
            const double t2 = __dmul_rn(z, prism_safe_log(__dadd_rn(y, r)));
            const double t3 = __dmul_rn(x, prism_safe_atan2(__dmul_rn(z, y), __dmul_rn(x, r)));
            return -__dsub_rn(__dadd_rn(t1, t2), t3);
        }
        case GI_FIELD_GY: {  // kernely, :46-47
            const double t1 = __dmul_rn(z, prism_safe_log(__dadd_rn(x, r)));
            const double t2 = __dmul_rn(x, prism_safe_log(__dadd_rn(z, r)));
            const double t3 = __dmul_rn(y, prism_safe_atan2(__dmul_rn(x, z), __dmul_rn(y, r)));
            return -__dsub_rn(__dadd_rn(t1, t2), t3);
        }
        case GI_FIELD_GZ: {  // kernelz, :49-50
            const double t1 = __dmul_rn(x, prism_safe_log(__dadd_rn(y, r)));
            const double t2 = __dmul_rn(y, prism_safe_log(__dadd_rn(x, r)));
            const double t3 = __dmul_rn(z, prism_safe_atan2(__dmul_rn(x, y), __dmul_rn(z, r)));
            return -__dsub_rn(__dadd_rn(t1, t2), t3);
        }
        case GI_FIELD_GXX: return k_xx(c);
        case GI_FIELD_GXY: return k_xy(c);
        case GI_FIELD_GXZ: return k_xz(c);
        case GI_FIELD_GYY: return k_yy(c);
        case GI_FIELD_GYZ: return k_yz(c);
        case GI_FIELD_GZZ: return k_zz(c);
        case GI_FIELD_TF: {  // _prism.pyx:94-111 with (v0, v1, v2) = (fx, fy, fz): kernelk = f . (V f)
            const double a = k_xx(c), b = k_xy(c), d = k_xz(c), e = k_yy(c), f = k_yz(c), g = k_zz(c);
            const double bxk = dot3(a, v0, b, v1, d, v2);
            const double byk = dot3(b, v0, e, v1, f, v2);
            const double bzk = dot3(d, v0, f, v1, g, v2);
            return dot3(v0, bxk, v1, byk, v2, bzk);
        }
        case GI_FIELD_VX:  // _prism.pyx:136-140 (bx): row x of V times (v0, v1, v2)
            return dot3(k_xx(c), v0, k_xy(c), v1, k_xz(c), v2);
        case GI_FIELD_VY:  // :166-170 (by)
            return dot3(k_xy(c), v0, k_yy(c), v1, k_yz(c), v2);
        default:           // :196-200 (bz)
            return dot3(k_xz(c), v0, k_yz(c), v1, k_zz(c), v2);
    }
}

constexpr int kFldThreads = 128;  // columns per CTA
constexpr int kFldRows = 8;       // observation rows per CTA tile

template <int FIELD>
__global__ void __launch_bounds__(kFldThreads)
prism_field_kernel(const double *__restrict__ xp, const double *__restrict__ yp, const double *__restrict__ zp,
                   int64_t nrows, const double *__restrict__ bounds, int64_t M, double scale, double v0,
                   double v1, double v2, double *__restrict__ G, int64_t ld) {
    const int64_t col = (int64_t)blockIdx.x * kFldThreads + threadIdx.x;
    if (col >= ld) return;
    const bool live = col < M;
    double bx[2] = {0, 0}, by[2] = {0, 0}, bz[2] = {0, 0}, ex = 0, ey = 0, ez = 0;
    if (live) {
        const double *b = bounds + 6 * col;
        bx[0] = b[1]; bx[1] = b[0];
        by[0] = b[3]; by[1] = b[2];
        bz[0] = b[5]; bz[1] = b[4];
        ex = __dmul_rn(0.00001, __dsub_rn(b[1], b[0]));
        ey = __dmul_rn(0.00001, __dsub_rn(b[3], b[2]));
        ez = __dmul_rn(0.00001, __dsub_rn(b[5], b[4]));
    }
    for (int64_t tile = blockIdx.y; tile * kFldRows < nrows; tile += gridDim.y) {
        const int64_t r0 = tile * kFldRows;
        const int nr = (int)min((int64_t)kFldRows, nrows - r0);
        for (int rr = 0; rr < nr; ++rr) {
            const int64_t row = r0 + rr;
            double acc = 0.0;
            if (live) {
                const double ox = __ldg(xp + row), oy = __ldg(yp + row), oz = __ldg(zp + row);
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    const int k = c >> 2, j = (c >> 1) & 1, i = c & 1;
                    const double dz = __dsub_rn(bz[k], oz);
                    const double dy = __dsub_rn(by[j], oy);
                    const double dx = __dsub_rn(bx[i], ox);
                    const double kern = corner_value<FIELD>(dx, dy, dz, ex, ey, ez, v0, v1, v2);
                    acc = __dadd_rn(acc, ((i + j + k) & 1) ? -kern : kern);
                }
                acc = __dmul_rn(acc, scale);
            }
            G[row * ld + col] = acc;
        }
    }
}

}  // namespace
}  // namespace gi

using namespace gi;

extern "C" int gi_prism_field_assemble(int32_t field, const double *xp, const double *yp, const double *zp,
                                       int64_t nrows, const double *bounds, int64_t M, double scale,
                                       const double *vec3_host, double *G, int64_t ld, void *stream) {
    GI_REQUIRE(field >= GI_FIELD_POTENTIAL && field <= GI_FIELD_VZ, "gi_prism_field_assemble: unknown field");
    GI_REQUIRE(nrows >= 0 && M >= 0 && ld >= M && ld % 4 == 0, "gi_prism_field_assemble: bad shape");
    if (nrows == 0 || ld == 0) return GI_OK;
    GI_REQUIRE(xp && yp && zp && G && (bounds || M == 0), "gi_prism_field_assemble: null pointer");
    GI_REQUIRE(field < GI_FIELD_TF || vec3_host, "gi_prism_field_assemble: this field needs a vector");
    const double v0 = vec3_host ? vec3_host[0] : 0.0, v1 = vec3_host ? vec3_host[1] : 0.0,
                 v2 = vec3_host ? vec3_host[2] : 0.0;
    dim3 grid((unsigned)ceil_div(ld, kFldThreads), (unsigned)std::min<int64_t>(65535, ceil_div(nrows, kFldRows)));
    cudaStream_t s = (cudaStream_t)stream;
#define GI_FLD(F)                                                                                  \
    case F:                                                                                        \
        prism_field_kernel<F><<<grid, kFldThreads, 0, s>>>(xp, yp, zp, nrows, bounds, M, scale, v0, v1, \
                                                           v2, G, ld);                             \
        break;
    switch (field) {
        GI_FLD(GI_FIELD_POTENTIAL) GI_FLD(GI_FIELD_GX) GI_FLD(GI_FIELD_GY) GI_FLD(GI_FIELD_GZ)
        GI_FLD(GI_FIELD_GXX) GI_FLD(GI_FIELD_GXY) GI_FLD(GI_FIELD_GXZ) GI_FLD(GI_FIELD_GYY)
        GI_FLD(GI_FIELD_GYZ) GI_FLD(GI_FIELD_GZZ) GI_FLD(GI_FIELD_TF) GI_FLD(GI_FIELD_VX)
        GI_FLD(GI_FIELD_VY) GI_FLD(GI_FIELD_VZ)
    }
#undef GI_FLD
    GI_LAUNCH_CHECK();
    return GI_OK;
}
