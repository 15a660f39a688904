// tess_fields.cu -- sensitivity matrices of the tesseroid fields other than gz:
// potential, gx, gy, gxx, gxy, gxz, gyy, gyz, gzz by 2x2x2 Gauss-Legendre quadrature with the
// reference's adaptive LIFO subdivision.
//
// Reference: gravmag/_tesseroid_numba.py -- engine :25-72, scale_nodes :75-91, distance_size :94-111,
// split :114-132, divisions :135-157, kernels kernelV :160-173, kernelx :176-189, kernely :192-205,
// kernelz :208-223, kernelxx :226-239, kernelxy :242-257, kernelxz :260-274, kernelyy :277-292,
// kernelyz :295-311, kernelzz :314-328; called per tesseroid by gravmag/tesseroid.py:189-232 with the
// distance-size ratios RATIO_V = 1 / RATIO_G = 1.6 / RATIO_GG = 8 (:76-78).
//
// Same layout as tess_gz_kernel: a thread owns one column (cell) and strides over observation rows;
// the split test of the un-split cell is hoisted out of the row loop.  Operation order follows the
// reference with explicit round-to-nearest intrinsics; x**1.5 and x**2.5 are x*sqrt(x) and
// x*x*sqrt(x) (within an ulp of pow).
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "tess_math.cuh"

namespace gi {
namespace {

#define M2(a, b) __dmul_rn(a, b)
#define A2(a, b) __dadd_rn(a, b)
#define S2(a, b) __dsub_rn(a, b)
#define D2(a, b) __ddiv_rn(a, b)

template <int FIELD>
__device__ __forceinline__ double tess_field_leaf(double lon, double coslat, double sinlat, double radius,
                                                  const TessCell &c) {
    TessLeafC L;
    tess_leaf_consts(c, L);
    if (FIELD == GI_FIELD_GZ) return tess_leaf_eval(lon, coslat, sinlat, radius, L);
    const double r_sqr = M2(radius, radius);
    double result = 0.0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double coslon = cos(S2(lon, L.lonc[i]));
        const double sinlon = sin(S2(L.lonc[i], lon));
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double kphi = S2(M2(coslat, L.sinlatc[j]), M2(M2(sinlat, L.coslatc[j]), coslon));
            const double cospsi = A2(M2(sinlat, L.sinlatc[j]), M2(M2(coslat, L.coslatc[j]), coslon));
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const double rc = L.rc[k], rc_sqr = L.rck2[k], kappa = L.kappa[j][k];
                const double l_sqr = S2(A2(r_sqr, rc_sqr), M2(M2(M2(2.0, radius), rc), cospsi));
                const double rt = __dsqrt_rn(l_sqr);
                const double l_3 = M2(l_sqr, rt), l_5 = M2(M2(l_sqr, l_sqr), rt);
                const double deltay = M2(M2(rc, L.coslatc[j]), sinlon);
                const double deltaz = S2(M2(rc, cospsi), radius);
                double term;
                switch (FIELD) {
                    case GI_FIELD_POTENTIAL: term = D2(kappa, rt); break;
                    case GI_FIELD_GX: term = D2(M2(M2(kappa, rc), kphi), l_3); break;
                    case GI_FIELD_GY: term = M2(kappa, D2(deltay, l_3)); break;
                    case GI_FIELD_GXX: {
                        const double t = M2(rc, kphi);
                        term = D2(M2(kappa, S2(M2(3.0, M2(t, t)), l_sqr)), l_5);
                        break;
                    }
                    case GI_FIELD_GXY:
                        term = D2(M2(M2(M2(M2(M2(kappa, 3.0), rc_sqr), kphi), L.coslatc[j]), sinlon), l_5);
                        break;
                    case GI_FIELD_GXZ: term = D2(M2(M2(M2(M2(kappa, 3.0), rc), kphi), deltaz), l_5); break;
                    case GI_FIELD_GYY: term = D2(M2(kappa, S2(M2(3.0, M2(deltay, deltay)), l_sqr)), l_5); break;
                    case GI_FIELD_GYZ: term = D2(M2(M2(M2(kappa, 3.0), deltay), deltaz), l_5); break;
                    default: term = D2(M2(kappa, S2(M2(3.0, M2(deltaz, deltaz)), l_sqr)), l_5); break;
                }
                result = A2(result, term);
            }
        }
    }
    return M2(L.scale, result);
}

#undef M2
#undef A2
#undef S2
#undef D2

constexpr int kTfThreads = 128;

template <int FIELD>
__global__ void __launch_bounds__(kTfThreads)
tess_field_kernel(const double *__restrict__ lon, const double *__restrict__ sinlat,
                  const double *__restrict__ coslat, const double *__restrict__ radius, int64_t nrows,
                  const double *__restrict__ bounds, int64_t M, double ratio, double scale1, double scale2,
                  double *__restrict__ G, int64_t ld, int32_t *__restrict__ status) {
    const int64_t col = (int64_t)blockIdx.x * kTfThreads + threadIdx.x;
    if (col >= ld) return;
    const bool live = col < M;
    TessCell root;
    TessDivC rootD;
    if (live) {
        const double *b = bounds + 6 * col;
        root.w = b[0]; root.e = b[1]; root.s = b[2]; root.n = b[3]; root.top = b[4]; root.bottom = b[5];
        tess_div_consts(root, ratio, rootD);
    }
    TessCell stack[kStackSize];  // local memory; only touched when a cell subdivides
    int errsum = 0;
    bool overflow = false;
    for (int64_t row = blockIdx.y; row < nrows; row += gridDim.y) {
        double acc = 0.0;
        if (live) {
            const double olon = __ldg(lon + row), osin = __ldg(sinlat + row), ocos = __ldg(coslat + row),
                         orad = __ldg(radius + row);
            // engine (_tesseroid_numba.py:32-71): LIFO stack, children pushed lon-major
            int top = -1, err;
            TessCell cur = root;
            int div = tess_div_eval(olon, ocos, osin, orad, rootD, &err);
            errsum += err;
            bool have = true;
            while (true) {
                if (!have) {
                    if (top < 0) break;
                    cur = stack[top--];
                    div = tess_divisions(olon, ocos, osin, orad, cur, ratio, &err);
                    errsum += err;
                }
                have = false;
                const int nlon = div & 3, nlat = (div >> 2) & 3, nr = (div >> 4) & 3;
                const int ncell = nlon * nlat * nr;
                if (ncell > 1) {
                    if (ncell + (top + 1) > kStackSize) {
                        overflow = true;
                        acc = nan("");
                        break;
                    }
                    const double dlon = __ddiv_rn(__dsub_rn(cur.e, cur.w), (double)nlon);
                    const double dlat = __ddiv_rn(__dsub_rn(cur.n, cur.s), (double)nlat);
                    const double dr = __ddiv_rn(__dsub_rn(cur.top, cur.bottom), (double)nr);
                    for (int i = 0; i < nlon; ++i)
                        for (int j = 0; j < nlat; ++j)
                            for (int k = 0; k < nr; ++k) {
                                TessCell c;
                                c.w = __dadd_rn(cur.w, __dmul_rn((double)i, dlon));
                                c.e = __dadd_rn(cur.w, __dmul_rn((double)(i + 1), dlon));
                                c.s = __dadd_rn(cur.s, __dmul_rn((double)j, dlat));
                                c.n = __dadd_rn(cur.s, __dmul_rn((double)(j + 1), dlat));
                                c.top = __dadd_rn(cur.bottom, __dmul_rn((double)(k + 1), dr));
                                c.bottom = __dadd_rn(cur.bottom, __dmul_rn((double)k, dr));
                                stack[++top] = c;
                            }
                } else {
                    acc = __dadd_rn(acc, tess_field_leaf<FIELD>(olon, ocos, osin, orad, cur));
                }
            }
            acc = __dmul_rn(__dmul_rn(acc, scale1), scale2);  // tesseroid.py:375-507
        }
        G[row * ld + col] = acc;
    }
    if (errsum != 0) atomicAdd(status, errsum);
    if (overflow) atomicExch(status + 1, 1);
}

}  // namespace
}  // namespace gi

using namespace gi;

extern "C" int gi_tess_field_assemble(int32_t field, const double *lon, const double *sinlat,
                                      const double *coslat, const double *radius, int64_t nrows,
                                      const double *bounds, int64_t M, double ratio, double scale1,
                                      double scale2, double *G, int64_t ld, int32_t *status, void *stream) {
    GI_REQUIRE(field >= GI_FIELD_POTENTIAL && field <= GI_FIELD_GZZ, "gi_tess_field_assemble: unknown field");
    GI_REQUIRE(nrows >= 0 && M >= 0 && ld >= M && ld % 4 == 0, "gi_tess_field_assemble: bad shape");
    GI_REQUIRE(ratio > 0, "gi_tess_field_assemble: ratio must be > 0");
    if (nrows == 0 || ld == 0) return GI_OK;
    GI_REQUIRE(lon && sinlat && coslat && radius && G && status && (bounds || M == 0),
               "gi_tess_field_assemble: null pointer");
    const int64_t xblocks = ceil_div(ld, kTfThreads);
    int64_t y = (16LL * 8 * sm_count()) / xblocks + 1;
    y = std::max<int64_t>(1, std::min<int64_t>(y, std::min<int64_t>(nrows, 65535)));
    dim3 grid((unsigned)xblocks, (unsigned)y);
    cudaStream_t s = (cudaStream_t)stream;
#define GI_TF(F)                                                                                   \
    case F:                                                                                        \
        tess_field_kernel<F><<<grid, kTfThreads, 0, s>>>(lon, sinlat, coslat, radius, nrows, bounds, M, \
                                                         ratio, scale1, scale2, G, ld, status);    \
        break;
    switch (field) {
        GI_TF(GI_FIELD_POTENTIAL) GI_TF(GI_FIELD_GX) GI_TF(GI_FIELD_GY) GI_TF(GI_FIELD_GZ)
        GI_TF(GI_FIELD_GXX) GI_TF(GI_FIELD_GXY) GI_TF(GI_FIELD_GXZ) GI_TF(GI_FIELD_GYY)
        GI_TF(GI_FIELD_GYZ) GI_TF(GI_FIELD_GZZ)
    }
#undef GI_TF
    GI_LAUNCH_CHECK();
    return GI_OK;
}
