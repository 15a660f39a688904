// tess_math.cuh -- the tesseroid building blocks of gravmag/_tesseroid_numba.py shared by the gz
// assembly (assemble.cu) and the other fields (tess_fields.cu): GLQ node scaling (scale_nodes
// :75-91), the gz leaf kernel (:207-222) and the split decision (distance_size :94-111, divisions
// :135-157), each separated into a cell-only part and a per-observation part.
#pragma once
#include <math.h>

namespace gi {

constexpr double kEarthRadius = 6378137.0;  // constants.py:44
constexpr int kStackSize = 100;             // gravmag/tesseroid.py:79
constexpr double kNodeLo = -0.577350269189625731058868041146;
constexpr double kNodeHi = 0.577350269189625731058868041146;
constexpr double kNpPi = 3.141592653589793;  // np.pi

struct TessCell {
    double w, e, s, n, top, bottom;
};

// Everything that depends on the cell only is separated from the per-observation part, so that
// the thread owning a column evaluates it once and reuses it for all its observation rows (the
// values and the operation order per pair are unchanged -> same bits).
struct TessLeafC {
    double lonc[2], sinlatc[2], coslatc[2], rc[2], rck2[2], kappa[2][2], scale;
};

__device__ __forceinline__ void tess_leaf_consts(const TessCell &c, TessLeafC &L) {
    // scale_nodes (_tesseroid_numba.py:75-91) + the cell-only factors of kernelz (:207-222)
    const double d2r = kNpPi / 180;
    const double dlon = __dmul_rn(d2r, __dsub_rn(c.e, c.w));
    const double dlat = __dmul_rn(d2r, __dsub_rn(c.n, c.s));
    const double dr = __dsub_rn(c.top, c.bottom);
    const double nodes[2] = {kNodeLo, kNodeHi};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        L.lonc[i] = __dadd_rn(__dmul_rn(__dmul_rn(0.5, dlon), nodes[i]),
                              __dmul_rn(__dmul_rn(d2r, 0.5), __dadd_rn(c.e, c.w)));
        const double latc = __dadd_rn(__dmul_rn(__dmul_rn(0.5, dlat), nodes[i]),
                                      __dmul_rn(__dmul_rn(d2r, 0.5), __dadd_rn(c.n, c.s)));
        L.sinlatc[i] = sin(latc);
        L.coslatc[i] = cos(latc);
        L.rc[i] = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(0.5, dr), nodes[i]),
                                      __dmul_rn(0.5, __dadd_rn(c.top, c.bottom))),
                            kEarthRadius);
    }
    L.scale = __dmul_rn(__dmul_rn(__dmul_rn(dlon, dlat), dr), 0.125);
#pragma unroll
    for (int k = 0; k < 2; ++k) L.rck2[k] = __dmul_rn(L.rc[k], L.rc[k]);
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 2; ++k) L.kappa[j][k] = __dmul_rn(L.rck2[k], L.coslatc[j]);
}

__device__ __forceinline__ double tess_leaf_eval(double lon, double coslat, double sinlat, double radius,
                                                 const TessLeafC &L) {
    const double r_sqr = __dmul_rn(radius, radius);
    double result = 0.0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double coslon = cos(__dsub_rn(lon, L.lonc[i]));
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double cospsi = __dadd_rn(__dmul_rn(sinlat, L.sinlatc[j]),
                                            __dmul_rn(__dmul_rn(coslat, L.coslatc[j]), coslon));
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const double l_sqr = __dsub_rn(
                    __dadd_rn(r_sqr, L.rck2[k]),
                    __dmul_rn(__dmul_rn(__dmul_rn(2.0, radius), L.rc[k]), cospsi));
                const double num = __dmul_rn(L.kappa[j][k], __dsub_rn(__dmul_rn(L.rc[k], cospsi), radius));
                // l_sqr**1.5 ; pow(x, 1.5) == x*sqrt(x) to within 1 ulp
                result = __dadd_rn(result, __ddiv_rn(num, __dmul_rn(l_sqr, __dsqrt_rn(l_sqr))));
            }
        }
    }
    return __dmul_rn(L.scale, -result);
}

__device__ __forceinline__ double tess_leaf(double lon, double coslat, double sinlat, double radius,
                                            const TessCell &c) {
    TessLeafC L;
    tess_leaf_consts(c, L);
    return tess_leaf_eval(lon, coslat, sinlat, radius, L);
}

// Split decision of one cell (distance_size :94-111 + divisions :135-157), cell-only part
struct TessDivC {
    double rt2, rt, lont, sinlatt, coslatt, rLlon, rLlat, rLr;
    bool lon_small, lat_small, r_small;
};

__device__ __forceinline__ void tess_div_consts(const TessCell &c, double ratio, TessDivC &D) {
    const double d2r = kNpPi / 180;
    D.rt = __dadd_rn(__dmul_rn(0.5, __dadd_rn(c.top, c.bottom)), kEarthRadius);
    D.rt2 = __dmul_rn(D.rt, D.rt);
    D.lont = __dmul_rn(__dmul_rn(d2r, 0.5), __dadd_rn(c.w, c.e));
    const double latt = __dmul_rn(__dmul_rn(d2r, 0.5), __dadd_rn(c.s, c.n));
    D.sinlatt = sin(latt);
    D.coslatt = cos(latt);
    const double rtop = __dadd_rn(c.top, kEarthRadius);
    const double Llon = __dmul_rn(
        rtop, acos(__dadd_rn(__dmul_rn(D.sinlatt, D.sinlatt),
                             __dmul_rn(__dmul_rn(D.coslatt, D.coslatt),
                                       cos(__dmul_rn(d2r, __dsub_rn(c.e, c.w)))))));
    const double dn = __dmul_rn(d2r, c.n), ds = __dmul_rn(d2r, c.s);
    const double Llat = __dmul_rn(
        rtop, acos(__dadd_rn(__dmul_rn(sin(dn), sin(ds)), __dmul_rn(cos(dn), cos(ds)))));
    const double Lr = __dsub_rn(c.top, c.bottom);
    D.rLlon = __dmul_rn(ratio, Llon);
    D.rLlat = __dmul_rn(ratio, Llat);
    D.rLr = __dmul_rn(ratio, Lr);
    D.lon_small = Llon <= 0.1;
    D.lat_small = Llat <= 0.1;
    D.r_small = Lr <= 1e3;
}

// Returns nlon | nlat<<2 | nr<<4, err in *err (0 or -1).
__device__ __forceinline__ int tess_div_eval(double lon, double coslat, double sinlat, double radius,
                                             const TessDivC &D, int *err) {
    const double cospsi = __dadd_rn(__dmul_rn(sinlat, D.sinlatt),
                                    __dmul_rn(__dmul_rn(coslat, D.coslatt), cos(__dsub_rn(lon, D.lont))));
    const double distance = __dsqrt_rn(
        __dsub_rn(__dadd_rn(__dmul_rn(radius, radius), D.rt2),
                  __dmul_rn(__dmul_rn(__dmul_rn(2.0, radius), D.rt), cospsi)));
    int nlon = 1, nlat = 1, nr = 1, e = 0;
    if (distance <= D.rLlon) {
        if (D.lon_small) e = -1; else nlon = 2;
    }
    if (distance <= D.rLlat) {
        if (D.lat_small) e = -1; else nlat = 2;
    }
    if (distance <= D.rLr) {
        if (D.r_small) e = -1; else nr = 2;
    }
    *err = e;
    return nlon | (nlat << 2) | (nr << 4);
}

__device__ __forceinline__ int tess_divisions(double lon, double coslat, double sinlat, double radius,
                                              const TessCell &c, double ratio, int *err) {
    TessDivC D;
    tess_div_consts(c, ratio, D);
    return tess_div_eval(lon, coslat, sinlat, radius, D, err);
}

}  // namespace gi
