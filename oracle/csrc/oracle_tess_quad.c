/* TEST INFRASTRUCTURE ONLY -- the tesseroid gz kernel of gravmag/_tesseroid_numba.py evaluated in
 * IEEE binary128 (__float128, libquadmath), used to decide which of two FP64 implementations (the
 * reference's numba engine / the CUDA kernel) is closer to the exact value of the SAME quadrature
 * on the deeply subdivided near-field pairs, where l^2 = r^2 + rc^2 - 2 r rc cos(psi) cancels ~1e9-fold
 * and a 1-ulp difference between two libm's cos() moves the FP64 result by ~1e-7.
 *
 * Same algorithm as oracle_tess.c (and therefore as the reference):
 *   - the subdivision DECISIONS (_tesseroid_numba.py:94-157 distance_size / divisions) and the leaf
 *     bounds (:114-132 split) are computed in FP64 exactly like the reference, so the set of leaves
 *     is identical;
 *   - every leaf's 2x2x2 Gauss-Legendre sum (:75-91 scale_nodes, :207-222 kernelz) and the
 *     accumulation over leaves are carried out in binary128 from the same FP64 inputs
 *     (lon, sinlat, coslat, radius of the observation; w, e, s, n, top, bottom of the leaf), and
 *     rounded to FP64 once at the end.
 * Never linked into, imported by or called from the product.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared oracle_tess_quad.c -lquadmath -lm
 */
#include <math.h>
#include <quadmath.h>
#include <stddef.h>
#include <stdint.h>

#define MEAN_EARTH_RADIUS 6378137.0
#define STACK_SIZE 100
typedef __float128 q_t;

static const double PI_NP = 3.141592653589793; /* np.pi, the FP64 constant the reference multiplies by */
/* the reference's nodes are the FP64 values +-0.5773502691896257...: keep the same FP64 inputs */
static const double NODES[2] = {-0.577350269189625731058868041146, 0.577350269189625731058868041146};

static void distance_size(double lon, double coslat, double sinlat, double radius, double w, double e,
                          double s, double n, double top, double bottom, double *distance,
                          double *Llon, double *Llat, double *Lr)
{
    const double d2r = PI_NP / 180;
    double rt = 0.5 * (top + bottom) + MEAN_EARTH_RADIUS;
    double lont = d2r * 0.5 * (w + e);
    double latt = d2r * 0.5 * (s + n);
    double sinlatt = sin(latt);
    double coslatt = cos(latt);
    double cospsi = sinlat * sinlatt + coslat * coslatt * cos(lon - lont);
    *distance = sqrt(radius * radius + rt * rt - 2 * radius * rt * cospsi);
    double rtop = top + MEAN_EARTH_RADIUS;
    *Llon = rtop * acos(sinlatt * sinlatt + (coslatt * coslatt) * cos(d2r * (e - w)));
    *Llat = rtop * acos(sin(d2r * n) * sin(d2r * s) + cos(d2r * n) * cos(d2r * s));
    *Lr = top - bottom;
}

static q_t leaf_quad(double lon, double coslat, double sinlat, double radius, double w, double e,
                     double s, double n, double top, double bottom)
{
    const q_t d2r = (q_t)PI_NP / 180;
    q_t dlon = d2r * ((q_t)e - w), dlat = d2r * ((q_t)n - s), dr = (q_t)top - bottom;
    q_t lonc[2], sinlatc[2], coslatc[2], rc[2];
    for (int i = 0; i < 2; ++i) {
        lonc[i] = 0.5Q * dlon * NODES[i] + d2r * 0.5Q * ((q_t)e + w);
        q_t latc = 0.5Q * dlat * NODES[i] + d2r * 0.5Q * ((q_t)n + s);
        sinlatc[i] = sinq(latc);
        coslatc[i] = cosq(latc);
        rc[i] = 0.5Q * dr * NODES[i] + 0.5Q * ((q_t)top + bottom) + MEAN_EARTH_RADIUS;
    }
    q_t scale = dlon * dlat * dr * 0.125Q;
    q_t r = radius, r_sqr = r * r, result = 0;
    for (int i = 0; i < 2; ++i) {
        q_t coslon = cosq((q_t)lon - lonc[i]);
        for (int j = 0; j < 2; ++j) {
            q_t cospsi = (q_t)sinlat * sinlatc[j] + (q_t)coslat * coslatc[j] * coslon;
            for (int k = 0; k < 2; ++k) {
                q_t l_sqr = r_sqr + rc[k] * rc[k] - 2 * r * rc[k] * cospsi;
                q_t kappa = rc[k] * rc[k] * coslatc[j];
                result += kappa * (rc[k] * cospsi - r) / (l_sqr * sqrtq(l_sqr));
            }
        }
    }
    return -scale * result;
}

/* raw (unscaled) kernel of one (observation, tesseroid) pair: leaves as the FP64 engine picks them,
 * leaf sums and accumulation in binary128; returns NaN on stack overflow; *leaves = leaf count */
double oracle_tess_gz_pair_quad(double lon, double sinlat, double coslat, double radius,
                                const double *bounds, double ratio, int32_t *leaves)
{
    double stack[STACK_SIZE][6];
    q_t acc = 0;
    int32_t nleaf = 0;
    for (int i = 0; i < 6; ++i) stack[0][i] = bounds[i];
    int stktop = 0;
    while (stktop >= 0) {
        double w = stack[stktop][0], e = stack[stktop][1], s = stack[stktop][2], n = stack[stktop][3],
               top = stack[stktop][4], bottom = stack[stktop][5];
        stktop -= 1;
        double distance, Llon, Llat, Lr;
        distance_size(lon, coslat, sinlat, radius, w, e, s, n, top, bottom, &distance, &Llon, &Llat, &Lr);
        int nlon = 1, nlat = 1, nr = 1;
        if (distance <= ratio * Llon && !(Llon <= 0.1)) nlon = 2;
        if (distance <= ratio * Llat && !(Llat <= 0.1)) nlat = 2;
        if (distance <= ratio * Lr && !(Lr <= 1e3)) nr = 2;
        int new_cells = nlon * nlat * nr;
        if (new_cells > 1) {
            if (new_cells + (stktop + 1) > STACK_SIZE) return NAN;
            double dlon = (e - w) / nlon, dlat = (n - s) / nlat, dr = (top - bottom) / nr;
            for (int i = 0; i < nlon; ++i)
                for (int j = 0; j < nlat; ++j)
                    for (int k = 0; k < nr; ++k) {
                        stktop += 1;
                        stack[stktop][0] = w + i * dlon;
                        stack[stktop][1] = w + (i + 1) * dlon;
                        stack[stktop][2] = s + j * dlat;
                        stack[stktop][3] = s + (j + 1) * dlat;
                        stack[stktop][4] = bottom + (k + 1) * dr;
                        stack[stktop][5] = bottom + k * dr;
                    }
        } else {
            acc += leaf_quad(lon, coslat, sinlat, radius, w, e, s, n, top, bottom);
            nleaf += 1;
        }
    }
    if (leaves) *leaves = nleaf;
    return (double)acc;
}

/* out[i] = scale1*scale2 * kernel of pair (obs[i], cell[i]) for npairs listed pairs */
void oracle_tess_gz_pairs_quad(const double *lon, const double *sinlat, const double *coslat,
                               const double *radius, const int64_t *obs, const double *bounds,
                               const int64_t *cell, int64_t npairs, double ratio, double scale1,
                               double scale2, double *out, int32_t *leaves)
{
    for (int64_t i = 0; i < npairs; ++i) {
        const int64_t l = obs[i];
        int32_t nl = 0;
        double v = oracle_tess_gz_pair_quad(lon[l], sinlat[l], coslat[l], radius[l], bounds + 6 * cell[i],
                                            ratio, &nl);
        out[i] = v * scale1 * scale2;
        if (leaves) leaves[i] = nl;
    }
}
