/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's tesseroid gz
 * sensitivity assembly (adaptive 2x2x2 Gauss-Legendre with LIFO subdivision stack).
 * Never linked into, imported by or called from the product.
 *
 * Follows, in plain C:
 *   gravmag/_tesseroid_numba.py:21-22    nodes
 *   gravmag/_tesseroid_numba.py:32-71    engine (stack pop order, overflow check, accumulate)
 *   gravmag/_tesseroid_numba.py:75-91    scale_nodes
 *   gravmag/_tesseroid_numba.py:94-111   distance_size
 *   gravmag/_tesseroid_numba.py:114-132  split (push order lon, lat, r)
 *   gravmag/_tesseroid_numba.py:135-157  divisions
 *   gravmag/_tesseroid_numba.py:207-222  kernelz
 *   gravmag/tesseroid.py:109-123         _convert_coords   (done by the caller in numpy)
 *   gravmag/tesseroid.py:429-430         scale SI2MGAL*G applied after accumulation
 * Pinned against the reference run in this container (tests/golden/, oracle/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -lm  (numba/LLVM does not contract FMAs).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define MEAN_EARTH_RADIUS 6378137.0 /* constants.py:44 */
#define STACK_SIZE 100              /* gravmag/tesseroid.py:79 */

static const double NODES[2] = {-0.577350269189625731058868041146,
                                0.577350269189625731058868041146};
static const double PI_NP = 3.141592653589793; /* np.pi */

static double scale_nodes(double w, double e, double s, double n, double top, double bottom,
                          double *lonc, double *sinlatc, double *coslatc, double *rc)
{
    const double d2r = PI_NP / 180;
    double dlon = d2r * (e - w);
    double dlat = d2r * (n - s);
    double dr = top - bottom;
    for (int i = 0; i < 2; ++i) {
        lonc[i] = 0.5 * dlon * NODES[i] + d2r * 0.5 * (e + w);
        double latc = 0.5 * dlat * NODES[i] + d2r * 0.5 * (n + s);
        sinlatc[i] = sin(latc);
        coslatc[i] = cos(latc);
        rc[i] = (0.5 * dr * NODES[i] + 0.5 * (top + bottom) + MEAN_EARTH_RADIUS);
    }
    return dlon * dlat * dr * 0.125;
}

static void distance_size(double lon, double coslat, double sinlat, double radius,
                          double w, double e, double s, double n, double top, double bottom,
                          double *distance, double *Llon, double *Llat, double *Lr)
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

static double kernelz(double lon, double coslat, double sinlat, double radius,
                      const double *lonc, const double *sinlatc, const double *coslatc,
                      const double *rc)
{
    double r_sqr = radius * radius;
    double result = 0;
    for (int i = 0; i < 2; ++i) {
        double coslon = cos(lon - lonc[i]);
        for (int j = 0; j < 2; ++j) {
            double cospsi = sinlat * sinlatc[j] + coslat * coslatc[j] * coslon;
            for (int k = 0; k < 2; ++k) {
                double l_sqr = r_sqr + rc[k] * rc[k] - 2 * radius * rc[k] * cospsi;
                double kappa = (rc[k] * rc[k]) * coslatc[j];
                result += kappa * (rc[k] * cospsi - radius) / pow(l_sqr, 1.5);
            }
        }
    }
    result *= -1;
    return result;
}

/* The other GLQ kernels, gravmag/_tesseroid_numba.py:160-328 (kernelV, kernelx, kernely, kernelxx,
 * kernelxy, kernelxz, kernelyy, kernelyz, kernelzz), in numba's evaluation order. */
enum { TF_POT = 0, TF_GX, TF_GY, TF_GZ, TF_GXX, TF_GXY, TF_GXZ, TF_GYY, TF_GYZ, TF_GZZ };

static double kernel_field(int field, double lon, double coslat, double sinlat, double radius,
                           const double *lonc, const double *sinlatc, const double *coslatc,
                           const double *rc)
{
    if (field == TF_GZ)
        return kernelz(lon, coslat, sinlat, radius, lonc, sinlatc, coslatc, rc);
    double r_sqr = radius * radius;
    double result = 0;
    for (int i = 0; i < 2; ++i) {
        double coslon = cos(lon - lonc[i]);
        double sinlon = sin(lonc[i] - lon);
        for (int j = 0; j < 2; ++j) {
            double kphi = coslat * sinlatc[j] - sinlat * coslatc[j] * coslon;
            double cospsi = sinlat * sinlatc[j] + coslat * coslatc[j] * coslon;
            for (int k = 0; k < 2; ++k) {
                double rc_sqr = rc[k] * rc[k];
                double l_sqr = r_sqr + rc_sqr - 2 * radius * rc[k] * cospsi;
                double kappa = rc_sqr * coslatc[j];
                double deltay = rc[k] * coslatc[j] * sinlon;
                double deltaz = rc[k] * cospsi - radius;
                switch (field) {
                case TF_POT: result += kappa / sqrt(l_sqr); break;
                case TF_GX: result += kappa * rc[k] * kphi / pow(l_sqr, 1.5); break;
                case TF_GY: result += kappa * (rc[k] * coslatc[j] * sinlon / pow(l_sqr, 1.5)); break;
                case TF_GXX: {
                    double t = rc[k] * kphi;
                    result += kappa * (3 * (t * t) - l_sqr) / pow(l_sqr, 2.5);
                    break;
                }
                case TF_GXY: result += kappa * 3 * rc_sqr * kphi * coslatc[j] * sinlon / pow(l_sqr, 2.5); break;
                case TF_GXZ: result += kappa * 3 * rc[k] * kphi * deltaz / pow(l_sqr, 2.5); break;
                case TF_GYY: result += kappa * (3 * (deltay * deltay) - l_sqr) / pow(l_sqr, 2.5); break;
                case TF_GYZ: result += kappa * 3. * deltay * deltaz / pow(l_sqr, 2.5); break;
                default: result += kappa * (3 * (deltaz * deltaz) - l_sqr) / pow(l_sqr, 2.5); break;
                }
            }
        }
    }
    return result;
}

/* One (observation, tesseroid) pair, raw (unscaled) kernel value.
 * *err accumulates the reference's error_code (-1 per refused split);
 * returns NaN and sets *overflow=1 where the reference raises OverflowError.
 * stats (optional): [0] += leaves evaluated, [1] = max(stack depth).  */
double oracle_tess_field_pair(int field, double lon, double sinlat, double coslat, double radius,
                              const double *bounds, double ratio, int *err, int *overflow,
                              int64_t *stats)
{
    double stack[STACK_SIZE][6];
    double lonc[2], sinlatc[2], coslatc[2], rc[2];
    double acc = 0.0;
    for (int i = 0; i < 6; ++i)
        stack[0][i] = bounds[i];
    int stktop = 0;
    while (stktop >= 0) {
        double w = stack[stktop][0], e = stack[stktop][1], s = stack[stktop][2],
               n = stack[stktop][3], top = stack[stktop][4], bottom = stack[stktop][5];
        stktop -= 1;
        double distance, Llon, Llat, Lr;
        distance_size(lon, coslat, sinlat, radius, w, e, s, n, top, bottom, &distance, &Llon,
                      &Llat, &Lr);
        int nlon = 1, nlat = 1, nr = 1, error = 0;
        if (distance <= ratio * Llon) {
            if (Llon <= 0.1)
                error = -1;
            else
                nlon = 2;
        }
        if (distance <= ratio * Llat) {
            if (Llat <= 0.1)
                error = -1;
            else
                nlat = 2;
        }
        if (distance <= ratio * Lr) {
            if (Lr <= 1e3)
                error = -1;
            else
                nr = 2;
        }
        if (err)
            *err += error;
        int new_cells = nlon * nlat * nr;
        if (new_cells > 1) {
            if (new_cells + (stktop + 1) > STACK_SIZE) {
                if (overflow)
                    *overflow = 1;
                return NAN;
            }
            double dlon = (e - w) / nlon;
            double dlat = (n - s) / nlat;
            double dr = (top - bottom) / nr;
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
            if (stats && stktop + 1 > stats[1])
                stats[1] = stktop + 1;
        } else {
            double scale = scale_nodes(w, e, s, n, top, bottom, lonc, sinlatc, coslatc, rc);
            acc += scale * kernel_field(field, lon, coslat, sinlat, radius, lonc, sinlatc, coslatc, rc);
            if (stats)
                stats[0] += 1;
        }
    }
    return acc;
}

double oracle_tess_gz_pair(double lon, double sinlat, double coslat, double radius,
                           const double *bounds, double ratio, int *err, int *overflow,
                           int64_t *stats)
{
    return oracle_tess_field_pair(TF_GZ, lon, sinlat, coslat, radius, bounds, ratio, err, overflow, stats);
}

/* kernel2d for any field (gravmag/tesseroid.py:324-510); dens (optional, [M]) also accumulates the
 * forward result res[l] += density * scale * kernel in the reference's order
 * (_tesseroid_numba.py:62-64), scaled like kernel2d. */
int oracle_tess_field(int field, const double *lon, const double *sinlat, const double *coslat,
                      const double *radius, int64_t N, const double *bounds, int64_t M, double ratio,
                      double scale1, double scale2, double *kernel2d, int64_t ld, const double *dens,
                      double *res, int *overflow)
{
    int err = 0;
    if (field < TF_POT || field > TF_GZZ) return -9999;
    for (int64_t c = 0; c < M; ++c)
        for (int64_t l = 0; l < N; ++l) {
            double v = oracle_tess_field_pair(field, lon[l], sinlat[l], coslat[l], radius[l],
                                              bounds + 6 * c, ratio, &err, overflow, NULL);
            if (kernel2d) kernel2d[l * ld + c] = v * scale1 * scale2;
            if (dens && res) res[l] += dens[c] * v;
        }
    if (dens && res)
        for (int64_t l = 0; l < N; ++l) res[l] = res[l] * scale1 * scale2;
    return err;
}

/* kernel2d[N][ld] row-major for M active tesseroids, bounds[M][6] = w,e,s,n,top,bottom.
 * Inputs are the converted coordinates of tesseroid.py:109-123 (lon in rad, sinlat, coslat,
 * radius).  Returns the summed error_code; *overflow set if any pair overflowed the stack.
 * The reference scales in two steps, kernel2d*SI2MGAL*G (tesseroid.py:430) -> scale1, scale2. */
int oracle_tess_gz(const double *lon, const double *sinlat, const double *coslat,
                   const double *radius, int64_t N, const double *bounds, int64_t M,
                   double ratio, double scale1, double scale2, double *kernel2d, int64_t ld,
                   int *overflow, int64_t *stats)
{
    int err = 0;
    for (int64_t c = 0; c < M; ++c)
        for (int64_t l = 0; l < N; ++l) {
            double v = oracle_tess_gz_pair(lon[l], sinlat[l], coslat[l], radius[l],
                                           bounds + 6 * c, ratio, &err, overflow, stats);
            kernel2d[l * ld + c] = v * scale1 * scale2; /* (k*SI2MGAL)*G, tesseroid.py:430 */
        }
    return err;
}

/* Subdivision bookkeeping only: leaves[l*ld + c] = number of leaf cells the adaptive subdivision of
 * _tesseroid_numba.py:32-71 evaluates for pair (l, c) (-1 where the stack would overflow). */
int oracle_tess_leaves(const double *lon, const double *sinlat, const double *coslat,
                       const double *radius, int64_t N, const double *bounds, int64_t M,
                       double ratio, int32_t *leaves, int64_t ld)
{
    for (int64_t c = 0; c < M; ++c)
        for (int64_t l = 0; l < N; ++l) {
            int64_t st[2] = {0, 0};
            int err = 0, ovf = 0;
            oracle_tess_gz_pair(lon[l], sinlat[l], coslat[l], radius[l], bounds + 6 * c, ratio, &err,
                                &ovf, st);
            leaves[l * ld + c] = ovf ? -1 : (int32_t)st[0];
        }
    return 0;
}
