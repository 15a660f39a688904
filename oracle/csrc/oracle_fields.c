/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's other prism fields
 * (SURVEY.md 8(f3)).  Never linked into, imported by or called from the product.
 *
 * Follows, in plain C:
 *   gravmag/_prism.pyx:36-70    kernelpot, kernelx, kernely, kernelz, kernelxx .. kernelzz
 *   gravmag/_prism.pyx:484-509  potential   :206-290 gx, gy, gz   :294-480 gxx .. gzz (with the
 *                               displaced radius of gxy :345-350, gxz :380-385, gyz :442-447)
 *   gravmag/_prism.pyx:72-112   tf (kernel1D = f.(V f), res = f.(V m))   :116-202 bx, by, bz
 *   gravmag/prism.py:102-732    one kernel2d column per prism, scale applied after the 8-corner sum
 * Pinned against the compiled reference (oracle/_ref/_prism*.so) and tests/golden/fields.npz.
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction, like the reference's x86-64 build).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

static const double PI_LIT = 3.1415926535897931159979634685441851615906; /* _prism.pyx:21 */

static inline double safe_atan2(double y, double x)
{
    if (y == 0) return 0;
    if ((y > 0) && (x < 0)) return atan2(y, x) - PI_LIT;
    if ((y < 0) && (x < 0)) return atan2(y, x) + PI_LIT;
    return atan2(y, x);
}

static inline double safe_log(double x) { return (x == 0) ? 0.0 : log(x); }

static inline double kernelpot(double x, double y, double z, double r)
{
    return (x * y * safe_log(z + r) + y * z * safe_log(x + r) + x * z * safe_log(y + r)
            - 0.5 * (x * x) * safe_atan2(z * y, x * r) - 0.5 * (y * y) * safe_atan2(z * x, y * r)
            - 0.5 * (z * z) * safe_atan2(x * y, z * r));
}
static inline double kernelx(double x, double y, double z, double r)
{ return -(y * safe_log(z + r) + z * safe_log(y + r) - x * safe_atan2(z * y, x * r)); }
static inline double kernely(double x, double y, double z, double r)
{ return -(z * safe_log(x + r) + x * safe_log(z + r) - y * safe_atan2(x * z, y * r)); }
static inline double kernelz(double x, double y, double z, double r)
{ return -(x * safe_log(y + r) + y * safe_log(x + r) - z * safe_atan2(x * y, z * r)); }
static inline double kernelxx(double x, double y, double z, double r) { return -safe_atan2(z * y, x * r); }
static inline double kernelxy(double x, double y, double z, double r) { (void)x; (void)y; return safe_log(z + r); }
static inline double kernelxz(double x, double y, double z, double r) { (void)x; (void)z; return safe_log(y + r); }
static inline double kernelyy(double x, double y, double z, double r) { return -safe_atan2(z * x, y * r); }
static inline double kernelyz(double x, double y, double z, double r) { (void)y; (void)z; return safe_log(x + r); }
static inline double kernelzz(double x, double y, double z, double r) { return -safe_atan2(x * y, z * r); }

enum { F_POT = 0, F_GX, F_GY, F_GZ, F_GXX, F_GXY, F_GXZ, F_GYY, F_GYZ, F_GZZ, F_TF, F_VX, F_VY, F_VZ };

/* kernel2d[N][ld] for `field`; vec = (fx,fy,fz) for tf, the vector for vx/vy/vz.
 * wts: optional per-prism weights [M][nw] (nw = 1: density; nw = 3: magnetisation vectors for the
 * forward result of tf / bx / by / bz); res[l] accumulates in the reference's order. */
int oracle_prism_field(int field, const double *xp, const double *yp, const double *zp, int64_t N,
                       const double *bounds, int64_t M, double scale, const double *vec,
                       double *kernel2d, int64_t ld, const double *wts, int nw, double *res)
{
    if (field < F_POT || field > F_VZ) return -1;
    for (int64_t c = 0; c < M; ++c) {
        const double *b = bounds + 6 * c;
        const double x[2] = {b[1], b[0]}, y[2] = {b[3], b[2]}, z[2] = {b[5], b[4]};
        const double x1 = b[0], x2 = b[1], y1 = b[2], y2 = b[3], z1 = b[4], z2 = b[5];
        for (int64_t l = 0; l < N; ++l) {
            double acc = 0.0;
            for (int k = 0; k < 2; ++k) {
                double dz = z[k] - zp[l];
                for (int j = 0; j < 2; ++j) {
                    double dy = y[j] - yp[l];
                    for (int i = 0; i < 2; ++i) {
                        double dx = x[i] - xp[l];
                        double r, kern = 0.0, kres = 0.0;
                        double sign = ((i + j + k) & 1) ? -1.0 : 1.0;
                        if (field == F_GXY && dx == 0 && dy == 0 && dz < 0) {
                            double t1 = 0.00001 * (x2 - x1), t2 = 0.00001 * (y2 - y1);
                            r = sqrt(t1 * t1 + t2 * t2 + dz * dz);
                        } else if (field == F_GXZ && dx == 0 && dz == 0 && dy < 0) {
                            double t1 = 0.00001 * (x2 - x1), t2 = 0.00001 * (z2 - z1);
                            r = sqrt(t1 * t1 + t2 * t2 + dy * dy);
                        } else if (field == F_GYZ && dy == 0 && dz == 0 && dx < 0) {
                            double t1 = 0.00001 * (y2 - y1), t2 = 0.00001 * (z2 - z1);
                            r = sqrt(t1 * t1 + t2 * t2 + dx * dx);
                        } else {
                            r = sqrt(dx * dx + dy * dy + dz * dz);
                        }
                        switch (field) {
                        case F_POT: kern = kernelpot(dx, dy, dz, r); break;
                        case F_GX: kern = kernelx(dx, dy, dz, r); break;
                        case F_GY: kern = kernely(dx, dy, dz, r); break;
                        case F_GZ: kern = kernelz(dx, dy, dz, r); break;
                        case F_GXX: kern = kernelxx(dx, dy, dz, r); break;
                        case F_GXY: kern = kernelxy(dx, dy, dz, r); break;
                        case F_GXZ: kern = kernelxz(dx, dy, dz, r); break;
                        case F_GYY: kern = kernelyy(dx, dy, dz, r); break;
                        case F_GYZ: kern = kernelyz(dx, dy, dz, r); break;
                        case F_GZZ: kern = kernelzz(dx, dy, dz, r); break;
                        default: {
                            double v1 = kernelxx(dx, dy, dz, r), v2 = kernelxy(dx, dy, dz, r);
                            double v3 = kernelxz(dx, dy, dz, r), v4 = kernelyy(dx, dy, dz, r);
                            double v5 = kernelyz(dx, dy, dz, r), v6 = kernelzz(dx, dy, dz, r);
                            double fx = vec[0], fy = vec[1], fz = vec[2];
                            if (field == F_TF) {
                                double bxk = (v1 * fx + v2 * fy + v3 * fz);
                                double byk = (v2 * fx + v4 * fy + v5 * fz);
                                double bzk = (v3 * fx + v5 * fy + v6 * fz);
                                kern = fx * bxk + fy * byk + fz * bzk;
                                if (wts && nw == 3) {
                                    double mx = wts[3 * c], my = wts[3 * c + 1], mz = wts[3 * c + 2];
                                    double bx = (v1 * mx + v2 * my + v3 * mz);
                                    double by = (v2 * mx + v4 * my + v5 * mz);
                                    double bz = (v3 * mx + v5 * my + v6 * mz);
                                    kres = fx * bx + fy * by + fz * bz;
                                }
                            } else {
                                double ux = vec[0], uy = vec[1], uz = vec[2];
                                if (wts && nw == 3) { ux = wts[3 * c]; uy = wts[3 * c + 1]; uz = wts[3 * c + 2]; }
                                if (field == F_VX) kern = (v1 * ux + v2 * uy + v3 * uz);
                                else if (field == F_VY) kern = (v2 * ux + v4 * uy + v5 * uz);
                                else kern = (v3 * ux + v5 * uy + v6 * uz);
                                kres = kern;
                            }
                        }
                        }
                        acc += sign * kern;
                        if (res && wts) {
                            if (nw == 1) res[l] += sign * kern * wts[c];
                            else res[l] += sign * kres;
                        }
                    }
                }
            }
            if (kernel2d) kernel2d[l * ld + c] = acc * scale;
        }
    }
    if (res && wts)
        for (int64_t l = 0; l < N; ++l) res[l] *= scale;
    return 0;
}
