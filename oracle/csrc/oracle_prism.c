/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's prism gz
 * sensitivity assembly.  Never linked into, imported by or called from the product
 * (gravinv3dhmc_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * Follows, in plain C:
 *   gravmag/_prism.pyx:16-26   safe_atan2
 *   gravmag/_prism.pyx:28-34   safe_log
 *   gravmag/_prism.pyx:49-50   kernelz
 *   gravmag/_prism.pyx:263-290 gz  (loop order k{z2,z1} j{y2,y1} i{x2,x1}, sign (-1)^(i+j+k))
 *   gravmag/prism.py:291-316   _gz (one column per non-masked prism, scale G*SI2MGAL applied
 *                                   AFTER the 8-term sum)
 * Pinned against the compiled reference (oracle/_ref/_prism*.so) and the golden vectors in
 * tests/golden/ by tests/test_oracle_pinning.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no FMA contraction: the reference is
 * generic x86-64 code without FMA).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

static const double PI_LIT = 3.1415926535897931159979634685441851615906; /* _prism.pyx:21 */

static inline double safe_atan2(double y, double x)
{
    double res;
    if (y == 0)
        res = 0;
    else if ((y > 0) && (x < 0))
        res = atan2(y, x) - PI_LIT;
    else if ((y < 0) && (x < 0))
        res = atan2(y, x) + PI_LIT;
    else
        res = atan2(y, x);
    return res;
}

static inline double safe_log(double x)
{
    return (x == 0) ? 0.0 : log(x);
}

static inline double kernelz(double x, double y, double z, double r)
{
    return -(x * safe_log(y + r) + y * safe_log(x + r) - z * safe_atan2(x * y, z * r));
}

/* one (observation, prism) pair: the raw 8-corner sum, before the G*SI2MGAL scale */
double oracle_prism_gz_pair(double xp, double yp, double zp,
                            double x1, double x2, double y1, double y2, double z1, double z2)
{
    const double x[2] = {x2, x1}, y[2] = {y2, y1}, z[2] = {z2, z1};
    double acc = 0.0;
    for (int k = 0; k < 2; ++k) {
        double dz = z[k] - zp;
        for (int j = 0; j < 2; ++j) {
            double dy = y[j] - yp;
            for (int i = 0; i < 2; ++i) {
                double dx = x[i] - xp;
                double r = sqrt(dx * dx + dy * dy + dz * dz);
                double kern = kernelz(dx, dy, dz, r);
                double sign = ((i + j + k) & 1) ? -1.0 : 1.0;
                acc += sign * kern;
            }
        }
    }
    return acc;
}

/* kernel2d[N][ld] (row-major) for M active prisms with bounds[M][6] = x1,x2,y1,y2,z1,z2.
 * If dens != NULL also accumulates res[l] += sign*kernel*dens[k] in the reference order
 * (prism-major, then corner order) and scales it.  */
void oracle_prism_gz(const double *xp, const double *yp, const double *zp, int64_t N,
                     const double *bounds, int64_t M, double scale,
                     double *kernel2d, int64_t ld, const double *dens, double *res)
{
    for (int64_t c = 0; c < M; ++c) {
        const double *b = bounds + 6 * c;
        const double x[2] = {b[1], b[0]}, y[2] = {b[3], b[2]}, z[2] = {b[5], b[4]};
        for (int64_t l = 0; l < N; ++l) {
            double acc = 0.0;
            for (int k = 0; k < 2; ++k) {
                double dz = z[k] - zp[l];
                for (int j = 0; j < 2; ++j) {
                    double dy = y[j] - yp[l];
                    for (int i = 0; i < 2; ++i) {
                        double dx = x[i] - xp[l];
                        double r = sqrt(dx * dx + dy * dy + dz * dz);
                        double kern = kernelz(dx, dy, dz, r);
                        double sign = ((i + j + k) & 1) ? -1.0 : 1.0;
                        acc += sign * kern;
                        if (dens && res)
                            res[l] += sign * kern * dens[c];
                    }
                }
            }
            kernel2d[l * ld + c] = acc * scale;
        }
    }
    if (dens && res)
        for (int64_t l = 0; l < N; ++l)
            res[l] *= scale;
}
