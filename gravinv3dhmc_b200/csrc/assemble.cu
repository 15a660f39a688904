// assemble.cu -- sensitivity-matrix assembly and column weighting on sm_100a.
//
// Replaces (reference paths relative to the reference root):
//   gravmag/_prism.pyx:263-290 gz + :49-50 kernelz + :16-34 safe_atan2/safe_log, driven per prism
//     by gravmag/prism.py:291-316 _gz                      -> prism_gz_kernel
//   gravmag/_tesseroid_numba.py:25-157,207-222 engine/scale_nodes/distance_size/split/divisions/
//     kernelz, driven per tesseroid by gravmag/tesseroid.py:189-232   -> tess_gz_kernel
//   inversion/potential.py:232-264 sensitivityWeighting     -> colsumsq / weights / scale_columns
//
// Layout: G is row-major [nrows][ld]; consecutive threads own consecutive COLUMNS (cells) of one
// observation row, so every warp store is one contiguous 256 B segment.  The arithmetic keeps the
// reference's operation order and uses explicit round-to-nearest intrinsics (__dmul_rn/__dadd_rn)
// so ptxas cannot contract multiply-adds: the closed-form prism kernel cancels ~1e7..1e10 and the
// reference is plain x86-64 code without FMA.
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "prism_math.cuh"
#include "tess_math.cuh"

namespace gi {

// ---------------------------------------------------------------------------------------------
// prism gz
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double prism_corner(double x, double y, double z) {
    // r = sqrt(x**2 + y**2 + z**2)  (_prism.pyx:286);  kernelz (_prism.pyx:49-50):
    // -(x*log(y + r) + y*log(x + r) - z*atan2(x*y, z*r))
    const double r =
        __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    const double t1 = __dmul_rn(x, prism_safe_log(__dadd_rn(y, r)));
    const double t2 = __dmul_rn(y, prism_safe_log(__dadd_rn(x, r)));
    const double t3 = __dmul_rn(z, prism_safe_atan2(__dmul_rn(x, y), __dmul_rn(z, r)));
    return -__dsub_rn(__dadd_rn(t1, t2), t3);
}

constexpr int kAsmThreads = 128;  // columns per CTA
constexpr int kAsmRows = 8;       // observation rows per CTA

__global__ void __launch_bounds__(kAsmThreads)
prism_gz_kernel(const double *__restrict__ xp, const double *__restrict__ yp,
                const double *__restrict__ zp, int64_t nrows, const double *__restrict__ bounds,
                int64_t M, double scale, double *__restrict__ G, int64_t ld) {
    const int64_t col = (int64_t)blockIdx.x * kAsmThreads + threadIdx.x;
    if (col >= ld) return;
    const bool live = col < M;
    double bx[2], by[2], bz[2];
    if (live) {
        const double *b = bounds + 6 * col;
        // corner order of _prism.pyx:272-274: x = [x2, x1], y = [y2, y1], z = [z2, z1]
        bx[0] = b[1]; bx[1] = b[0];
        by[0] = b[3]; by[1] = b[2];
        bz[0] = b[5]; bz[1] = b[4];
    }
    for (int64_t tile = blockIdx.y; tile * kAsmRows < nrows; tile += gridDim.y) {
        const int64_t r0 = tile * kAsmRows;
        const int nr = (int)min((int64_t)kAsmRows, nrows - r0);
        for (int rr = 0; rr < nr; ++rr) {
            const int64_t row = r0 + rr;
            double acc = 0.0;
            if (live) {
                const double ox = __ldg(xp + row), oy = __ldg(yp + row), oz = __ldg(zp + row);
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    // loop nest k (z) outer, j (y), i (x) inner; sign (-1)^(i+j+k)
                    const int k = c >> 2, j = (c >> 1) & 1, i = c & 1;
                    const double dz = __dsub_rn(bz[k], oz);
                    const double dy = __dsub_rn(by[j], oy);
                    const double dx = __dsub_rn(bx[i], ox);
                    const double kern = prism_corner(dx, dy, dz);
                    acc = __dadd_rn(acc, ((i + j + k) & 1) ? -kern : kern);
                }
                acc = __dmul_rn(acc, scale);  // prism.py:314-315, after the 8-term sum
            }
            G[row * ld + col] = acc;
        }
    }
}

// Structured-grid variant: neighbouring prisms of a regular / stretched / segmented mesh share their
// corners, so the corner term of _prism.pyx:286-289 is evaluated once per grid NODE and each cell
// combines its 8 node values in the reference's order and signs.  Bit-identical to the per-cell
// kernel whenever the upper edge of cell i is the same double as the lower edge of cell i+1 on every
// axis (the host checks this; otherwise the general kernel runs): the corner term only sees
// (edge - observation), and the 8-term sum is formed in the same order.
// A CTA owns a (4 x 4 x 32)-cell tile = 825 nodes for 512 cells -> 1.6 corner evaluations per cell
// instead of 8; node values live in shared memory; cell rows are 256-B contiguous stores.
constexpr int kGridTX = 32, kGridTY = 4, kGridTZ = 4;
constexpr int kGridThreads = 128;
constexpr int kGridNodes = (kGridTX + 1) * (kGridTY + 1) * (kGridTZ + 1);

__global__ void __launch_bounds__(kGridThreads)
prism_gz_grid_kernel(const double *__restrict__ xp, const double *__restrict__ yp,
                     const double *__restrict__ zp, int64_t nrows, const double *__restrict__ xn,
                     const double *__restrict__ yn, const double *__restrict__ zn, int nx, int ny, int nz,
                     int tiles_x, int tiles_y, const int32_t *__restrict__ colmap, double scale,
                     double *__restrict__ G, int64_t ld) {
    __shared__ double F[kGridNodes];
    __shared__ double ex[kGridTX + 1], ey[kGridTY + 1], ez[kGridTZ + 1];
    int tile = blockIdx.x;
    const int tx = tile % tiles_x;
    tile /= tiles_x;
    const int ty = tile % tiles_y, tz = tile / tiles_y;
    const int i0 = tx * kGridTX, j0 = ty * kGridTY, k0 = tz * kGridTZ;
    const int cx = min(kGridTX, nx - i0), cy = min(kGridTY, ny - j0), cz = min(kGridTZ, nz - k0);
    for (int t = threadIdx.x; t <= cx; t += kGridThreads) ex[t] = xn[i0 + t];
    if (threadIdx.x <= cy) ey[threadIdx.x] = yn[j0 + threadIdx.x];
    if (threadIdx.x <= cz) ez[threadIdx.x] = zn[k0 + threadIdx.x];
    __syncthreads();
    // compile-time tile indexing (no integer divisions by runtime values); edge tiles are masked
    constexpr int PX = kGridTX + 1, PXY = PX * (kGridTY + 1);
    constexpr int SX = 1, SY = PX, SZ = PXY;
    constexpr int NPASS = (kGridNodes + kGridThreads - 1) / kGridThreads;
    constexpr int CPASS = kGridTX * kGridTY * kGridTZ / kGridThreads;
    // this thread's nodes and cells are the same for every observation row
    double nxe[NPASS], nye[NPASS], nze[NPASS];
    bool nlive[NPASS];
#pragma unroll
    for (int u = 0; u < NPASS; ++u) {
        const int n = threadIdx.x + u * kGridThreads;
        const int c = n / PXY, rem = n - c * PXY, b = rem / PX, a = rem - b * PX;
        nlive[u] = n < kGridNodes && a <= cx && b <= cy && c <= cz;
        nxe[u] = nlive[u] ? ex[a] : 0.0;
        nye[u] = nlive[u] ? ey[b] : 0.0;
        nze[u] = nlive[u] ? ez[c] : 0.0;
    }
    int64_t ccol[CPASS];
    int cofs[CPASS];
#pragma unroll
    for (int u = 0; u < CPASS; ++u) {
        const int q = threadIdx.x + u * kGridThreads;
        const int a = q % kGridTX, b = (q / kGridTX) % kGridTY, c = q / (kGridTX * kGridTY);
        cofs[u] = c * SZ + b * SY + a;
        ccol[u] = -1;
        if (a < cx && b < cy && c < cz) {
            const int64_t cell = ((int64_t)(k0 + c) * ny + (j0 + b)) * nx + (i0 + a);
            ccol[u] = colmap ? (int64_t)colmap[cell] : cell;
        }
    }
    for (int64_t row = blockIdx.y; row < nrows; row += gridDim.y) {
        const double ox = __ldg(xp + row), oy = __ldg(yp + row), oz = __ldg(zp + row);
#pragma unroll
        for (int u = 0; u < NPASS; ++u)
            if (nlive[u])
                F[threadIdx.x + u * kGridThreads] =
                    prism_corner(__dsub_rn(nxe[u], ox), __dsub_rn(nye[u], oy), __dsub_rn(nze[u], oz));
        __syncthreads();
        double *grow = G + row * ld;
#pragma unroll
        for (int u = 0; u < CPASS; ++u) {
            if (ccol[u] < 0) continue;
            const double *f = F + cofs[u];
            // _prism.pyx:281-290: k over [z2, z1], j over [y2, y1], i over [x2, x1], sign (-1)^(i+j+k)
            double acc = 0.0;
            acc = __dadd_rn(acc, f[SZ + SY + SX]);
            acc = __dadd_rn(acc, -f[SZ + SY]);
            acc = __dadd_rn(acc, -f[SZ + SX]);
            acc = __dadd_rn(acc, f[SZ]);
            acc = __dadd_rn(acc, -f[SY + SX]);
            acc = __dadd_rn(acc, f[SY]);
            acc = __dadd_rn(acc, f[SX]);
            acc = __dadd_rn(acc, -f[0]);
            grow[ccol[u]] = __dmul_rn(acc, scale);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// tesseroid gz
// ---------------------------------------------------------------------------------------------
constexpr int kTessThreads = 128;

// COUNT = true writes the number of leaf cells of the subdivision instead of the kernel value
// (index bookkeeping of the adaptive engine, compared bit-exactly with the oracle's).
template <bool COUNT>
__global__ void __launch_bounds__(kTessThreads)
tess_gz_kernel(const double *__restrict__ lon, const double *__restrict__ sinlat,
               const double *__restrict__ coslat, const double *__restrict__ radius, int64_t nrows,
               const double *__restrict__ bounds, int64_t M, double ratio, double scale1,
               double scale2, double *__restrict__ G, int64_t ld, int32_t *__restrict__ status) {
    const int64_t col = (int64_t)blockIdx.x * kTessThreads + threadIdx.x;
    if (col >= ld) return;
    const bool live = col < M;
    TessCell root;
    if (live) {
        const double *b = bounds + 6 * col;
        root.w = b[0]; root.e = b[1]; root.s = b[2]; root.n = b[3]; root.top = b[4]; root.bottom = b[5];
    }
    TessCell stack[kStackSize];  // local memory; only touched when a cell subdivides
    int errsum = 0;
    bool overflow = false;
    // cell-only parts of the split test and of the quadrature of the un-split cell: once per thread
    TessDivC rootD;
    TessLeafC rootL;
    if (live) {
        tess_div_consts(root, ratio, rootD);
        tess_leaf_consts(root, rootL);
    }
    for (int64_t row = blockIdx.y; row < nrows; row += gridDim.y) {
        double acc = 0.0;
        if (live) {
            const double olon = __ldg(lon + row), osin = __ldg(sinlat + row),
                         ocos = __ldg(coslat + row), orad = __ldg(radius + row);
            int err;
            const int div0 = tess_div_eval(olon, ocos, osin, orad, rootD, &err);
            errsum += err;
            if (div0 == (1 | (1 << 2) | (1 << 4))) {
                // fast path: no subdivision
                acc = COUNT ? 1.0 : tess_leaf_eval(olon, ocos, osin, orad, rootL);
            } else {
                // engine (_tesseroid_numba.py:32-71): LIFO stack, children pushed lon-major
                int top = -1;
                TessCell cur = root;
                int div = div0;
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
                            acc = COUNT ? -1.0 : nan("");
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
                        acc = __dadd_rn(acc, COUNT ? 1.0 : tess_leaf(olon, ocos, osin, orad, cur));
                    }
                }
            }
            if (!COUNT) acc = __dmul_rn(__dmul_rn(acc, scale1), scale2);  // tesseroid.py:430
        }
        G[row * ld + col] = acc;
    }
    if (errsum != 0) atomicAdd(status, errsum);
    if (overflow) atomicExch(status + 1, 1);
}

// ---------------------------------------------------------------------------------------------
// sensitivity weighting
// ---------------------------------------------------------------------------------------------
constexpr int kColThreads = 128;

// One thread per 4 columns, rows summed sequentially in row order (the reference's order,
// potential.py:241-244), a*a and the add rounded separately.
__global__ void __launch_bounds__(kColThreads)
colsumsq_kernel(const double *__restrict__ G, int64_t nrows, int64_t ld, double *__restrict__ out,
                int accumulate) {
    const int64_t c4 = ((int64_t)blockIdx.x * kColThreads + threadIdx.x) * 4;
    if (c4 >= ld) return;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    if (accumulate) { s0 = out[c4]; s1 = out[c4 + 1]; s2 = out[c4 + 2]; s3 = out[c4 + 3]; }
    const double *p = G + c4;
    int64_t r = 0;
    constexpr int U = 8;
    for (; r + U <= nrows; r += U) {
        double a[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) ldg_stream4(p + (r + u) * ld, a[u][0], a[u][1], a[u][2], a[u][3]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s0 = __dadd_rn(s0, __dmul_rn(a[u][0], a[u][0]));
            s1 = __dadd_rn(s1, __dmul_rn(a[u][1], a[u][1]));
            s2 = __dadd_rn(s2, __dmul_rn(a[u][2], a[u][2]));
            s3 = __dadd_rn(s3, __dmul_rn(a[u][3], a[u][3]));
        }
    }
    for (; r < nrows; ++r) {
        double a0, a1, a2, a3;
        ldg_stream4(p + r * ld, a0, a1, a2, a3);
        s0 = __dadd_rn(s0, __dmul_rn(a0, a0));
        s1 = __dadd_rn(s1, __dmul_rn(a1, a1));
        s2 = __dadd_rn(s2, __dmul_rn(a2, a2));
        s3 = __dadd_rn(s3, __dmul_rn(a3, a3));
    }
    out[c4] = s0; out[c4 + 1] = s1; out[c4 + 2] = s2; out[c4 + 3] = s3;
}

__global__ void weights_kernel(const double *__restrict__ sumsq, int64_t M, double wf,
                               double *__restrict__ wm, double *__restrict__ wminv,
                               double *__restrict__ wmsq) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    // potential.py:245-253: ADiag = power(ADiagSquare, weightfactor); 1.0/ADiag; ADiag*ADiag
    const double s = sumsq[i];
    const double d = (wf == 0.5) ? __dsqrt_rn(s) : pow(s, wf);
    wm[i] = d;
    wminv[i] = __ddiv_rn(1.0, d);
    wmsq[i] = __dmul_rn(d, d);
}

__global__ void __launch_bounds__(kColThreads)
scale_columns_kernel(double *__restrict__ G, int64_t nrows, int64_t M, int64_t ld,
                     const double *__restrict__ cs) {
    const int64_t c4 = ((int64_t)blockIdx.x * kColThreads + threadIdx.x) * 4;
    if (c4 >= ld) return;
    double w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = (c4 + k < M) ? cs[c4 + k] : 0.0;
    for (int64_t r = blockIdx.y; r < nrows; r += gridDim.y) {
        double *p = G + r * ld + c4;
        double a0, a1, a2, a3;
        ldg4_cg(p, a0, a1, a2, a3);  // coherent load: the same addresses are stored below (.nc is for read-only data)
        stg4(p, __dmul_rn(a0, w[0]), __dmul_rn(a1, w[1]), __dmul_rn(a2, w[2]), __dmul_rn(a3, w[3]));
    }
}

}  // namespace gi

using namespace gi;

extern "C" int gi_prism_gz_assemble(const double *xp, const double *yp, const double *zp,
                                    int64_t nrows, const double *bounds, int64_t M, double scale,
                                    double *G, int64_t ld, void *stream) {
    GI_REQUIRE(nrows >= 0 && M >= 0 && ld >= M && ld % 4 == 0, "gi_prism_gz_assemble: bad shape");
    if (nrows == 0 || ld == 0) return GI_OK;
    GI_REQUIRE(xp && yp && zp && G && (bounds || M == 0), "gi_prism_gz_assemble: null pointer");
    dim3 grid((unsigned)ceil_div(ld, kAsmThreads),
              (unsigned)min((int64_t)65535, ceil_div(nrows, kAsmRows)));
    prism_gz_kernel<<<grid, kAsmThreads, 0, (cudaStream_t)stream>>>(xp, yp, zp, nrows, bounds, M,
                                                                    scale, G, ld);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

// rows are strided over gridDim.y CTAs: enough CTAs for ~16 waves of 8 CTAs per SM, so that each
// thread keeps its cell constants for as many observation rows as possible
static unsigned tess_row_split(int64_t ld, int64_t nrows) {
    const int64_t xblocks = ceil_div(ld, kTessThreads);
    int64_t y = (16LL * 8 * sm_count()) / xblocks + 1;
    y = std::max<int64_t>(1, std::min<int64_t>(y, std::min<int64_t>(nrows, 65535)));
    return (unsigned)y;
}

extern "C" int gi_prism_gz_assemble_grid(const double *xp, const double *yp, const double *zp,
                                         int64_t nrows, const double *xn, const double *yn,
                                         const double *zn, int32_t nx, int32_t ny, int32_t nz,
                                         const int32_t *colmap, int64_t M, double scale, double *G,
                                         int64_t ld, void *stream) {
    GI_REQUIRE(nrows >= 0 && M >= 0 && ld >= M && ld % 4 == 0 && nx > 0 && ny > 0 && nz > 0,
               "gi_prism_gz_assemble_grid: bad shape");
    GI_REQUIRE(colmap || (int64_t)nx * ny * nz == M,
               "gi_prism_gz_assemble_grid: M must equal nx*ny*nz without a column map");
    if (nrows == 0 || ld == 0) return GI_OK;
    GI_REQUIRE(xp && yp && zp && xn && yn && zn && G, "gi_prism_gz_assemble_grid: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    // padding columns [M, ld) (and, with a column map, nothing else) must read as zero
    if (ld > M)
        GI_CUDA(cudaMemset2DAsync(G + M, sizeof(double) * ld, 0, sizeof(double) * (ld - M), nrows, s));
    const int tiles_x = (int)ceil_div(nx, kGridTX), tiles_y = (int)ceil_div(ny, kGridTY),
              tiles_z = (int)ceil_div(nz, kGridTZ);
    const int64_t tiles = (int64_t)tiles_x * tiles_y * tiles_z;
    GI_REQUIRE(tiles < (1LL << 31), "gi_prism_gz_assemble_grid: mesh too large");
    // enough CTAs for ~16 waves; each CTA strides over observation rows
    int64_t ysplit = std::max<int64_t>(1, std::min<int64_t>(nrows, (16LL * 8 * sm_count()) / tiles + 1));
    ysplit = std::min<int64_t>(ysplit, 65535);
    dim3 grid((unsigned)tiles, (unsigned)ysplit);
    prism_gz_grid_kernel<<<grid, kGridThreads, 0, s>>>(xp, yp, zp, nrows, xn, yn, zn, nx, ny, nz,
                                                      tiles_x, tiles_y, colmap, scale, G, ld);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_tess_gz_assemble(const double *lon, const double *sinlat, const double *coslat,
                                   const double *radius, int64_t nrows, const double *bounds,
                                   int64_t M, double ratio, double scale1, double scale2, double *G,
                                   int64_t ld, int32_t *status, void *stream) {
    GI_REQUIRE(nrows >= 0 && M >= 0 && ld >= M && ld % 4 == 0, "gi_tess_gz_assemble: bad shape");
    GI_REQUIRE(ratio > 0, "gi_tess_gz_assemble: ratio must be > 0");
    if (nrows == 0 || ld == 0) return GI_OK;
    GI_REQUIRE(lon && sinlat && coslat && radius && G && status && (bounds || M == 0),
               "gi_tess_gz_assemble: null pointer");
    dim3 grid((unsigned)ceil_div(ld, kTessThreads), tess_row_split(ld, nrows));
    tess_gz_kernel<false><<<grid, kTessThreads, 0, (cudaStream_t)stream>>>(
        lon, sinlat, coslat, radius, nrows, bounds, M, ratio, scale1, scale2, G, ld, status);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_tess_gz_leafcount(const double *lon, const double *sinlat, const double *coslat,
                                    const double *radius, int64_t nrows, const double *bounds,
                                    int64_t M, double ratio, double *leaves, int64_t ld,
                                    int32_t *status, void *stream) {
    GI_REQUIRE(nrows >= 0 && M >= 0 && ld >= M && ld % 4 == 0, "gi_tess_gz_leafcount: bad shape");
    GI_REQUIRE(ratio > 0, "gi_tess_gz_leafcount: ratio must be > 0");
    if (nrows == 0 || ld == 0) return GI_OK;
    GI_REQUIRE(lon && sinlat && coslat && radius && leaves && status && (bounds || M == 0),
               "gi_tess_gz_leafcount: null pointer");
    dim3 grid((unsigned)ceil_div(ld, kTessThreads), tess_row_split(ld, nrows));
    tess_gz_kernel<true><<<grid, kTessThreads, 0, (cudaStream_t)stream>>>(
        lon, sinlat, coslat, radius, nrows, bounds, M, ratio, 1.0, 1.0, leaves, ld, status);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_colsumsq(const double *G, int64_t nrows, int64_t M, int64_t ld, double *out,
                           int accumulate, void *stream) {
    GI_REQUIRE(nrows >= 0 && M >= 0 && ld >= M && ld % 4 == 0, "gi_colsumsq: bad shape");
    if (ld == 0) return GI_OK;
    GI_REQUIRE(G && out, "gi_colsumsq: null pointer");
    colsumsq_kernel<<<(unsigned)ceil_div(ld / 4, kColThreads), kColThreads, 0,
                      (cudaStream_t)stream>>>(G, nrows, ld, out, accumulate);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_weights_from_sumsq(const double *sumsq, int64_t M, double wf, double *wm,
                                     double *wminv, double *wmsq, void *stream) {
    GI_REQUIRE(M >= 0, "gi_weights_from_sumsq: bad shape");
    if (M == 0) return GI_OK;
    GI_REQUIRE(sumsq && wm && wminv && wmsq, "gi_weights_from_sumsq: null pointer");
    weights_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(sumsq, M, wf, wm,
                                                                                 wminv, wmsq);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_scale_columns(double *G, int64_t nrows, int64_t M, int64_t ld, const double *cs,
                                void *stream) {
    GI_REQUIRE(nrows >= 0 && M >= 0 && ld >= M && ld % 4 == 0, "gi_scale_columns: bad shape");
    if (nrows == 0 || ld == 0) return GI_OK;
    GI_REQUIRE(G && cs, "gi_scale_columns: null pointer");
    dim3 grid((unsigned)ceil_div(ld / 4, kColThreads), (unsigned)min((int64_t)4096, nrows));
    scale_columns_kernel<<<grid, kColThreads, 0, (cudaStream_t)stream>>>(G, nrows, M, ld, cs);
    GI_LAUNCH_CHECK();
    return GI_OK;
}
