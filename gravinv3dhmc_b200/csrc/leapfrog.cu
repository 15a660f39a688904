// leapfrog.cu -- the steady-state hot loop of the sampler on sm_100a: d = Aw x, r, g = Aw^T r,
// regulariser gradient + leapfrog update + clamp, Metropolis test.
//
// Replaces (reference paths relative to the reference root):
//   inversion/potential.py:688-717 data_all      (np.dot(Aw, mw), mean removal, norm, np.dot(Aw.T, r))
//   inversion/potential.py:719-810 model_{MS,Damping,Smoothness,TV}_all  (+ fd3d :266-361)
//   inversion/potential.py:812-845 misfit_and_grad
//   inversion/hmc.py:85-177 _leapfrog, :44-50 _kinetic
//
// Both big passes stream the row-major FP64 matrix exactly once with 256-bit loads
// (LDG.E.256, L1 no-allocate, L2 evict-first) and reduce deterministically:
//   fwd: a CTA owns R rows x one column chunk; every x value is loaded once per thread and
//        reused for the R rows from registers; warp-shuffle + fixed-order cross-warp sum;
//        partial[chunk][row] is summed over chunks in a fixed order by the finish kernel.
//   adj: a CTA owns a 1024-column strip x one row chunk; r is staged in shared memory and
//        broadcast; each thread keeps 4 column accumulators; partial[chunk][col] is summed over
//        chunks in a fixed order by the update kernel.
// No floating-point atomics anywhere: results are bitwise reproducible run to run.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "plan.cuh"
#include "sink.cuh"

namespace gi {

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return GI_ERR_CUDA;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

// ---------------------------------------------------------------------------------------------
// forward pass  d = G x
// ---------------------------------------------------------------------------------------------
constexpr int kFwdThreads = 256;
constexpr int kFwdCols = kFwdThreads * 4;  // columns per CTA iteration

template <int R>
__global__ void __launch_bounds__(kFwdThreads, 2)
gemv_fwd_kernel(const double *__restrict__ G, int64_t ld, const double *__restrict__ x, int64_t nrows,
                int64_t chunk, int64_t rowblocks, double *__restrict__ part) {
    __shared__ double red[kFwdThreads / 32][R];
    const int64_t tile = blockIdx.x;
    const int64_t ck = tile / rowblocks, rb = tile - ck * rowblocks;
    const int64_t r0 = rb * R;
    const int64_t c0 = ck * chunk, c1 = min(c0 + chunk, ld);
    const double *rowp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) rowp[r] = G + min(r0 + r, nrows - 1) * ld;
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0;
    for (int64_t c = c0 + threadIdx.x * 4; c < c1; c += kFwdCols) {
        double xv[4], g[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) ldg_stream4(rowp[r] + c, g[r][0], g[r][1], g[r][2], g[r][3]);
        ldg4(x + c, xv[0], xv[1], xv[2], xv[3]);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            acc[r] = fma(g[r][0], xv[0], acc[r]);
            acc[r] = fma(g[r][1], xv[1], acc[r]);
            acc[r] = fma(g[r][2], xv[2], acc[r]);
            acc[r] = fma(g[r][3], xv[3], acc[r]);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const double v = warp_sum(acc[r]);
        if (lane == 0) red[warp][r] = v;
    }
    __syncthreads();
    if (threadIdx.x < R && r0 + threadIdx.x < nrows) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kFwdThreads / 32; ++w) t += red[w][threadIdx.x];
        part[ck * nrows + r0 + threadIdx.x] = t;
    }
}

// d[row] = sum over chunks (fixed order)
__global__ void fwd_reduce_kernel(const double *__restrict__ part, int64_t nrows, int64_t nchunks,
                                  double *__restrict__ d) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    double t = 0.0;
    for (int64_t c = 0; c < nchunks; ++c) t += part[c * nrows + row];
    d[row] = t;
}

// Single-CTA data-misfit kernel (potential.py:699-706).
//  mode 0: d = sum part; s0 = sum(d + fix); mean = s0/n_total; r = (d+fix-mean) - dobs_c; s1 = sum r^2
//  mode 1: s0 = sum(d + fix) only (d given)
//  mode 2: r, s1 from given d and sums[0] (already reduced over ranks)
//  mode 3: d given (e.g. the wavelet-compressed forward): s0, r, s1
constexpr int kFinThreads = 1024;
__global__ void __launch_bounds__(kFinThreads)
data_misfit_kernel(int mode, const double *__restrict__ part, int64_t nchunks, int64_t nrows,
                   int64_t n_total, double *__restrict__ d, const double *__restrict__ fix,
                   const double *__restrict__ dobs_c, double *__restrict__ r,
                   double *__restrict__ sums) {
    __shared__ double scratch[32];
    double s0 = 0.0;
    if (mode != 2) {
        for (int64_t row = threadIdx.x; row < nrows; row += kFinThreads) {
            double t;
            if (mode == 0) {
                t = 0.0;
                for (int64_t c = 0; c < nchunks; ++c) t += part[c * nrows + row];
                d[row] = t;
            } else {
                t = d[row];
            }
            s0 += fix ? t + fix[row] : t;
        }
        s0 = block_sum(s0, scratch);
        if (threadIdx.x == 0) sums[0] = s0;
        if (mode == 1) return;
    } else {
        s0 = sums[0];
    }
    const double mean = n_total > 0 ? s0 / (double)n_total : 0.0;  // n_total <= 0: no mean removal
    double s1 = 0.0;
    for (int64_t row = threadIdx.x; row < nrows; row += kFinThreads) {
        const double t = d[row];
        const double dinv = fix ? t + fix[row] : t;
        const double rr = (dinv - mean) - dobs_c[row];
        r[row] = rr;
        s1 += rr * rr;
    }
    s1 = block_sum(s1, scratch);
    if (threadIdx.x == 0) sums[1] = s1;
}

// ---------------------------------------------------------------------------------------------
// adjoint pass  g = G^T r
// ---------------------------------------------------------------------------------------------
constexpr int kAdjThreads = 256;
constexpr int kAdjCols = kAdjThreads * 4;  // columns per strip
constexpr int kAdjUnroll = 8;
constexpr int kAdjMaxRows = 2048;  // rows per chunk staged in shared memory (16 KB)

__global__ void __launch_bounds__(kAdjThreads, 2)
gemv_adj_kernel(const double *__restrict__ G, int64_t ld, const double *__restrict__ r, int64_t nrows,
                int64_t rows_per_chunk, int64_t strips, double *__restrict__ part) {
    __shared__ double rs[kAdjMaxRows];
    const int64_t tile = blockIdx.x;
    const int64_t ck = tile / strips, st = tile - ck * strips;
    const int64_t r0 = ck * rows_per_chunk;
    const int nr = (int)min(rows_per_chunk, nrows - r0);
    for (int i = threadIdx.x; i < nr; i += kAdjThreads) rs[i] = r[r0 + i];
    __syncthreads();
    const int64_t c = st * kAdjCols + threadIdx.x * 4;
    if (c >= ld) return;
    const double *p = G + r0 * ld + c;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int i = 0;
    for (; i + kAdjUnroll <= nr; i += kAdjUnroll) {
        double g[kAdjUnroll][4];
#pragma unroll
        for (int u = 0; u < kAdjUnroll; ++u)
            ldg_stream4(p + (int64_t)(i + u) * ld, g[u][0], g[u][1], g[u][2], g[u][3]);
#pragma unroll
        for (int u = 0; u < kAdjUnroll; ++u) {
            const double rv = rs[i + u];
            a0 = fma(g[u][0], rv, a0);
            a1 = fma(g[u][1], rv, a1);
            a2 = fma(g[u][2], rv, a2);
            a3 = fma(g[u][3], rv, a3);
        }
    }
    for (; i < nr; ++i) {
        double g0, g1, g2, g3;
        ldg_stream4(p + (int64_t)i * ld, g0, g1, g2, g3);
        const double rv = rs[i];
        a0 = fma(g0, rv, a0);
        a1 = fma(g1, rv, a1);
        a2 = fma(g2, rv, a2);
        a3 = fma(g3, rv, a3);
    }
    stg4(part + ck * ld + c, a0, a1, a2, a3);
}

__global__ void adj_reduce_kernel(const double *__restrict__ part, int64_t ld, int64_t nchunks,
                                  double *__restrict__ g) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ld) return;
    double t = 0.0;
    for (int64_t k = 0; k < nchunks; ++k) t += part[k * ld + c];
    g[c] = t;
}

// mw = mw(x) for the start state
__global__ void transform_kernel(const double *__restrict__ x, const double *__restrict__ low,
                                 const double *__restrict__ high, double log_factor, int64_t M,
                                 double *__restrict__ mw) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const double ex = pow(2.718281828459045, log_factor * x[j]);
    mw[j] = (low[j] + high[j] * ex) / (1.0 + ex);
}

}  // namespace gi

using namespace gi;

// =============================================================================================
// plan
// =============================================================================================
static int plan_tiles(gi_plan *p) {
    const int sms = sm_count();
    const int64_t target = 32LL * 2 * sms;  // >= 32 waves of 2 CTAs/SM: tail effect < ~3 %
    const bool small = (double)p->nrows * (double)p->ld * 8.0 < 256e6;
    // forward: R rows per tile; small problems use R = 4 to expose more tiles
    p->fwd_R = small ? 4 : 8;
    p->fwd_rowblocks = ceil_div(p->nrows, p->fwd_R);
    const int64_t min_chunk = small ? kFwdCols : 16 * kFwdCols;
    int64_t nch = ceil_div(target, p->fwd_rowblocks);
    nch = std::max<int64_t>(1, std::min<int64_t>(nch, ceil_div(p->ld, min_chunk)));
    p->fwd_chunk = ceil_div(ceil_div(p->ld, nch), kFwdCols) * kFwdCols;
    p->fwd_nchunks = ceil_div(p->ld, p->fwd_chunk);
    // adjoint: 1024-column strips x row chunks
    p->adj_strips = ceil_div(p->ld, kAdjCols);
    const int64_t min_rows = small ? 16 : 128;
    int64_t nrc = ceil_div(target, p->adj_strips);
    nrc = std::max<int64_t>(1, std::min<int64_t>(nrc, ceil_div(p->nrows, min_rows)));
    p->adj_rows = std::min<int64_t>(kAdjMaxRows, ceil_div(p->nrows, nrc));
    p->adj_rows = ceil_div(p->adj_rows, kAdjUnroll) * kAdjUnroll;
    p->adj_rows = std::min<int64_t>(kAdjMaxRows, p->adj_rows);
    p->adj_nchunks = ceil_div(p->nrows, p->adj_rows);
    p->upd_blocks = ceil_div(std::max<int64_t>(p->M, 1), (int64_t)kUpdThreads * kUpdVec * 4);
    return GI_OK;
}

extern "C" int gi_plan_create(int64_t nrows, int64_t M, int64_t ld, int32_t nchains, gi_plan **out) {
    GI_REQUIRE(out, "gi_plan_create: null out");
    GI_REQUIRE(nrows > 0 && M > 0 && ld >= M && ld % 4 == 0, "gi_plan_create: bad shape");
    GI_REQUIRE(nchains >= 1 && nchains <= 64, "gi_plan_create: 1..64 chains per plan");
    gi_plan *p = new gi_plan();
    memset(p, 0, sizeof(*p));
    p->nrows = nrows; p->M = M; p->ld = ld; p->nchains = nchains;
    plan_tiles(p);
    const size_t b_fwd = sizeof(double) * p->fwd_nchunks * nrows;
    const size_t b_adj = sizeof(double) * p->adj_nchunks * ld;
    const size_t b_blk = sizeof(double) * 3 * p->upd_blocks;
    cudaError_t e = cudaMalloc(&p->fwd_part, b_fwd);
    if (e == cudaSuccess) e = cudaMalloc(&p->adj_part, b_adj);
    if (e == cudaSuccess) e = cudaMalloc(&p->blockpart, b_blk);
    if (e == cudaSuccess) e = cudaMalloc(&p->counter, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(p->counter, 0, sizeof(unsigned int));
    if (e != cudaSuccess) {
        gi_plan_destroy(p);
        return cuda_fail(e, "plan workspace", __FILE__, __LINE__);
    }
    p->workspace_bytes = (int64_t)(b_fwd + b_adj + b_blk);
    if (nchains > 1) {
        int rc = batched_plan_init(p);
        if (rc) {
            gi_plan_destroy(p);
            return rc;
        }
    }
    *out = p;
    return GI_OK;
}

extern "C" int gi_plan_destroy(gi_plan *p) {
    if (!p) return GI_OK;
    cudaFree(p->fwd_part);
    cudaFree(p->adj_part);
    cudaFree(p->blockpart);
    cudaFree(p->counter);
    batched_plan_free(p);
    delete p;
    return GI_OK;
}

extern "C" int gi_plan_info(const gi_plan *p, int64_t *fwd_tiles, int64_t *adj_tiles,
                            int64_t *workspace_bytes) {
    GI_REQUIRE(p, "gi_plan_info: null plan");
    if (fwd_tiles) *fwd_tiles = p->fwd_nchunks * p->fwd_rowblocks;
    if (adj_tiles) *adj_tiles = p->adj_nchunks * p->adj_strips;
    if (workspace_bytes) *workspace_bytes = p->workspace_bytes;
    return GI_OK;
}

static int launch_fwd_partial(gi_plan *p, const double *G, const double *x, cudaStream_t s) {
    const unsigned grid = (unsigned)(p->fwd_nchunks * p->fwd_rowblocks);
    if (p->fwd_R == 8)
        gemv_fwd_kernel<8><<<grid, kFwdThreads, 0, s>>>(G, p->ld, x, p->nrows, p->fwd_chunk,
                                                        p->fwd_rowblocks, p->fwd_part);
    else
        gemv_fwd_kernel<4><<<grid, kFwdThreads, 0, s>>>(G, p->ld, x, p->nrows, p->fwd_chunk,
                                                        p->fwd_rowblocks, p->fwd_part);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

static int launch_adj_partial(gi_plan *p, const double *G, const double *r, cudaStream_t s) {
    const unsigned grid = (unsigned)(p->adj_nchunks * p->adj_strips);
    gemv_adj_kernel<<<grid, kAdjThreads, 0, s>>>(G, p->ld, r, p->nrows, p->adj_rows, p->adj_strips,
                                                 p->adj_part);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_gemv_fwd(gi_plan *p, const double *G, const double *x, double *d, void *stream) {
    GI_REQUIRE(p && G && x && d, "gi_gemv_fwd: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    int rc = launch_fwd_partial(p, G, x, s);
    if (rc) return rc;
    fwd_reduce_kernel<<<(unsigned)ceil_div(p->nrows, 256), 256, 0, s>>>(p->fwd_part, p->nrows,
                                                                       p->fwd_nchunks, d);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_data_sum(gi_plan *p, const double *d, const double *fix, double *sums,
                           void *stream) {
    GI_REQUIRE(p && d && sums, "gi_data_sum: null pointer");
    data_misfit_kernel<<<1, kFinThreads, 0, (cudaStream_t)stream>>>(
        1, nullptr, 0, p->nrows, p->nrows, const_cast<double *>(d), fix, nullptr, nullptr, sums);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_residual(gi_plan *p, const double *d, const double *fix, const double *dobs_c,
                           int64_t n_total, double *r, double *sums, void *stream) {
    GI_REQUIRE(p && d && dobs_c && r && sums, "gi_residual: bad argument");
    data_misfit_kernel<<<1, kFinThreads, 0, (cudaStream_t)stream>>>(
        2, nullptr, 0, p->nrows, n_total, const_cast<double *>(d), fix, dobs_c, r, sums);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_gemv_adj(gi_plan *p, const double *G, const double *r, double *g, void *stream) {
    GI_REQUIRE(p && G && r && g, "gi_gemv_adj: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    int rc = launch_adj_partial(p, G, r, s);
    if (rc) return rc;
    adj_reduce_kernel<<<(unsigned)ceil_div(p->ld, 256), 256, 0, s>>>(p->adj_part, p->ld,
                                                                    p->adj_nchunks, g);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

int gi::check_reg(const gi_reg_params *reg, int64_t M) {
    GI_REQUIRE(reg, "null regulariser parameters");
    GI_REQUIRE(reg->reg_kind >= GI_REG_DAMPING && reg->reg_kind <= GI_REG_TV,
               "Please choose regularization from 'MS','Damping', 'Smoothness', 'TV'.");
    GI_REQUIRE(reg->constraint == GI_CONSTRAINT_MANDATORY ||
                   reg->constraint == GI_CONSTRAINT_LOGARITHMIC,
               "Please choose right boundary constraint(mandatory, logarithmic)!");
    if (reg->reg_kind == GI_REG_SMOOTHNESS || reg->reg_kind == GI_REG_TV)
        GI_REQUIRE((int64_t)reg->nz * reg->ny * reg->nx == M,
                   "Smoothness/TV need the full (nz, ny, nx) grid: nz*ny*nx != M");
    return GI_OK;
}

static int launch_update(gi_plan *p, const gi_reg_params *reg, const double *grad_in,
                         const double *gpart, int64_t gparts, const double *x_in,
                         const double *mw_in, const double *mwapr, const double *wmsq,
                         const double *low, const double *high, double *pm, double *x_out,
                         double *mw_out, double *grad_out, double pcoef, double dt, int advance,
                         double *sums, cudaStream_t s) {
    UpdateArgs a;
    a.save_k0 = 0;
    a.p_in = nullptr;
    a.gp_ldk = 0;
    a.gp_piece_stride = 0;
    a.col0 = a.col1 = 0;
    a.gp_nsrc = a.gp_src_stride = a.gp_pitch = 0;
    a.grad_in = grad_in; a.gpart = gpart; a.gparts = gparts;
    a.x_in = x_in; a.mw_in = mw_in; a.mwapr = mwapr; a.wmsq = wmsq; a.low = low; a.high = high;
    a.p = pm; a.x_out = x_out; a.mw_out = mw_out; a.grad_out = grad_out;
    a.pcoef = pcoef; a.dt = dt; a.advance = advance; a.M = p->M; a.ld = p->ld; a.reg = *reg;
    a.blockpart = p->blockpart; a.counter = p->counter; a.sums = sums;
    update_kernel<<<(unsigned)p->upd_blocks, kUpdThreads, 0, s>>>(a);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_update(gi_plan *p, const gi_reg_params *reg, const double *gdata,
                         const double *x_in, const double *mw_in, const double *mwapr,
                         const double *wmsq, const double *low, const double *high, double *pm,
                         double *x_out, double *mw_out, double *grad_out, double pcoef, double dt,
                         int advance, double *sums, void *stream) {
    GI_REQUIRE(p && gdata && x_in && mw_in && mwapr && pm && sums, "gi_update: null pointer");
    int rc = check_reg(reg, p->M);
    if (rc) return rc;
    GI_REQUIRE(reg->reg_kind != GI_REG_MS || wmsq, "gi_update: MS needs wmsq");
    GI_REQUIRE(!advance || (x_out && mw_out && low && high && x_out != x_in),
               "gi_update: advance needs x_out/mw_out/low/high");
    return launch_update(p, reg, nullptr, gdata, 1, x_in, mw_in, mwapr, wmsq, low, high, pm, x_out,
                         mw_out, grad_out, pcoef, dt, advance, sums, (cudaStream_t)stream);
}

// =============================================================================================
// single-GPU device-resident sampler
// =============================================================================================
struct gi_hmc {
    gi_hmc_config cfg;
    gi_plan *plan;
    const double *G;
    cudaStream_t stream;
    // M-vectors (ld entries, zero padded)
    double *x_cur, *mw_cur, *g_cur, *xa, *xb, *mwa, *mwb, *p, *gnew, *low, *high, *mwapr, *wmsq;
    // N-vectors
    double *d_cur, *d, *r, *dobs_c, *fix;
    double *sums;  // [8]
    DevState *st;
    DevState *st_host;  // pinned
    bool has_state;
    int64_t launches;
    // wavelet-compressed forward (0 = off)
    int wv_kind, wv_nz, wv_ny, wv_nx;
    const int64_t *wv_indptr;
    const int32_t *wv_indices;
    const double *wv_data;
    int64_t wv_ncoef;
    double *wv_coef;
    // single-pass gradient evaluation (fused.cu): -1 not tried yet, 0 unavailable, 1 in use
    int fused_state;
    gi_fused *fused;
    double *gfused;
    // on-device sample sink (gi_hmc_attach_stats)
    gi_stats *stats;
    int32_t stats_slot;
    const double *stats_scale;
};

static void hmc_free(gi_hmc *h) {
    if (!h) return;
    double *bufs[] = {h->x_cur, h->g_cur, h->xa, h->xb, h->p, h->gnew, h->low, h->high, h->mwapr,
                      h->wmsq, h->d_cur, h->d, h->r, h->dobs_c, h->fix, h->sums};
    for (double *b : bufs) cudaFree(b);
    if (h->cfg.reg.constraint == GI_CONSTRAINT_LOGARITHMIC) {
        cudaFree(h->mw_cur); cudaFree(h->mwa); cudaFree(h->mwb);
    }
    cudaFree(h->st);
    cudaFree(h->wv_coef);
    cudaFree(h->gfused);
    gi_fused_destroy(h->fused);
    if (h->st_host) cudaFreeHost(h->st_host);
    gi_plan_destroy(h->plan);
    delete h;
}

// The fused single-pass evaluation pays one cross-SM hand-off per observation row (~1 us), so it only
// wins when a row is worth that: the two GEMV passes stream a row in 16 M bytes / 7.3 TB/s, which
// exceeds the hand-off above M ~ 460 000 voxels (measured: c5, M = 2^20: 27.0 vs 37.4 ms per
// evaluation; mid, M = 2^17: 4.1 vs 1.2 ms).  Default: rows of >= 4 MB (M >= 524 288) and kernels of
// >= 1 GB (GI_FUSED_GEMV=1 forces it on for any shape that fits, =0 switches it off).  The two-pass
// kernels (1.12 x the measured copy bandwidth) serve every other shape, including strips that do not
// fit one SM's shared memory (M > 148 x 7168); gi_hmc_eval_path reports which one is in use.
static bool fused_ready(gi_hmc *h) {
    if (h->fused_state >= 0) return h->fused_state == 1;
    h->fused_state = 0;
    const char *env = getenv("GI_FUSED_GEMV");
    if (env && env[0] == '0') return false;
    const bool force = env && env[0] == '1';
    if (!force && ((double)h->cfg.N * (double)h->cfg.ld * 8.0 < 1e9 || h->cfg.M < (1 << 19))) return false;
    if (gi_fused_create(h->cfg.N, h->cfg.M, h->cfg.ld, h->G, h->stream, &h->fused) != GI_OK) {
        h->fused = nullptr;
        return false;
    }
    if (cudaMalloc(&h->gfused, sizeof(double) * h->cfg.ld) != cudaSuccess) {
        gi_fused_destroy(h->fused);
        h->fused = nullptr;
        return false;
    }
    h->fused_state = 1;
    return true;
}

#define HMC_CUDA(h, call)                                                         \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) {                                                 \
            hmc_free(h);                                                          \
            return gi::cuda_fail(e__, #call, __FILE__, __LINE__);                 \
        }                                                                         \
    } while (0)

extern "C" int gi_hmc_create(const gi_hmc_config *cfg, const double *G, const double *dobs_host,
                             const double *gravfix_host, const double *low_host,
                             const double *high_host, const double *mwapr_host,
                             const double *wmsq_host, void *stream, gi_hmc **out) {
    GI_REQUIRE(cfg && G && dobs_host && low_host && high_host && mwapr_host && out,
               "gi_hmc_create: null pointer");
    GI_REQUIRE(cfg->N > 0 && cfg->M > 0 && cfg->ld >= cfg->M && cfg->ld % 4 == 0,
               "gi_hmc_create: bad shape");
    int rc = check_reg(&cfg->reg, cfg->M);
    if (rc) return rc;
    GI_REQUIRE(!cfg->fixed || gravfix_host, "gi_hmc_create: fixed=True needs grav_fix");
    gi_hmc *h = new gi_hmc();
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->G = G;
    h->stream = (cudaStream_t)stream;
    h->fused_state = -1;
    rc = gi_plan_create(cfg->N, cfg->M, cfg->ld, 1, &h->plan);
    if (rc) { delete h; return rc; }
    const size_t bm = sizeof(double) * cfg->ld, bn = sizeof(double) * cfg->N;
    double **mv[] = {&h->x_cur, &h->g_cur, &h->xa, &h->xb, &h->p, &h->gnew, &h->low, &h->high,
                     &h->mwapr, &h->wmsq};
    for (double **b : mv) {
        HMC_CUDA(h, cudaMalloc(b, bm));
        HMC_CUDA(h, cudaMemsetAsync(*b, 0, bm, h->stream));
    }
    if (cfg->reg.constraint == GI_CONSTRAINT_LOGARITHMIC) {
        double **lv[] = {&h->mw_cur, &h->mwa, &h->mwb};
        for (double **b : lv) {
            HMC_CUDA(h, cudaMalloc(b, bm));
            HMC_CUDA(h, cudaMemsetAsync(*b, 0, bm, h->stream));
        }
    } else {
        h->mw_cur = h->x_cur; h->mwa = h->xa; h->mwb = h->xb;
    }
    double **nv[] = {&h->d_cur, &h->d, &h->r, &h->dobs_c, &h->fix};
    for (double **b : nv) {
        HMC_CUDA(h, cudaMalloc(b, bn));
        HMC_CUDA(h, cudaMemsetAsync(*b, 0, bn, h->stream));
    }
    HMC_CUDA(h, cudaMalloc(&h->sums, sizeof(double) * 8));
    HMC_CUDA(h, cudaMemsetAsync(h->sums, 0, sizeof(double) * 8, h->stream));
    HMC_CUDA(h, cudaMalloc(&h->st, sizeof(DevState)));
    HMC_CUDA(h, cudaMemsetAsync(h->st, 0, sizeof(DevState), h->stream));
    HMC_CUDA(h, cudaMallocHost(&h->st_host, sizeof(DevState)));
    const size_t vm = sizeof(double) * cfg->M;
    HMC_CUDA(h, cudaMemcpyAsync(h->low, low_host, vm, cudaMemcpyHostToDevice, h->stream));
    HMC_CUDA(h, cudaMemcpyAsync(h->high, high_host, vm, cudaMemcpyHostToDevice, h->stream));
    HMC_CUDA(h, cudaMemcpyAsync(h->mwapr, mwapr_host, vm, cudaMemcpyHostToDevice, h->stream));
    if (wmsq_host)
        HMC_CUDA(h, cudaMemcpyAsync(h->wmsq, wmsq_host, vm, cudaMemcpyHostToDevice, h->stream));
    // dobs - mean(dobs) (potential.py:706), mean in plain sequential-pairwise double on the host
    {
        double *tmp = new double[cfg->N];
        long double acc = 0.0L;
        for (int64_t i = 0; i < cfg->N; ++i) acc += dobs_host[i];
        const double mean = cfg->nocenter ? 0.0 : (double)(acc / (long double)cfg->N);
        for (int64_t i = 0; i < cfg->N; ++i) tmp[i] = dobs_host[i] - mean;
        cudaError_t e = cudaMemcpyAsync(h->dobs_c, tmp, bn, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        delete[] tmp;
        HMC_CUDA(h, e);
    }
    if (cfg->fixed)
        HMC_CUDA(h, cudaMemcpyAsync(h->fix, gravfix_host, bn, cudaMemcpyHostToDevice, h->stream));
    HMC_CUDA(h, cudaStreamSynchronize(h->stream));
    *out = h;
    return GI_OK;
}

extern "C" int gi_hmc_destroy(gi_hmc *h) {
    if (h) cudaStreamSynchronize(h->stream);
    hmc_free(h);
    return GI_OK;
}

extern "C" int gi_hmc_set_reg(gi_hmc *h, const gi_reg_params *reg) {
    GI_REQUIRE(h && reg, "gi_hmc_set_reg: null pointer");
    int rc = check_reg(reg, h->cfg.M);
    if (rc) return rc;
    GI_REQUIRE(reg->constraint == h->cfg.reg.constraint,
               "gi_hmc_set_reg: the constraint is fixed at creation");
    h->cfg.reg = *reg;
    h->has_state = false;  // U, grad depend on the regulariser
    return GI_OK;
}

// one misfit_and_grad evaluation at (x_in, mw_in) followed by the fused update
static int grad_eval_and_update(gi_hmc *h, const double *x_in, const double *mw_in, double *x_out,
                                double *mw_out, double *grad_out, double pcoef, double dt,
                                int advance) {
    gi_plan *p = h->plan;
    cudaStream_t s = h->stream;
    const int64_t ncen = h->cfg.nocenter ? 0 : p->nrows;  // rows the mean is taken over (0: no mean removal)
    int rc;
    if (h->wv_kind) {
        // d = Awcp @ DWT(mw)  (compressor1D.py:45-60 / compressor3D.py:47-68)
        if (h->wv_kind == 1)
            rc = gi_dwt_db4_l2_1d(mw_in, h->cfg.M, h->wv_coef, nullptr, s);
        else
            rc = gi_dwt_db4_l2_3d(mw_in, h->wv_nz, h->wv_ny, h->wv_nx, h->wv_coef, nullptr, s);
        if (rc) return rc;
        rc = gi_csr_spmv(h->wv_indptr, h->wv_indices, h->wv_data, p->nrows, h->wv_coef, h->d, s);
        if (rc) return rc;
        data_misfit_kernel<<<1, kFinThreads, 0, s>>>(3, nullptr, 0, p->nrows, ncen, h->d,
                                                     h->cfg.fixed ? h->fix : nullptr, h->dobs_c,
                                                     h->r, h->sums);
        h->launches += (h->wv_kind == 1 ? 2 : 7);
    } else if (fused_ready(h)) {
        // d, Aw^T r in ONE pass over Aw (fused.cu); then U_data from d, and the usual fused update
        rc = gi_fused_pass(h->fused, mw_in, h->dobs_c, h->cfg.fixed ? h->fix : nullptr, h->cfg.nocenter ? 0 : 1,
                           h->d, h->gfused, s);
        if (rc) return rc;
        data_misfit_kernel<<<1, kFinThreads, 0, s>>>(3, nullptr, 0, p->nrows, ncen, h->d,
                                                     h->cfg.fixed ? h->fix : nullptr, h->dobs_c,
                                                     h->r, h->sums);
        GI_LAUNCH_CHECK();
        rc = launch_update(p, &h->cfg.reg, nullptr, h->gfused, 1, x_in, mw_in, h->mwapr, h->wmsq, h->low,
                           h->high, h->p, x_out, mw_out, grad_out, pcoef, dt, advance, h->sums, s);
        h->launches += 3;
        return rc;
    } else {
        rc = launch_fwd_partial(p, h->G, mw_in, s);
        if (rc) return rc;
        data_misfit_kernel<<<1, kFinThreads, 0, s>>>(0, p->fwd_part, p->fwd_nchunks, p->nrows,
                                                     ncen, h->d, h->cfg.fixed ? h->fix : nullptr,
                                                     h->dobs_c, h->r, h->sums);
    }
    GI_LAUNCH_CHECK();
    rc = launch_adj_partial(p, h->G, h->r, s);
    if (rc) return rc;
    rc = launch_update(p, &h->cfg.reg, nullptr, p->adj_part, p->adj_nchunks, x_in, mw_in, h->mwapr,
                       h->wmsq, h->low, h->high, h->p, x_out, mw_out, grad_out, pcoef, dt, advance,
                       h->sums, s);
    h->launches += 4;
    return rc;
}

extern "C" int gi_hmc_set_state(gi_hmc *h, const double *x_host) {
    GI_REQUIRE(h && x_host, "gi_hmc_set_state: null pointer");
    cudaStream_t s = h->stream;
    const size_t vm = sizeof(double) * h->cfg.M;
    GI_CUDA(cudaMemcpyAsync(h->x_cur, x_host, vm, cudaMemcpyHostToDevice, s));
    if (h->cfg.reg.constraint == GI_CONSTRAINT_LOGARITHMIC) {
        transform_kernel<<<(unsigned)ceil_div(h->cfg.M, 256), 256, 0, s>>>(
            h->x_cur, h->low, h->high, h->cfg.reg.log_factor, h->cfg.M, h->mw_cur);
        GI_LAUNCH_CHECK();
        h->launches += 1;
    }
    GI_CUDA(cudaMemsetAsync(h->p, 0, sizeof(double) * h->cfg.ld, s));
    int rc = grad_eval_and_update(h, h->x_cur, h->mw_cur, nullptr, nullptr, h->g_cur, 0.0, 0.0, 0);
    if (rc) return rc;
    // commit U, Ud, Um of the start state: a forced-accept Metropolis with K = 0
    metropolis_kernel<<<1, 1, 0, s>>>(h->st, h->sums, h->cfg.reg.alpha, 0, 1);
    GI_LAUNCH_CHECK();
    GI_CUDA(cudaMemcpyAsync(h->d_cur, h->d, sizeof(double) * h->cfg.N, cudaMemcpyDeviceToDevice, s));
    h->launches += 1;
    GI_CUDA(cudaStreamSynchronize(s));
    h->has_state = true;
    return GI_OK;
}

extern "C" int gi_hmc_get_state(gi_hmc *h, double *x_host, double *d_host, double *mw_host) {
    GI_REQUIRE(h && h->has_state, "gi_hmc_get_state: no state set");
    cudaStream_t s = h->stream;
    if (x_host)
        GI_CUDA(cudaMemcpyAsync(x_host, h->x_cur, sizeof(double) * h->cfg.M, cudaMemcpyDeviceToHost, s));
    if (mw_host)
        GI_CUDA(cudaMemcpyAsync(mw_host, h->mw_cur, sizeof(double) * h->cfg.M, cudaMemcpyDeviceToHost, s));
    if (d_host)
        GI_CUDA(cudaMemcpyAsync(d_host, h->d_cur, sizeof(double) * h->cfg.N, cudaMemcpyDeviceToHost, s));
    GI_CUDA(cudaStreamSynchronize(s));
    return GI_OK;
}

extern "C" int gi_hmc_get_misfit(gi_hmc *h, double *U, double *Ud, double *Um, double *grad_host) {
    GI_REQUIRE(h && h->has_state, "gi_hmc_get_misfit: no state set");
    cudaStream_t s = h->stream;
    GI_CUDA(cudaMemcpyAsync(h->st_host, h->st, sizeof(DevState), cudaMemcpyDeviceToHost, s));
    if (grad_host)
        GI_CUDA(cudaMemcpyAsync(grad_host, h->g_cur, sizeof(double) * h->cfg.M, cudaMemcpyDeviceToHost, s));
    GI_CUDA(cudaStreamSynchronize(s));
    if (U) *U = h->st_host->U;
    if (Ud) *Ud = h->st_host->Ud;
    if (Um) *Um = h->st_host->Um;
    return GI_OK;
}

// trajectory from the current state with the momentum already in h->p
static int run_trajectory(gi_hmc *h, int32_t L, double dt, gi_hmc_result *result,
                          double *trace_x_host, double *trace_U_host, bool metropolis) {
    cudaStream_t s = h->stream;
    gi_plan *p = h->plan;
    const int64_t M = h->cfg.M;
    const bool logc = h->cfg.reg.constraint == GI_CONSTRAINT_LOGARITHMIC;
    const double alpha = h->cfg.reg.alpha;
    double hs[8];
    if (trace_x_host)
        GI_CUDA(cudaMemcpyAsync(trace_x_host, h->x_cur, sizeof(double) * M, cudaMemcpyDeviceToHost, s));
    if (trace_U_host) {
        GI_CUDA(cudaMemcpyAsync(h->st_host, h->st, sizeof(DevState), cudaMemcpyDeviceToHost, s));
        GI_CUDA(cudaStreamSynchronize(s));
        trace_U_host[0] = h->st_host->U;
    }
    // hmc.py:104-118: K0 = 0.5 p.p ; p -= 0.5 dt grad(x_cur) ; x += dt p ; clamp
    int rc = launch_update(p, &h->cfg.reg, h->g_cur, nullptr, 0, h->x_cur, h->mw_cur, h->mwapr,
                           h->wmsq, h->low, h->high, h->p, h->xa, h->mwa, nullptr, 0.5 * dt, dt, 1,
                           h->sums, s);
    if (rc) return rc;
    h->launches += 1;
    // K0 lives in sums[4]; later updates overwrite it, so park it in sums[5]
    GI_CUDA(cudaMemcpyAsync(h->sums + 5, h->sums + 4, sizeof(double), cudaMemcpyDeviceToDevice, s));
    double *xin = h->xa, *xout = h->xb, *mwin = h->mwa, *mwout = h->mwb;
    for (int i = 1; i <= L; ++i) {
        const bool last = (i == L) && metropolis;
        rc = grad_eval_and_update(h, xin, mwin, xout, mwout, last ? h->gnew : nullptr,
                                  last ? 0.5 * dt : dt, dt, last ? 0 : 1);
        if (rc) return rc;
        if (trace_x_host)
            GI_CUDA(cudaMemcpyAsync(trace_x_host + (int64_t)i * M, xin, sizeof(double) * M,
                                    cudaMemcpyDeviceToHost, s));
        if (trace_U_host) {
            GI_CUDA(cudaMemcpyAsync(hs, h->sums, sizeof(hs), cudaMemcpyDeviceToHost, s));
            GI_CUDA(cudaStreamSynchronize(s));
            trace_U_host[i] = hs[1] + alpha * hs[2];
        }
        if (!last) {
            double *t = xin; xin = xout; xout = t;
            if (logc) { t = mwin; mwin = mwout; mwout = t; }
            else { mwin = xin; mwout = xout; }
        }
    }
    if (!metropolis) return GI_OK;
    GI_CUDA(cudaMemcpyAsync(h->sums + 4, h->sums + 5, sizeof(double), cudaMemcpyDeviceToDevice, s));
    metropolis_kernel<<<1, 1, 0, s>>>(h->st, h->sums, alpha, L, 0);
    GI_LAUNCH_CHECK();
    const int64_t n = std::max(M, h->cfg.N);
    commit_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(h->st, M, h->cfg.N, xin, mwin, h->gnew,
                                                            h->d, h->x_cur, h->mw_cur, h->g_cur,
                                                            h->d_cur);
    GI_LAUNCH_CHECK();
    h->launches += 2;
    if (h->stats) {  // the accepted model goes to the sink without leaving the device
        StatsMap map;
        memset(&map, 0, sizeof(map));
        map.fin[0] = 1;
        rc = stats_add_chains(h->stats, map, h->st, h->mw_cur, h->cfg.ld, 1, h->stats_slot,
                              h->stats_scale, s);
        if (rc) return rc;
        h->launches += 2;
    }
    GI_CUDA(cudaMemcpyAsync(h->st_host, h->st, sizeof(DevState), cudaMemcpyDeviceToHost, s));
    GI_CUDA(cudaStreamSynchronize(s));
    if (result) *result = h->st_host->res;
    return GI_OK;
}

extern "C" int gi_hmc_attach_stats(gi_hmc *h, gi_stats *stats, int32_t slot, const double *scale_dev) {
    GI_REQUIRE(h, "gi_hmc_attach_stats: null handle");
    h->stats = stats;
    h->stats_slot = slot;
    h->stats_scale = scale_dev;
    return GI_OK;
}

extern "C" int gi_hmc_propose(gi_hmc *h, const double *p0_host, int32_t L, double dt, double u,
                              gi_hmc_result *result, double *trace_x_host, double *trace_U_host) {
    GI_REQUIRE(h && p0_host && result, "gi_hmc_propose: null pointer");
    GI_REQUIRE(h->has_state, "gi_hmc_propose: call gi_hmc_set_state first");
    GI_REQUIRE(L >= 1, "gi_hmc_propose: L must be >= 1");
    cudaStream_t s = h->stream;
    GI_CUDA(cudaMemcpyAsync(h->p, p0_host, sizeof(double) * h->cfg.M, cudaMemcpyHostToDevice, s));
    GI_CUDA(cudaMemcpyAsync(&h->st->u, &u, sizeof(double), cudaMemcpyHostToDevice, s));
    return run_trajectory(h, L, dt, result, trace_x_host, trace_U_host, true);
}

extern "C" int gi_hmc_propose_philox(gi_hmc *h, uint64_t seed, uint64_t counter, double sigma,
                                     int32_t L, double dt, gi_hmc_result *result) {
    GI_REQUIRE(h && result, "gi_hmc_propose_philox: null pointer");
    GI_REQUIRE(h->has_state, "gi_hmc_propose_philox: call gi_hmc_set_state first");
    GI_REQUIRE(L >= 1, "gi_hmc_propose_philox: L must be >= 1");
    const int64_t pairs = ceil_div(h->cfg.M, 2);
    philox_normal_kernel<<<(unsigned)ceil_div(pairs, 256), 256, 0, h->stream>>>(
        seed, counter, sigma, h->cfg.M, h->p, h->st);
    GI_LAUNCH_CHECK();
    h->launches += 1;
    return run_trajectory(h, L, dt, result, nullptr, nullptr, true);
}

extern "C" int gi_hmc_leapfrog_steps(gi_hmc *h, const double *p0_dev, int32_t nsteps, double dt) {
    GI_REQUIRE(h, "gi_hmc_leapfrog_steps: null handle");
    GI_REQUIRE(h->has_state, "gi_hmc_leapfrog_steps: call gi_hmc_set_state first");
    GI_REQUIRE(nsteps >= 1, "gi_hmc_leapfrog_steps: nsteps must be >= 1");
    cudaStream_t s = h->stream;
    if (p0_dev)
        GI_CUDA(cudaMemcpyAsync(h->p, p0_dev, sizeof(double) * h->cfg.M, cudaMemcpyDeviceToDevice, s));
    else
        GI_CUDA(cudaMemsetAsync(h->p, 0, sizeof(double) * h->cfg.ld, s));
    return run_trajectory(h, nsteps, dt, nullptr, nullptr, nullptr, false);
}

extern "C" int gi_hmc_set_wavelet(gi_hmc *h, int32_t kind, int32_t nz, int32_t ny, int32_t nx,
                                  const int64_t *indptr, const int32_t *indices, const double *data,
                                  int64_t ncoef) {
    GI_REQUIRE(h, "gi_hmc_set_wavelet: null handle");
    GI_REQUIRE(kind == 0 || kind == 1 || kind == 3, "gi_hmc_set_wavelet: kind must be 0, 1 or 3");
    h->has_state = false;
    cudaFree(h->wv_coef);
    h->wv_coef = nullptr;
    h->wv_kind = 0;
    if (kind == 0) return GI_OK;
    GI_REQUIRE(indptr && indices && data && ncoef > 0, "gi_hmc_set_wavelet: null CSR arrays");
    int64_t want = 0;
    if (kind == 1) {
        int rc = gi_dwt_db4_l2_1d(nullptr, h->cfg.M, nullptr, &want, nullptr);
        if (rc) return rc;
    } else {
        GI_REQUIRE((int64_t)nz * ny * nx == h->cfg.M,
                   "gi_hmc_set_wavelet: the 3-D wavelet needs the full (nz, ny, nx) grid");
        int32_t shp[3];
        int rc = gi_dwt_db4_l2_3d(nullptr, nz, ny, nx, nullptr, shp, nullptr);
        if (rc) return rc;
        want = (int64_t)shp[0] * shp[1] * shp[2];
    }
    GI_REQUIRE(want == ncoef, "gi_hmc_set_wavelet: CSR column count does not match the transform");
    GI_CUDA(cudaMalloc(&h->wv_coef, sizeof(double) * ncoef));
    h->wv_kind = kind; h->wv_nz = nz; h->wv_ny = ny; h->wv_nx = nx;
    h->wv_indptr = indptr; h->wv_indices = indices; h->wv_data = data; h->wv_ncoef = ncoef;
    return GI_OK;
}

extern "C" int64_t gi_hmc_launch_count(const gi_hmc *h) { return h ? h->launches : 0; }
// 0: the two GEMV passes, 1: the single-pass evaluation (fused.cu), 2: wavelet-compressed forward
extern "C" int32_t gi_hmc_eval_path(gi_hmc *h) {
    if (!h) return -1;
    if (h->wv_kind) return 2;
    return fused_ready(h) ? 1 : 0;
}
extern "C" void *gi_hmc_stream(const gi_hmc *h) { return h ? (void *)h->stream : nullptr; }

// =============================================================================================
// misc
// =============================================================================================
extern "C" int gi_abi_version(void) { return GI_ABI_VERSION; }
extern "C" const char *gi_last_error(void) { return gi::g_err; }

extern "C" int gi_device_info(int *sms, int *cc_major, int *cc_minor, int64_t *l2_bytes) {
    int dev = 0;
    GI_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    GI_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sms) *sms = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (l2_bytes) *l2_bytes = prop.l2CacheSize;
    return GI_OK;
}
