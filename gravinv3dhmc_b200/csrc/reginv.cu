// reginv.cu -- device-resident regularised conjugate gradient and its bootstrap
// (reference inversion/reginv.py: ConjugateGradient.CG :357-491,
// BootStrap.CG :631-713, BootStrap.BSCG :715-748).
//
// One handle runs `ncols` independent CG problems in lockstep over ONE weighted kernel Aw:
//   ncols == 1  : the reference's ConjugateGradient.CG, G passes through the GEMV kernels (HBM bound)
//   ncols 2..64 : bootstrap replicates batched as columns of the DMMA contractions (batched.cu).
// A bootstrap replicate resamples the observation rows WITH replacement (reginv.py:733-739); instead
// of gathering rows into a second N x M matrix, replicate c carries the multiplicity w_c[l] of
// row l:  |AwS m - dS|^2 = sum_l w_l (Aw_l m - d_l)^2,  AwS^T rS = Aw^T (w .* r),
// |AwS Iw|^2 = sum_l w_l (Aw_l Iw)^2  -- the same sums in a different order, so Aw is streamed once
// for all replicates and never copied.
//
// Per iteration and column (reference lines in the kernels):  alpha control -> Gt = Aw^T(w r) ->
// I = 2 Gt + alpha dR -> mu, Iw = I + mu Iw -> Q = Aw Iw -> kstep -> mw = Wm clamp(WmInv (mw -
// kstep Iw)) -> R(mw), dR(mw) -> D = Aw mw -> r, data.  Three passes over Aw per iteration
// (the reference makes ~12: it re-evaluates data(mw) for every use).  All scalars (alpha, mu,
// kstep, norms) stay on the device; the host reads one 4-byte "columns still active" word per
// iteration for the reference's early stop.
#include <algorithm>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "plan.cuh"

using namespace gi;

namespace {

enum { S_ALPHA = 0, S_MU, S_KSTEP, S_DATA_CUR, S_DATA_NEW, S_MODEL_NEW, S_II, S_II_OLD, S_IWI, S_IWIW,
       S_QQ, S_ACTIVE, S_ITERS, S_STRIDE = 16 };

constexpr int kVecThreads = 256;

struct CgDims {
    int64_t M, ld, N, npad;
    int32_t C, variant;
    gi_reg_params reg;
};

// deterministic grid reduction of up to 3 values per column: per-CTA partials, the last CTA of the
// column adds them in index order (same scheme as update_body in plan.cuh)
__device__ __forceinline__ void column_reduce3(double v0, double v1, double v2, double *blockpart,
                                               unsigned int *counter, double *out0, double *out1,
                                               double *out2) {
    __shared__ double scratch[32];
    __shared__ bool is_last;
    const int64_t nb = gridDim.x;
    double *bp = blockpart + (int64_t)blockIdx.y * nb * 3;
    v0 = block_sum(v0, scratch);
    v1 = block_sum(v1, scratch);
    v2 = block_sum(v2, scratch);
    if (threadIdx.x == 0) {
        bp[3 * (int64_t)blockIdx.x + 0] = v0;
        bp[3 * (int64_t)blockIdx.x + 1] = v1;
        bp[3 * (int64_t)blockIdx.x + 2] = v2;
        __threadfence();
        is_last = atomicAdd(counter + blockIdx.y, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double s0 = 0, s1 = 0, s2 = 0;
    for (int64_t b = threadIdx.x; b < nb; b += blockDim.x) {
        s0 += __ldcg(bp + 3 * b + 0);
        s1 += __ldcg(bp + 3 * b + 1);
        s2 += __ldcg(bp + 3 * b + 2);
    }
    s0 = block_sum(s0, scratch);
    s1 = block_sum(s1, scratch);
    s2 = block_sum(s2, scratch);
    if (threadIdx.x == 0) {
        if (out0) *out0 = s0;
        if (out1) *out1 = s1;
        if (out2) *out2 = s2;
        counter[blockIdx.y] = 0;
    }
}

// numpy's  np.linalg.norm(v) ** 2  (reginv.py:256,425,...): square of the rounded square root
__device__ __forceinline__ double norm_sq(double sumsq) {
    const double n = sqrt(sumsq);
    return __dmul_rn(n, n);
}

// regularisation-factor control, one thread per column (reginv.py:384-402 / 648-658)
__global__ void cg_alpha_kernel(double *S, int C, int k, double q, double *hist_alpha, int maxk) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double *s = S + (int64_t)c * S_STRIDE;
    if (s[S_ACTIVE] == 0.0) return;
    double alpha = s[S_ALPHA];
    if (k == 0) alpha = 0.0;
    else if (k == 1) alpha = s[S_DATA_NEW] / s[S_MODEL_NEW];
    else if (__dsub_rn(s[S_DATA_CUR], s[S_DATA_NEW]) < __dmul_rn(0.01, s[S_DATA_CUR])) alpha = __dmul_rn(q, alpha);
    s[S_ALPHA] = alpha;
    hist_alpha[(int64_t)c * maxk + k] = alpha;
    if (k > 0) {  // mw = mw_new, I_old = I  (reginv.py:441-444)
        s[S_DATA_CUR] = s[S_DATA_NEW];
        s[S_II_OLD] = s[S_II];
    }
}

// I = 2 Aw^T r + alpha dR(mw)   (reginv.py:411-421, 446-455);  S_II = sum I^2
__global__ void __launch_bounds__(kVecThreads)
cg_grad_kernel(CgDims dm, const double *__restrict__ Gt, const double *__restrict__ gR,
               double *__restrict__ I, double *S, double *blockpart, unsigned int *counter) {
    const int c = blockIdx.y;
    double *s = S + (int64_t)c * S_STRIDE;
    if (s[S_ACTIVE] == 0.0) return;
    const double alpha = s[S_ALPHA];
    const int64_t off = (int64_t)c * dm.ld;
    double acc = 0.0;
    const int64_t j = (int64_t)blockIdx.x * kVecThreads + threadIdx.x;
    if (j < dm.M) {
        const double v = __dadd_rn(__dmul_rn(2.0, Gt[off + j]), __dmul_rn(alpha, gR[off + j]));
        I[off + j] = v;
        acc = v * v;
    }
    column_reduce3(acc, 0.0, 0.0, blockpart, counter, s + S_II, nullptr, nullptr);
}

// mu = |I|^2 / |I_old|^2;  Iw = I + mu Iw_old  (reginv.py:456-458; k == 0: Iw = I, :423)
// S_IWI = Iw.I, S_IWIW = sum Iw^2
__global__ void __launch_bounds__(kVecThreads)
cg_dir_kernel(CgDims dm, int k, const double *__restrict__ I, double *__restrict__ Iw, double *S,
              double *blockpart, unsigned int *counter) {
    const int c = blockIdx.y;
    double *s = S + (int64_t)c * S_STRIDE;
    if (s[S_ACTIVE] == 0.0) return;
    const double mu = k == 0 ? 0.0 : norm_sq(s[S_II]) / norm_sq(s[S_II_OLD]);
    const int64_t off = (int64_t)c * dm.ld;
    double a0 = 0.0, a1 = 0.0;
    const int64_t j = (int64_t)blockIdx.x * kVecThreads + threadIdx.x;
    if (j < dm.M) {
        const double iv = I[off + j];
        const double w = k == 0 ? iv : __dadd_rn(iv, __dmul_rn(mu, Iw[off + j]));
        Iw[off + j] = w;
        a0 = w * iv;
        a1 = w * w;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) s[S_MU] = mu;
    column_reduce3(a0, a1, 0.0, blockpart, counter, s + S_IWI, s + S_IWIW, nullptr);
}

// one CTA per column over this rank's rows of Q = Aw Iw:  red[c] = sum w q^2  (summed over the row
// shards by the caller's hook before cg_kstep_kernel)
__global__ void __launch_bounds__(1024)
cg_qq_kernel(CgDims dm, const double *__restrict__ Q, const double *__restrict__ W, const double *S,
             double *__restrict__ red) {
    __shared__ double scratch[32];
    const int c = blockIdx.x;
    if (S[(int64_t)c * S_STRIDE + S_ACTIVE] == 0.0) {
        if (threadIdx.x == 0) red[c] = 0.0;
        return;
    }
    const double *q = Q + (int64_t)c * dm.N;
    const double *w = W ? W + (int64_t)c * dm.N : nullptr;
    double acc = 0.0;
    for (int64_t l = threadIdx.x; l < dm.N; l += 1024) {
        const double v = q[l];
        acc += w ? w[l] * v * v : v * v;
    }
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) red[c] = acc;
}

// the step length  kstep = Iw.I / (|Aw Iw|^2 + alpha |Iw|^2)   (reginv.py:425, 460)
__global__ void cg_kstep_kernel(int C, const double *__restrict__ red, double *S) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double *s = S + (int64_t)c * S_STRIDE;
    if (s[S_ACTIVE] == 0.0) return;
    s[S_QQ] = red[c];
    const double den = __dadd_rn(norm_sq(red[c]), __dmul_rn(s[S_ALPHA], norm_sq(s[S_IWIW])));
    s[S_KSTEP] = s[S_IWI] / den;
}

// mw_new = Wm clamp(WmInv (mw - kstep Iw), rhomin, rhomax)   (reginv.py:427-432, 462-467)
__global__ void __launch_bounds__(kVecThreads)
cg_step_kernel(CgDims dm, const double *__restrict__ Iw, const double *__restrict__ wm,
               const double *__restrict__ wminv, double rhomin, double rhomax, double *__restrict__ mw,
               const double *S) {
    const int c = blockIdx.y;
    const double *s = S + (int64_t)c * S_STRIDE;
    if (s[S_ACTIVE] == 0.0) return;
    const double kstep = s[S_KSTEP];
    const int64_t j = (int64_t)blockIdx.x * kVecThreads + threadIdx.x;
    if (j >= dm.M) return;
    const int64_t o = (int64_t)c * dm.ld + j;
    double t = __dmul_rn(wminv[j], __dsub_rn(mw[o], __dmul_rn(kstep, Iw[o])));
    if (t < rhomin) t = rhomin;
    if (t > rhomax) t = rhomax;
    mw[o] = __dmul_rn(wm[j], t);
}

// model term and its gradient at mw:  S_MODEL_NEW = R(mw), gR = dR/dmw.
//   GI_CG_REGINV   : reginv.py:271-355 (MS gradient: denominator (mw^2 + beta)^2 -- mw, not mw - mwapr)
//   GI_CG_BOOTSTRAP: reginv.py:599-606, 620-629 (MS only, beta^2, no prior)
// Smoothness / TV apply D^T D and D^T(t / sqrt(t^2 + beta)) as 7-point stencils of the forward
// difference matrix fd3d (reginv.py:151-246), like update_body in plan.cuh.
__global__ void __launch_bounds__(kVecThreads)
cg_model_kernel(CgDims dm, const double *__restrict__ mw, const double *__restrict__ mwapr,
                const double *__restrict__ wmsq, double *__restrict__ gR, double *S, double *blockpart,
                unsigned int *counter) {
    const int c = blockIdx.y;
    double *s = S + (int64_t)c * S_STRIDE;
    if (s[S_ACTIVE] == 0.0) return;
    const double *m = mw + (int64_t)c * dm.ld;
    double um = 0.0;
    const int64_t j = (int64_t)blockIdx.x * kVecThreads + threadIdx.x;
    if (j < dm.M) {
        const double beta = dm.reg.beta;
        const double v = m[j];
        double gm = 0.0;
        if (dm.variant == GI_CG_BOOTSTRAP) {
            const double b2 = __dmul_rn(beta, beta), sq = __dmul_rn(v, v), den = __dadd_rn(sq, b2);
            um = __dmul_rn(wmsq[j], sq) / den;
            gm = __dmul_rn(__dmul_rn(2.0, wmsq[j]), __dmul_rn(v, b2)) / __dmul_rn(den, den);
        } else {
            const double dl = __dsub_rn(v, mwapr[j]);
            switch (dm.reg.reg_kind) {
                case GI_REG_DAMPING:
                    um = __dmul_rn(dl, dl);
                    gm = __dmul_rn(2.0, dl);
                    break;
                case GI_REG_MS: {
                    const double sq = __dmul_rn(dl, dl);
                    um = __dmul_rn(wmsq[j], sq) / __dadd_rn(sq, beta);
                    const double den = __dadd_rn(__dmul_rn(v, v), beta);
                    gm = __dmul_rn(__dmul_rn(__dmul_rn(2.0, beta), wmsq[j]), dl) / __dmul_rn(den, den);
                    break;
                }
                default: {
                    const int nx = dm.reg.nx, ny = dm.reg.ny, nz = dm.reg.nz;
                    const int64_t nxy = (int64_t)nx * ny;
                    const int kz = (int)(j / nxy);
                    const int rem = (int)(j - (int64_t)kz * nxy);
                    const int jy = rem / nx, ix = rem - jy * nx;
                    const bool tv = dm.reg.reg_kind == GI_REG_TV;
                    const int64_t offs[3] = {1, nx, nxy};
                    const bool has_f[3] = {ix + 1 < nx, jy + 1 < ny, kz + 1 < nz};
                    const bool has_b[3] = {ix > 0, jy > 0, kz > 0};
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        if (has_f[d]) {  // row of D owned by this cell: t = dl - dl_next
                            const double t = dl - (m[j + offs[d]] - mwapr[j + offs[d]]);
                            if (tv) {
                                const double sq = sqrt(t * t + beta);
                                um += sq;
                                gm += t / sq;
                            } else {
                                um += t * t;
                                gm += 2.0 * t;
                            }
                        }
                        if (has_b[d]) {  // row owned by the previous cell: t = dl_prev - dl
                            const double t = (m[j - offs[d]] - mwapr[j - offs[d]]) - dl;
                            if (tv) gm -= t / sqrt(t * t + beta);
                            else gm -= 2.0 * t;
                        }
                    }
                    break;
                }
            }
        }
        gR[(int64_t)c * dm.ld + j] = gm;
    }
    column_reduce3(um, 0.0, 0.0, blockpart, counter, s + S_MODEL_NEW, nullptr, nullptr);
}

// one CTA per column over this rank's rows of D = Aw mw:  r = d - dobs, R = w r (input of the
// adjoint), red[c] = sum w r^2  (summed over the row shards by the caller's hook before cg_book_kernel)
__global__ void __launch_bounds__(1024)
cg_resid_kernel(CgDims dm, const double *__restrict__ D, const double *__restrict__ dobs,
                const double *__restrict__ W, double *__restrict__ R, const double *S,
                double *__restrict__ red) {
    __shared__ double scratch[32];
    const int c = blockIdx.x;
    if (S[(int64_t)c * S_STRIDE + S_ACTIVE] == 0.0) {
        if (threadIdx.x == 0) red[c] = 0.0;
        return;
    }
    const double *d = D + (int64_t)c * dm.N;
    const double *w = W ? W + (int64_t)c * dm.N : nullptr;
    double *r = R + (int64_t)c * dm.npad;
    double acc = 0.0;
    for (int64_t l = threadIdx.x; l < dm.npad; l += 1024) {
        double rr = 0.0, wr = 0.0;
        if (l < dm.N) {
            rr = __dsub_rn(d[l], dobs[l]);
            wr = w ? w[l] * rr : rr;
        }
        r[l] = wr;
        acc += wr * rr;
    }
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) red[c] = acc;
}

// bookkeeping of the iteration from the (global) data term (reginv.py:469-488 / 692-702), one
// thread per column; n_total = observations over all row shards:
//   k < 0  : start point -- data_cur = data_new = data(mw0); the REGINV variant records entry 0
//   REGINV : record data/N, model/M at index k, then stop the column when data/N < tol
//   BOOT   : stop the column when data < tol BEFORE recording; records go to index k-1
__global__ void cg_book_kernel(CgDims dm, int k, int maxk, double tol, int64_t n_total,
                               const double *__restrict__ red, double *S, double *hist_data,
                               double *hist_model, int *nactive) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= dm.C) return;
    double *s = S + (int64_t)c * S_STRIDE;
    if (s[S_ACTIVE] == 0.0) return;
    const double data = norm_sq(red[c]);
    s[S_DATA_NEW] = data;
    const bool boot = dm.variant == GI_CG_BOOTSTRAP;
    if (k < 0) {
        s[S_DATA_CUR] = data;
        if (!boot) {
            hist_data[(int64_t)c * maxk] = data / (double)n_total;
            hist_model[(int64_t)c * maxk] = s[S_MODEL_NEW] / (double)dm.M;
        }
        return;
    }
    s[S_ITERS] = (double)(k + 1);
    if (k == 0) return;
    bool stop;
    if (boot) {
        stop = data < tol;
        if (!stop) {
            hist_data[(int64_t)c * maxk + k - 1] = data / (double)n_total;
            hist_model[(int64_t)c * maxk + k - 1] = s[S_MODEL_NEW] / (double)dm.M;
        }
    } else {
        hist_data[(int64_t)c * maxk + k] = data / (double)n_total;
        hist_model[(int64_t)c * maxk + k] = s[S_MODEL_NEW] / (double)dm.M;
        stop = data / (double)n_total < tol;
    }
    if (stop) {
        s[S_ACTIVE] = 0.0;
        atomicSub(nactive, 1);
    }
}

// model_inv = WmInv mw (reginv.py:489, 710) and the operand Wm model_inv of data_inv = A model_inv
// (reginv.py:490 multiplies the UNWEIGHTED kernel; here A = Aw Wm is applied as Aw (Wm model_inv))
__global__ void cg_finish_kernel(CgDims dm, const double *__restrict__ mw, const double *__restrict__ wm,
                                 const double *__restrict__ wminv, double *__restrict__ model,
                                 double *__restrict__ back) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dm.M) return;
    const int64_t o = (int64_t)blockIdx.y * dm.ld + j;
    const double mi = __dmul_rn(wminv[j], mw[o]);
    model[o] = mi;
    back[o] = __dmul_rn(wm[j], mi);
}

__global__ void cg_broadcast_kernel(CgDims dm, int ncols, const double *__restrict__ v, double *__restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dm.ld) return;
    out[(int64_t)blockIdx.y * dm.ld + j] = ((int)blockIdx.y < ncols && j < dm.M) ? v[j] : 0.0;
}

__global__ void cg_reset_kernel(double *S, int C, int ncols, int *nactive) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0) *nactive = ncols;
    if (c >= C) return;
    double *s = S + (int64_t)c * S_STRIDE;
    for (int i = 0; i < S_STRIDE; ++i) s[i] = 0.0;
    s[S_ACTIVE] = c < ncols ? 1.0 : 0.0;
}

}  // namespace

// =============================================================================================
struct gi_cg {
    gi_cg_config cfg;
    CgDims dm;
    gi_plan *plan;
    const double *G, *wm, *wminv, *wmsq;
    cudaStream_t s;
    int64_t vblocks;
    double *dobs, *mwapr, *W;                 // [N], [ld], [C][N] or null
    double *mw, *Iw, *I, *gR, *Gt, *D, *R;    // [C][ld] x5, [C][N], [C][npad]
    double *S, *blockpart, *hist;             // [C][16], [C][vblocks][3], [3][C][maxk]
    double *v0;                               // [ld] staging of the start model
    unsigned int *counter;
    int *nactive_dev, *nactive_host;
    int32_t hist_maxk;
    int64_t launches;
    // single-pass data term (fused.cu): D = Aw mw and Gt = Aw^T (D - dobs) in ONE pass over Aw, so an
    // iteration streams Aw twice (Aw Iw, then this) instead of three times.  -1 not tried, 0 off, 1 on
    int fused_state;
    gi_fused *fused;
    bool gt_valid;  // Gt already holds Aw^T r of the current point
    // wavelet-compressed forward of the data terms (gi_cg_set_wavelet; 0 = off)
    int wv_kind, wv_nz, wv_ny, wv_nx;
    const int64_t *wv_indptr;
    const int32_t *wv_indices;
    const double *wv_data;
    double *wv_coef;
    // row-sharded mode (gi_cg_set_shard): caller-owned reduction buffers + sum-reduce hook
    int64_t n_total;
    double *red, *gt_ext;  // [C] scalars; [C][ld] adjoint output
    double *red_own;
    gi_cg_hook hook;
    void *hook_user;
};

// plain single-column problems on kernels >= 1 GB (GI_FUSED_GEMV=1 forces, =0 disables)
static bool cg_fused_ready(gi_cg *h) {
    if (h->cfg.ncols != 1 || h->W || h->hook || h->wv_kind) return false;
    if (h->fused_state >= 0) return h->fused_state == 1;
    h->fused_state = 0;
    const char *env = getenv("GI_FUSED_GEMV");
    if (env && env[0] == '0') return false;
    const bool force = env && env[0] == '1';
    if (!force && (double)h->cfg.N * (double)h->cfg.ld * 8.0 < 1e9) return false;
    if (gi_fused_create(h->cfg.N, h->cfg.M, h->cfg.ld, h->G, h->s, &h->fused) != GI_OK) {
        h->fused = nullptr;
        return false;
    }
    h->fused_state = 1;
    return true;
}

static int cg_reduce(gi_cg *h, int what) {
    if (!h->hook) return GI_OK;
    if (h->hook(h->hook_user, what)) {
        set_error("gi_cg: the reduction hook failed");
        return GI_ERR_CUDA;
    }
    return GI_OK;
}

static void cg_free(gi_cg *h) {
    if (!h) return;
    if (h->plan) gi_plan_destroy(h->plan);
    double *bufs[] = {h->dobs, h->mwapr, h->W, h->mw, h->Iw, h->I, h->gR, h->Gt, h->D, h->R, h->S,
                      h->blockpart, h->hist, h->v0};
    for (double *b : bufs) cudaFree(b);
    cudaFree(h->counter);
    cudaFree(h->nactive_dev);
    cudaFree(h->red_own);
    cudaFree(h->wv_coef);
    gi_fused_destroy(h->fused);
    if (h->nactive_host) cudaFreeHost(h->nactive_host);
    delete h;
}

extern "C" int gi_cg_create(const gi_cg_config *cfg, const double *G_dev, const double *dobs_host,
                            const double *wm_dev, const double *wminv_dev, const double *wmsq_dev,
                            const double *mwapr_host, const double *rowweight_host, void *stream,
                            gi_cg **out) {
    GI_REQUIRE(cfg && G_dev && dobs_host && wm_dev && wminv_dev && wmsq_dev && out, "gi_cg_create: null pointer");
    GI_REQUIRE(cfg->N > 0 && cfg->M > 0 && cfg->ld >= cfg->M && cfg->ld % 4 == 0, "gi_cg_create: bad shape");
    GI_REQUIRE(cfg->ncols >= 1 && cfg->ncols <= 64, "gi_cg_create: 1..64 columns per handle");
    GI_REQUIRE(cfg->variant == GI_CG_REGINV || cfg->variant == GI_CG_BOOTSTRAP, "gi_cg_create: bad variant");
    gi_reg_params reg = cfg->reg;
    reg.constraint = GI_CONSTRAINT_MANDATORY;
    if (cfg->variant == GI_CG_BOOTSTRAP) reg.reg_kind = GI_REG_MS;  // reginv.py:599-629: MS only
    int rc = check_reg(&reg, cfg->M);
    if (rc) return rc;
    gi_cg *h = new gi_cg();
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->cfg.reg = reg;
    h->G = G_dev; h->wm = wm_dev; h->wminv = wminv_dev; h->wmsq = wmsq_dev;
    h->s = (cudaStream_t)stream;
    rc = gi_plan_create(cfg->N, cfg->M, cfg->ld, cfg->ncols, &h->plan);
    if (rc) { cg_free(h); return rc; }
    int32_t C = 1;
    int64_t npad = cfg->N;
    if (cfg->ncols > 1) gi_plan_batch_info(h->plan, &C, &npad);
    h->dm.M = cfg->M; h->dm.ld = cfg->ld; h->dm.N = cfg->N; h->dm.npad = npad; h->dm.C = C;
    h->dm.variant = cfg->variant; h->dm.reg = reg;
    h->vblocks = ceil_div(cfg->M, kVecThreads);
    const size_t vec = sizeof(double) * C * cfg->ld;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](double **p, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, bytes, h->s);
    };
    alloc(&h->dobs, sizeof(double) * cfg->N);
    alloc(&h->mwapr, sizeof(double) * cfg->ld);
    if (rowweight_host) alloc(&h->W, sizeof(double) * C * cfg->N);
    alloc(&h->mw, vec); alloc(&h->Iw, vec); alloc(&h->I, vec); alloc(&h->gR, vec); alloc(&h->Gt, vec);
    alloc(&h->D, sizeof(double) * C * cfg->N);
    alloc(&h->R, sizeof(double) * C * npad);
    alloc(&h->S, sizeof(double) * C * S_STRIDE);
    alloc(&h->blockpart, sizeof(double) * 3 * C * h->vblocks);
    alloc(&h->v0, sizeof(double) * cfg->ld);
    alloc(&h->red_own, sizeof(double) * C);
    h->red = h->red_own;
    h->n_total = cfg->N;
    h->fused_state = -1;
    if (e == cudaSuccess) e = cudaMalloc(&h->counter, sizeof(unsigned int) * C);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->counter, 0, sizeof(unsigned int) * C, h->s);
    if (e == cudaSuccess) e = cudaMalloc(&h->nactive_dev, sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost(&h->nactive_host, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->dobs, dobs_host, sizeof(double) * cfg->N, cudaMemcpyHostToDevice, h->s);
    if (e == cudaSuccess && mwapr_host)
        e = cudaMemcpyAsync(h->mwapr, mwapr_host, sizeof(double) * cfg->M, cudaMemcpyHostToDevice, h->s);
    if (e == cudaSuccess && rowweight_host)
        e = cudaMemcpyAsync(h->W, rowweight_host, sizeof(double) * cfg->ncols * cfg->N, cudaMemcpyHostToDevice, h->s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->s);
    if (e != cudaSuccess) {
        cg_free(h);
        return cuda_fail(e, "gi_cg_create", __FILE__, __LINE__);
    }
    *out = h;
    return GI_OK;
}

extern "C" int gi_cg_set_shard(gi_cg *h, int64_t n_total, double *gt_dev, double *red_dev, gi_cg_hook hook,
                               void *user) {
    GI_REQUIRE(h && n_total >= h->cfg.N && gt_dev && red_dev && hook, "gi_cg_set_shard: bad argument");
    h->n_total = n_total;
    h->gt_ext = gt_dev;
    h->red = red_dev;
    h->hook = hook;
    h->hook_user = user;
    return GI_OK;
}

extern "C" int gi_cg_set_wavelet(gi_cg *h, int32_t kind, int32_t nz, int32_t ny, int32_t nx,
                                 const int64_t *indptr, const int32_t *indices, const double *data,
                                 int64_t ncoef) {
    GI_REQUIRE(h, "gi_cg_set_wavelet: null handle");
    GI_REQUIRE(kind == 0 || kind == 1 || kind == 3, "gi_cg_set_wavelet: kind must be 0, 1 or 3");
    cudaFree(h->wv_coef);
    h->wv_coef = nullptr;
    h->wv_kind = 0;
    if (kind == 0) return GI_OK;
    GI_REQUIRE(h->cfg.ncols == 1 && !h->hook, "gi_cg_set_wavelet: single-column, unsharded handles only");
    GI_REQUIRE(indptr && indices && data && ncoef > 0, "gi_cg_set_wavelet: null CSR arrays");
    int64_t want = 0;
    if (kind == 1) {
        int rc = gi_dwt_db4_l2_1d(nullptr, h->cfg.M, nullptr, &want, nullptr);
        if (rc) return rc;
    } else {
        GI_REQUIRE((int64_t)nz * ny * nx == h->cfg.M,
                   "gi_cg_set_wavelet: the 3-D wavelet needs the full (nz, ny, nx) grid");
        int32_t shp[3];
        int rc = gi_dwt_db4_l2_3d(nullptr, nz, ny, nx, nullptr, shp, nullptr);
        if (rc) return rc;
        want = (int64_t)shp[0] * shp[1] * shp[2];
    }
    GI_REQUIRE(want == ncoef, "gi_cg_set_wavelet: CSR column count does not match the transform");
    GI_CUDA(cudaMalloc(&h->wv_coef, sizeof(double) * ncoef));
    h->wv_kind = kind; h->wv_nz = nz; h->wv_ny = ny; h->wv_nx = nx;
    h->wv_indptr = indptr; h->wv_indices = indices; h->wv_data = data;
    return GI_OK;
}

extern "C" int gi_cg_destroy(gi_cg *h) {
    cg_free(h);
    return GI_OK;
}

// D[c] = Aw X[c] for every column.  `data_term`: this forward feeds data(mw) / data_gfun(mw)
// (reginv.py:248-269), which go through the wavelet-compressed kernel when one is set -- the step
// length's Aw @ Iw (reginv.py:425) and the final A @ model_inv (:490) stay dense, as in the reference.
static int cg_forward(gi_cg *h, const double *X, bool data_term = false) {
    if (data_term && cg_fused_ready(h)) {
        // r = d - dobs without mean removal (reginv.py:256): center = 0
        h->launches += 1;
        h->gt_valid = true;
        return gi_fused_pass(h->fused, X, h->dobs, nullptr, 0, h->D, h->Gt, h->s);
    }
    h->gt_valid = false;
    if (data_term && h->wv_kind) {
        int rc = h->wv_kind == 1
                     ? gi_dwt_db4_l2_1d(X, h->cfg.M, h->wv_coef, nullptr, h->s)
                     : gi_dwt_db4_l2_3d(X, h->wv_nz, h->wv_ny, h->wv_nx, h->wv_coef, nullptr, h->s);
        if (rc) return rc;
        h->launches += (h->wv_kind == 1 ? 2 : 7);
        return gi_csr_spmv(h->wv_indptr, h->wv_indices, h->wv_data, h->cfg.N, h->wv_coef, h->D, h->s);
    }
    h->launches += 2;
    if (h->cfg.ncols == 1) return gi_gemv_fwd(h->plan, h->G, X, h->D, h->s);
    return gi_gemm_fwd(h->plan, h->G, X, h->D, h->s);
}

// Gt = Aw^T R, summed over the row shards
static int cg_adjoint(gi_cg *h) {
    if (h->gt_valid) {  // the single-pass data term of the previous step already produced it
        h->gt_valid = false;
        return GI_OK;
    }
    double *gt = h->gt_ext ? h->gt_ext : h->Gt;
    int rc;
    if (h->cfg.ncols == 1) { h->launches += 2; rc = gi_gemv_adj(h->plan, h->G, h->R, gt, h->s); }
    else { h->launches += 1; rc = gi_gemm_adj(h->plan, h->G, h->R, gt, h->s); }
    return rc ? rc : cg_reduce(h, 0);
}

static int cg_model(gi_cg *h) {
    dim3 grid((unsigned)h->vblocks, (unsigned)h->dm.C);
    cg_model_kernel<<<grid, kVecThreads, 0, h->s>>>(h->dm, h->mw, h->mwapr, h->wmsq, h->gR, h->S,
                                                    h->blockpart, h->counter);
    GI_LAUNCH_CHECK();
    h->launches += 1;
    return GI_OK;
}

static int cg_resid(gi_cg *h, int k, int maxk) {
    double *hd = h->hist + (int64_t)h->dm.C * maxk, *hm = h->hist + 2 * (int64_t)h->dm.C * maxk;
    cg_resid_kernel<<<h->dm.C, 1024, 0, h->s>>>(h->dm, h->D, h->dobs, h->W, h->R, h->S, h->red);
    GI_LAUNCH_CHECK();
    int rc = cg_reduce(h, 1);
    if (rc) return rc;
    cg_book_kernel<<<1, 64, 0, h->s>>>(h->dm, k, maxk, h->cfg.stop_tol, h->n_total, h->red, h->S, hd, hm,
                                       h->nactive_dev);
    GI_LAUNCH_CHECK();
    h->launches += 2;
    return GI_OK;
}

extern "C" int gi_cg_run(gi_cg *h, const double *mw0_host, int32_t maxk, int32_t *iters_host,
                         double *regul_host, double *data_misfit_host, double *model_misfit_host) {
    GI_REQUIRE(h && mw0_host && maxk >= 1, "gi_cg_run: bad argument");
    const CgDims &dm = h->dm;
    const int C = dm.C, ncols = h->cfg.ncols;
    cudaStream_t s = h->s;
    if (h->hist_maxk < maxk) {
        cudaFree(h->hist);
        h->hist = nullptr;
        GI_CUDA(cudaMalloc(&h->hist, sizeof(double) * 3 * C * maxk));
        h->hist_maxk = maxk;
    }
    GI_CUDA(cudaMemsetAsync(h->hist, 0, sizeof(double) * 3 * C * maxk, s));
    GI_CUDA(cudaMemcpyAsync(h->v0, mw0_host, sizeof(double) * dm.M, cudaMemcpyHostToDevice, s));
    dim3 vgrid((unsigned)h->vblocks, (unsigned)C);
    cg_reset_kernel<<<1, 64, 0, s>>>(h->S, C, ncols, h->nactive_dev);
    cg_broadcast_kernel<<<dim3((unsigned)ceil_div(dm.ld, 256), (unsigned)C), 256, 0, s>>>(dm, ncols, h->v0, h->mw);
    GI_LAUNCH_CHECK();
    GI_CUDA(cudaMemsetAsync(h->Iw, 0, sizeof(double) * C * dm.ld, s));
    h->launches += 2;
    int rc;
    // start point: R(mw0), dR(mw0), d = Aw mw0, r, data(mw0)
    if ((rc = cg_model(h))) return rc;
    if ((rc = cg_forward(h, h->mw, true))) return rc;
    if ((rc = cg_resid(h, -1, maxk))) return rc;
    for (int k = 0; k < maxk; ++k) {
        cg_alpha_kernel<<<1, 64, 0, s>>>(h->S, C, k, h->cfg.q, h->hist, maxk);
        if ((rc = cg_adjoint(h))) return rc;
        cg_grad_kernel<<<vgrid, kVecThreads, 0, s>>>(dm, h->gt_ext ? h->gt_ext : h->Gt, h->gR, h->I, h->S,
                                                     h->blockpart, h->counter);
        cg_dir_kernel<<<vgrid, kVecThreads, 0, s>>>(dm, k, h->I, h->Iw, h->S, h->blockpart, h->counter);
        GI_LAUNCH_CHECK();
        if ((rc = cg_forward(h, h->Iw))) return rc;
        cg_qq_kernel<<<C, 1024, 0, s>>>(dm, h->D, h->W, h->S, h->red);
        GI_LAUNCH_CHECK();
        if ((rc = cg_reduce(h, 1))) return rc;
        cg_kstep_kernel<<<1, 64, 0, s>>>(C, h->red, h->S);
        cg_step_kernel<<<vgrid, kVecThreads, 0, s>>>(dm, h->Iw, h->wm, h->wminv, h->cfg.rhomin, h->cfg.rhomax,
                                                    h->mw, h->S);
        GI_LAUNCH_CHECK();
        h->launches += 6;
        if ((rc = cg_model(h))) return rc;
        if ((rc = cg_forward(h, h->mw, true))) return rc;
        if ((rc = cg_resid(h, k, maxk))) return rc;
        // the reference's early stop (reginv.py:486-488 / 693-696): one 4-byte read per iteration
        GI_CUDA(cudaMemcpyAsync(h->nactive_host, h->nactive_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
        GI_CUDA(cudaStreamSynchronize(s));
        if (*h->nactive_host <= 0) break;
    }
    // results
    std::vector<double> S((size_t)C * S_STRIDE), hist((size_t)3 * C * maxk);
    GI_CUDA(cudaMemcpyAsync(S.data(), h->S, sizeof(double) * S.size(), cudaMemcpyDeviceToHost, s));
    GI_CUDA(cudaMemcpyAsync(hist.data(), h->hist, sizeof(double) * hist.size(), cudaMemcpyDeviceToHost, s));
    GI_CUDA(cudaStreamSynchronize(s));
    for (int c = 0; c < ncols; ++c) {
        if (iters_host) iters_host[c] = (int32_t)S[(size_t)c * S_STRIDE + S_ITERS];
        for (int k = 0; k < maxk; ++k) {
            if (regul_host) regul_host[(size_t)c * maxk + k] = hist[(size_t)c * maxk + k];
            if (data_misfit_host) data_misfit_host[(size_t)c * maxk + k] = hist[((size_t)C + c) * maxk + k];
            if (model_misfit_host) model_misfit_host[(size_t)c * maxk + k] = hist[((size_t)2 * C + c) * maxk + k];
        }
    }
    return GI_OK;
}

extern "C" int gi_cg_get_result(gi_cg *h, double *model_host, double *data_host, double *mw_host) {
    GI_REQUIRE(h, "gi_cg_get_result: null handle");
    const CgDims &dm = h->dm;
    const int ncols = h->cfg.ncols;
    cudaStream_t s = h->s;
    // model_inv into I, Wm model_inv into Gt (both free after the run)
    cg_finish_kernel<<<dim3((unsigned)ceil_div(dm.M, 256), (unsigned)dm.C), 256, 0, s>>>(dm, h->mw, h->wm, h->wminv,
                                                                                      h->I, h->Gt);
    GI_LAUNCH_CHECK();
    h->launches += 1;
    if (data_host) {
        int rc = cg_forward(h, h->Gt);
        if (rc) return rc;
        GI_CUDA(cudaMemcpy2DAsync(data_host, sizeof(double) * dm.N, h->D, sizeof(double) * dm.N,
                                  sizeof(double) * dm.N, ncols, cudaMemcpyDeviceToHost, s));
    }
    if (model_host)
        GI_CUDA(cudaMemcpy2DAsync(model_host, sizeof(double) * dm.M, h->I, sizeof(double) * dm.ld,
                                  sizeof(double) * dm.M, ncols, cudaMemcpyDeviceToHost, s));
    if (mw_host)
        GI_CUDA(cudaMemcpy2DAsync(mw_host, sizeof(double) * dm.M, h->mw, sizeof(double) * dm.ld,
                                  sizeof(double) * dm.M, ncols, cudaMemcpyDeviceToHost, s));
    GI_CUDA(cudaStreamSynchronize(s));
    return GI_OK;
}

extern "C" int64_t gi_cg_launch_count(const gi_cg *h) { return h ? h->launches : 0; }
