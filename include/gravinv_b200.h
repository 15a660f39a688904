/* gravinv_b200.h -- C ABI of libgravinv_b200.so (sm_100a CUDA), the drop-in boundary for the
 * GravInv3DHMC inversion hot path.
 *
 * The reference (ChuWeiEr/GravInv3DHMC) is pure Python and has no FFI registry; the seams this
 * library replaces are its two native leaf calls and the numpy/scipy arithmetic of the sampler
 * (reference paths are relative to the reference root):
 *
 *   gi_prism_gz_assemble     <- gravmag/_prism.pyx:263-290 gz()  called per prism by
 *                               gravmag/prism.py:291-316 _gz (one kernel2d column per prism)
 *   gi_tess_gz_assemble      <- gravmag/_tesseroid_numba.py:25-72 engine(kernelz) called per
 *                               tesseroid by gravmag/tesseroid.py:189-232 _forward_model
 *   gi_colsumsq, gi_weights_from_sumsq, gi_scale_columns
 *                            <- inversion/potential.py:232-264 sensitivityWeighting
 *   gi_gemv_fwd / gi_residual / gi_gemv_adj / gi_update
 *                            <- inversion/potential.py:688-717 data_all, :719-810 model_*_all,
 *                               :812-845 misfit_and_grad and the body of
 *                               inversion/hmc.py:85-177 _leapfrog
 *   gi_hmc_*                 <- inversion/hmc.py:85-177 _leapfrog, :252-343 sample (one proposal)
 *   gi_gemm_* / gi_*_batched / gi_hmcb_*
 *                            <- the same lines for a batch of chains (the reference runs one process
 *                               per chain: example/uniformgrid/run_main.sh:18)
 *   gi_cg_*                  <- inversion/reginv.py:357-491 ConjugateGradient.CG, :631-748 BootStrap
 *   gi_stats_*               <- inversion/hmc.py:241-249 (model.dat sink) + example/uniformgrid/
 *                               plot_uniform.py:103-114 (posterior mean / std and their forward data)
 *   gi_dwt_db4_* / gi_csr_spmv
 *                            <- gravmag/compressor1D.py:45-60, compressor3D.py:47-68 modelcompressor
 *
 * Conventions
 *   - every function returns GI_OK (0) or a negative GI_ERR_* code and never throws;
 *     gi_last_error() returns a thread-local message for the last failure.
 *   - pointers named *_dev are CUDA device pointers owned by the caller (e.g. torch
 *     Tensor.data_ptr()); pointers named *_host are host pointers (pinned memory makes the
 *     copies asynchronous).  No torch types cross this boundary.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - all arithmetic is IEEE binary64; matrices are row-major with leading dimension `ld`
 *     (in doubles, a multiple of 4; columns [M, ld) must be zero).  Every device "M-vector"
 *     (x, mw, p, grad, low, high, wm ...) is allocated with ld entries, entries [M, ld) zero.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     GI_ERR_CUDA.
 */
#ifndef GRAVINV_B200_H
#define GRAVINV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GI_ABI_VERSION 1

#define GI_OK 0
#define GI_ERR_INVALID (-1)  /* bad argument (ValueError in the Python layer) */
#define GI_ERR_CUDA (-2)     /* CUDA runtime error / no device */
#define GI_ERR_OVERFLOW (-3) /* tesseroid subdivision stack overflow (OverflowError, _tesseroid_numba.py:53-54) */
#define GI_ERR_NOMEM (-4)
#define GI_ERR_BUSY (-5)     /* gi_hmcb_stream_feed: the chain's proposal queue is full */
#define GI_STREAM_QUEUE_DEPTH 4 /* proposals a chain may have queued behind the running one */

/* regularisers, inversion/potential.py:831-836 */
#define GI_REG_DAMPING 0
#define GI_REG_MS 1
#define GI_REG_SMOOTHNESS 2
#define GI_REG_TV 3
/* boundary constraint, inversion/hmc.py:271-278 */
#define GI_CONSTRAINT_MANDATORY 0
#define GI_CONSTRAINT_LOGARITHMIC 1

int gi_abi_version(void);
const char *gi_last_error(void);
/* sm count, compute capability, L2 bytes of the current device */
int gi_device_info(int *sm_count, int *cc_major, int *cc_minor, int64_t *l2_bytes);

/* ---- sensitivity-matrix assembly -------------------------------------------------------- */
/* G[l*ld + c] = scale * sum over the 8 corners of prism c at observation l (gz, closed form).
 * bounds_dev is [M][6] = x1,x2,y1,y2,z1,z2 of the ACTIVE prisms in mesh order.
 * Columns [M, ld) are written as zeros.  Rows [row0, row0+nrows) of the observation arrays
 * are assembled into rows [0, nrows) of G (row sharding across GPUs). */
int gi_prism_gz_assemble(const double *xp_dev, const double *yp_dev, const double *zp_dev,
                         int64_t nrows, const double *bounds_dev, int64_t M, double scale,
                         double *G_dev, int64_t ld, void *stream);

/* Same result for a STRUCTURED prism mesh (mesher/mesh.py PrismMesh / PrismMeshSegment) whose cells
 * share their edges bit for bit: xn[nx+1], yn[ny+1], zn[nz+1] are the node coordinates per axis
 * (cell (k,j,i) spans xn[i]..xn[i+1] etc.; flat cell index (k*ny + j)*nx + i as in mesh.py:229-236).
 * The corner term is evaluated once per node and shared by the up to 8 cells around it, in the
 * reference's summation order -> bit-identical to gi_prism_gz_assemble, ~5x fewer evaluations.
 * colmap_dev (optional, int32 [nx*ny*nz]) maps a cell to its column (-1 = carved out); without it
 * M must equal nx*ny*nz.  Columns [M, ld) are zeroed. */
int gi_prism_gz_assemble_grid(const double *xp_dev, const double *yp_dev, const double *zp_dev,
                              int64_t nrows, const double *xn_dev, const double *yn_dev,
                              const double *zn_dev, int32_t nx, int32_t ny, int32_t nz,
                              const int32_t *colmap_dev, int64_t M, double scale, double *G_dev,
                              int64_t ld, void *stream);

/* Tesseroid gz, 2x2x2 Gauss-Legendre with the reference's LIFO adaptive subdivision.
 * Inputs are the converted coordinates of gravmag/tesseroid.py:109-123.
 * status_dev is int32[2]: [0] accumulates the reference's error_code (-1 per refused split),
 * [1] is set to 1 if any pair overflowed the 100-entry stack (that entry is NaN).
 * G = (raw * scale1) * scale2  as in gravmag/tesseroid.py:430. */
int gi_tess_gz_assemble(const double *lon_dev, const double *sinlat_dev, const double *coslat_dev,
                        const double *radius_dev, int64_t nrows, const double *bounds_dev,
                        int64_t M, double ratio, double scale1, double scale2, double *G_dev,
                        int64_t ld, int32_t *status_dev, void *stream);

/* Subdivision bookkeeping of the same engine: leaves[l*ld + c] = number of leaf cells evaluated for
 * pair (l, c), as a double (-1 where the stack would overflow).  Lets callers and tests compare the
 * adaptive-subdivision DECISIONS (_tesseroid_numba.py:135-157) exactly, apart from the values. */
int gi_tess_gz_leafcount(const double *lon_dev, const double *sinlat_dev, const double *coslat_dev,
                         const double *radius_dev, int64_t nrows, const double *bounds_dev, int64_t M,
                         double ratio, double *leaves_dev, int64_t ld, int32_t *status_dev,
                         void *stream);

/* ---- the other prism fields (gravmag/_prism.pyx) --------------------------- */
/* G[l*ld + c] = scale * sum over the 8 corners of prism c at observation l of the corner kernel of
 * `field` (_prism.pyx:36-70), same layout / row sharding as gi_prism_gz_assemble.  Replaces the
 * per-prism Cython calls potential :484, gx :206, gy :235, gz :265, gxx :294, gxy :324, gxz :358,
 * gyy :392, gyz :421, gzz :455 (the three mixed components with the displaced radius of
 * :345-350 / :380-385 / :442-447), tf :72 (vec3 = the unit vector f of the regional field: the
 * kernel is f.(V f)) and the rows of V applied to vec3 (bx :116, by :146, bz :176). */
#define GI_FIELD_POTENTIAL 0
#define GI_FIELD_GX 1
#define GI_FIELD_GY 2
#define GI_FIELD_GZ 3
#define GI_FIELD_GXX 4
#define GI_FIELD_GXY 5
#define GI_FIELD_GXZ 6
#define GI_FIELD_GYY 7
#define GI_FIELD_GYZ 8
#define GI_FIELD_GZZ 9
#define GI_FIELD_TF 10
#define GI_FIELD_VX 11
#define GI_FIELD_VY 12
#define GI_FIELD_VZ 13
int gi_prism_field_assemble(int32_t field, const double *xp_dev, const double *yp_dev,
                            const double *zp_dev, int64_t nrows, const double *bounds_dev, int64_t M,
                            double scale, const double *vec3_host, double *G_dev, int64_t ld,
                            void *stream);
/* Tesseroid fields other than gz (GI_FIELD_POTENTIAL .. GI_FIELD_GZZ): the GLQ kernels of
 * gravmag/_tesseroid_numba.py:160-328 through the same adaptive engine (:25-157) as
 * gi_tess_gz_assemble; `ratio` is the distance-size ratio of gravmag/tesseroid.py:76-78 (1 / 1.6 / 8),
 * G = (raw * scale1) * scale2 as in tesseroid.py:375-507 (pass scale2 = 1 for a single factor).
 * Also serves gravmag/tesseroidforward.py (forward result = G @ density). */
int gi_tess_field_assemble(int32_t field, const double *lon_dev, const double *sinlat_dev,
                           const double *coslat_dev, const double *radius_dev, int64_t nrows,
                           const double *bounds_dev, int64_t M, double ratio, double scale1,
                           double scale2, double *G_dev, int64_t ld, int32_t *status_dev, void *stream);

/* ---- sensitivity weighting (potential.py:232-264) ---------------------------------------- */
/* out[c] (+)= sum_l G[l][c]^2, rows summed sequentially in row order. */
int gi_colsumsq(const double *G_dev, int64_t nrows, int64_t M, int64_t ld, double *out_dev,
                int accumulate, void *stream);
/* wm = sumsq^weightfactor, wminv = 1/wm, wmsq = wm*wm */
int gi_weights_from_sumsq(const double *sumsq_dev, int64_t M, double weightfactor, double *wm_dev,
                          double *wminv_dev, double *wmsq_dev, void *stream);
/* G[l][c] *= colscale[c] in place (Aw = A @ WmInv) */
int gi_scale_columns(double *G_dev, int64_t nrows, int64_t M, int64_t ld,
                     const double *colscale_dev, void *stream);

/* ---- leapfrog building blocks (used directly by the row-sharded multi-GPU driver) -------- */
typedef struct gi_plan gi_plan; /* tiling + workspace for one (nrows, M, ld, nchains) problem */

int gi_plan_create(int64_t nrows, int64_t M, int64_t ld, int32_t nchains, gi_plan **out);
int gi_plan_destroy(gi_plan *plan);
/* number of kernels the plan launches for one fwd / adj pass (for launch accounting) */
int gi_plan_info(const gi_plan *plan, int64_t *fwd_tiles, int64_t *adj_tiles,
                 int64_t *workspace_bytes);

/* d = G x  (x has ld entries, zero padded).  Deterministic two-stage reduction. */
int gi_gemv_fwd(gi_plan *plan, const double *G_dev, const double *x_dev, double *d_dev,
                void *stream);
/* sums_dev[0] = sum_l (d[l] + fix[l])   (fix_dev may be NULL) */
int gi_data_sum(gi_plan *plan, const double *d_dev, const double *fix_dev, double *sums_dev,
                void *stream);
/* r[l] = (d[l] + fix[l] - mean) - dobs_c[l] with mean = sums_dev[0] / n_total (sums_dev[0] is
 * the sum over ALL ranks' rows, n_total the global observation count);  sums_dev[1] = sum r^2. */
/* n_total <= 0: no mean removal (r = d + fix - dobs_c with dobs_c = dobs) */
int gi_residual(gi_plan *plan, const double *d_dev, const double *fix_dev,
                const double *dobs_c_dev, int64_t n_total, double *r_dev, double *sums_dev,
                void *stream);
/* g = G^T r (no factor 2; the caller's update applies it).  Deterministic. */
int gi_gemv_adj(gi_plan *plan, const double *G_dev, const double *r_dev, double *g_dev,
                void *stream);

typedef struct {
    int32_t reg_kind;   /* GI_REG_* */
    int32_t constraint; /* GI_CONSTRAINT_* */
    int32_t nz, ny, nx; /* grid shape (Smoothness / TV need nz*ny*nx == M) */
    int32_t reserved;
    double alpha;      /* RegulFactor */
    double beta;       /* MS / TV focusing parameter */
    double log_factor; /* logarithmic constraint */
} gi_reg_params;

/* One fused M-vector pass (hmc.py:114-152 + potential.py:719-810):
 *   grad = 2*gdata + alpha*dR/dmw(mw - mwapr);  Um = R(mw - mwapr)
 *   p   -= pcoef * grad
 *   if advance: x += dt*p; clamp to [low, high] and flip p (mandatory) ; mw_out = mw(x)
 * gdata_dev holds G^T r (already summed over ranks).  sums_dev[2] = Um, sums_dev[3] = 0.5*p.p
 * (after the update).  grad_out_dev may be NULL. x_in/x_out and mw_in/mw_out may not alias. */
int gi_update(gi_plan *plan, const gi_reg_params *reg, const double *gdata_dev,
              const double *x_in_dev, const double *mw_in_dev, const double *mwapr_dev,
              const double *wmsq_dev, const double *low_dev, const double *high_dev, double *p_dev,
              double *x_out_dev, double *mw_out_dev, double *grad_out_dev, double pcoef, double dt,
              int advance, double *sums_dev, void *stream);

/* ---- single-pass gradient evaluation (one chain) ------------------------------------------ */
/* d = G x and g = G^T r, r = (d + fix - mean(d + fix)) - dobs_c (potential.py:698-708), with ONE pass
 * over G instead of two: a persistent kernel (one CTA per SM, cooperative launch) keeps every row's
 * column strip in shared memory (TMA bulk copies into a 4-slot ring) between its forward dot product
 * and its adjoint update, using G^T r = G^T e - mean * (G^T 1).  Deterministic; DRAM traffic 8 N M
 * bytes per evaluation.  gi_fused_create fails with GI_ERR_INVALID when the strip of one SM
 * (ld / #SMs columns x 4 slots) does not fit its shared memory -- callers then use gi_gemv_fwd /
 * gi_gemv_adj.  gi_hmc and the single-column gi_cg use it automatically for kernels >= 1 GB (env
 * GI_FUSED_GEMV=0/1 overrides). */
typedef struct gi_fused gi_fused;
int gi_fused_create(int64_t nrows, int64_t M, int64_t ld, const double *G_dev, void *stream,
                    gi_fused **out);
int gi_fused_destroy(gi_fused *f);
/* center = 1: r = e - mean(e) with e = d + fix - dobs_c (potential.py:699-706); center = 0: r = e
 * (the residual of inversion/reginv.py:256, 267: no mean removal) */
int gi_fused_pass(gi_fused *f, const double *x_dev, const double *dobs_c_dev, const double *fix_dev,
                  int32_t center, double *d_dev, double *g_dev, void *stream);

/* ---- single-GPU device-resident sampler --------------------------------------------------- */
typedef struct gi_hmc gi_hmc;

typedef struct {
    int64_t N, M, ld;
    int32_t fixed; /* potential.py:699-703: add grav_fix to the forward data */
    int32_t nocenter; /* 0: r = (d - mean d) - (dobs - mean dobs) (GravMagModule, potential.py:706);
                         1: r = d - dobs (JointModule.data_all, potential.py:1665-1680) */
    gi_reg_params reg;
} gi_hmc_config;

typedef struct {
    int32_t accept;
    int32_t L;
    double U, U_data, U_model; /* of the state the chain is in AFTER the proposal */
    double Hcur, Hnew;
    double Unew, Unew_data, Unew_model; /* of the proposed end point */
} gi_hmc_result;

/* G_dev (N x ld, weighted kernel Aw) stays owned by the caller and must outlive the handle.
 * Host vectors (length M: low, high, mwapr, wmsq(optional); length N: dobs, grav_fix(optional))
 * are copied to the device. */
int gi_hmc_create(const gi_hmc_config *cfg, const double *G_dev, const double *dobs_host,
                  const double *gravfix_host, const double *low_host, const double *high_host,
                  const double *mwapr_host, const double *wmsq_host, void *stream, gi_hmc **out);
int gi_hmc_destroy(gi_hmc *h);
int gi_hmc_set_reg(gi_hmc *h, const gi_reg_params *reg);
/* set the current position x (hmc.py:271-276) and evaluate U, grad there */
int gi_hmc_set_state(gi_hmc *h, const double *x_host);
/* copy out the current position, forward data d = Aw mw (without grav_fix) and mw (any may be NULL) */
int gi_hmc_get_state(gi_hmc *h, double *x_host, double *d_host, double *mw_host);
/* current misfit and gradient (potential.py:812-845 return values) */
int gi_hmc_get_misfit(gi_hmc *h, double *U, double *U_data, double *U_model, double *grad_host);

/* One HMC proposal (hmc.py:85-177) with injected draws: p0_host = randn(M)*Sigma, u = rand().
 * trace_x_host ((L+1) x M) / trace_U_host (L+1) optionally receive the position and potential
 * after every gradient evaluation (index 0 = the start point). */
int gi_hmc_propose(gi_hmc *h, const double *p0_host, int32_t L, double dt, double u,
                   gi_hmc_result *result, double *trace_x_host, double *trace_U_host);
/* Same with device-generated draws (Philox4x32-10 + Box-Muller; NOT bit-compatible with numpy's
 * MT19937 stream): p0 = N(0,1)*sigma from (seed, counter). */
int gi_hmc_propose_philox(gi_hmc *h, uint64_t seed, uint64_t counter, double sigma, int32_t L,
                          double dt, gi_hmc_result *result);
/* Benchmark / warm-up helper: run `nsteps` leapfrog steps (hmc.py:117-152) from the current state
 * without a Metropolis test; the momentum starts at p0_dev (device, length ld) or zero. */
int gi_hmc_leapfrog_steps(gi_hmc *h, const double *p0_dev, int32_t nsteps, double dt);
/* Wavelet-compressed forward (potential.py:693-696): d = Awcp @ DWT(mw) replaces Aw @ mw in every
 * gradient evaluation of this handle; the gradient keeps the dense Aw^T (potential.py:708).
 * kind 1 = compressor1D (nz,ny,nx ignored), 3 = compressor3D on the (nz,ny,nx) grid, 0 = off.
 * The CSR arrays (N rows, ncoef columns) stay owned by the caller and must outlive the handle. */
int gi_hmc_set_wavelet(gi_hmc *h, int32_t kind, int32_t nz, int32_t ny, int32_t nx,
                       const int64_t *indptr_dev, const int32_t *indices_dev, const double *data_dev,
                       int64_t ncoef);
/* kernels launched by this handle since creation */
int64_t gi_hmc_launch_count(const gi_hmc *h);
/* which gradient evaluation the handle uses: 0 = the two GEMV passes, 1 = the single-pass evaluation
 * (rows of >= 4 MB in kernels of >= 1 GB whose column strip fits one SM), 2 = wavelet-compressed forward */
int32_t gi_hmc_eval_path(gi_hmc *h);
void *gi_hmc_stream(const gi_hmc *h);

/* ---- C independent chains batched as columns (FP64 tensor-core contractions) ----------------- */
/* The reference runs one OS process per chain (example/run_main.sh:18 `mpiexec -n 2`), each with its
 * own copy of Aw; here up to 64 chains share one pass over Aw: D = Aw X, Gt = Aw^T R as DMMA
 * (mma.sync m8n8k4 f64) contractions.  Per-chain vectors are chain-major: X, P, grad [Cp][ld];
 * D [Cp][nrows]; R [Cp][npad]; sums [Cp][8], where Cp >= nchains and npad >= nrows are the padded
 * sizes reported by gi_plan_batch_info (padding chains/rows must be zero-filled by the caller).
 * A plan created with nchains > 1 serves these entry points (ld must be a multiple of 32). */
int gi_plan_batch_info(const gi_plan *plan, int32_t *padded_chains, int64_t *padded_rows);
/* D[c][l] = sum_k G[l][k] X[c][k] */
int gi_gemm_fwd(gi_plan *plan, const double *G_dev, const double *X_dev, double *D_dev, void *stream);
/* sums[c][0] = sum_l (D[c][l] + fix[l]) */
int gi_data_sum_batched(gi_plan *plan, const double *D_dev, const double *fix_dev, double *sums_dev,
                        void *stream);
/* R[c][l] = (D[c][l] + fix[l] - sums[c][0]/n_total) - dobs_c[l];  sums[c][1] = sum_l R[c][l]^2 */
int gi_residual_batched(gi_plan *plan, const double *D_dev, const double *fix_dev,
                        const double *dobs_c_dev, int64_t n_total, double *R_dev, double *sums_dev,
                        void *stream);
/* Gt[c][k] = sum_l G[l][k] R[c][l] */
int gi_gemm_adj(gi_plan *plan, const double *G_dev, const double *R_dev, double *Gt_dev, void *stream);
/* gi_update for every chain of the batch.  The gradient is grad_in_dev ([Cp][ld], complete) when
 * given, else 2*gdata + alpha*dR.  Chain c at trajectory step `step` (L_dev[c] = its trajectory
 * length): step < L: p -= dt*grad, x += dt*p, clamp;  step == L: p -= dt/2*grad, grad_out written;
 * step > L: frozen; step == 0: opening half step (mode 3 below); L == 0: the chain is inactive
 * (frozen throughout).  With L_dev == NULL every chain takes uniform_mode (0 full step, 1 final half
 * step, 2 frozen, 3 opening half step + advance, which also parks K before the update in sums[c][5]).
 * low/high/mwapr/wmsq are shared [ld] vectors.  sums[c][2] = Um, [3] = K after, [4] = K before. */
int gi_update_batched(gi_plan *plan, const gi_reg_params *reg, const double *grad_in_dev,
                      const double *gdata_dev, const double *x_in_dev, const double *mw_in_dev,
                      const double *mwapr_dev, const double *wmsq_dev, const double *low_dev,
                      const double *high_dev, double *p_dev, double *x_out_dev, double *mw_out_dev,
                      double *grad_out_dev, double dt, const int32_t *L_dev, int32_t step,
                      int32_t uniform_mode, double *sums_dev, void *stream);

/* Device-resident sampler for a batch of 2..64 chains (same roles as gi_hmc_*; host arrays are
 * chain-major and un-padded: x [nchains][M], d [nchains][N], p0 [nchains][M]). */
typedef struct gi_hmcb gi_hmcb;
int gi_hmcb_create(const gi_hmc_config *cfg, int32_t nchains, const double *G_dev,
                   const double *dobs_host, const double *gravfix_host, const double *low_host,
                   const double *high_host, const double *mwapr_host, const double *wmsq_host,
                   void *stream, gi_hmcb **out);
int gi_hmcb_destroy(gi_hmcb *h);
int gi_hmcb_set_reg(gi_hmcb *h, const gi_reg_params *reg);
int gi_hmcb_set_state(gi_hmcb *h, const double *x_host);
int gi_hmcb_get_state(gi_hmcb *h, double *x_host, double *d_host, double *mw_host);
int gi_hmcb_get_misfit(gi_hmcb *h, double *U, double *U_data, double *U_model, double *grad_host);
/* one proposal per chain with injected draws; chain c runs L_host[c] leapfrog steps (chains whose
 * trajectory is shorter than the longest one idle; L = 0 sits the proposal out).  results[nchains]; optional traces
 * trace_x_host [(Lmax+1)][nchains][M], trace_U_host [(Lmax+1)][nchains] (entries past L_c repeat). */
int gi_hmcb_propose(gi_hmcb *h, const double *p0_host, const int32_t *L_host, double dt,
                    const double *u_host, gi_hmc_result *results, double *trace_x_host,
                    double *trace_U_host);
/* device draws: chain c uses the Philox key seed + c (like the reference's seed + myrank) */
int gi_hmcb_propose_philox(gi_hmcb *h, uint64_t seed, uint64_t counter, double sigma,
                           const int32_t *L_host, double dt, gi_hmc_result *results);
/* Row-sharded batches (G partitioned by observation rows over GPUs): the handle is created
 * over this rank's rows (cfg.N = local rows) and told the global row count; at the two exchange
 * points of every gradient evaluation it calls `hook` on its stream's behalf:
 *   what 0: sum-reduce red_dev[0 : Cp]      (sum of the forward data per chain -> global mean)
 *   what 1: sum-reduce gradient piece `piece` (gdata_dev + piece*Cp*(ld/npieces), Cp*(ld/npieces)
 *           doubles); async != 0: the reduction may complete later, what 3 waits for all of them --
 *           the adjoint contraction of the next piece overlaps the reduction of the previous one
 *   what 2: sum-reduce red_dev[Cp : 2*Cp]   (sum of squared residuals per chain)
 *   what 3: make the handle's stream wait for the outstanding asynchronous reductions
 * All reductions are in place over every rank and must be ordered after the work already queued on
 * the handle's stream (torch.distributed with NCCL does exactly this).  gdata_dev ([npieces][Cp]
 * [ld/npieces], ld/npieces a multiple of 256) and red_dev ([2*Cp]) are caller-owned so the caller's
 * collective library can address them.  dobs_c_host = this rank's rows of dobs - mean(all dobs).
 * Every rank must feed identical draws; the replicated chain state then stays bitwise identical. */
typedef int (*gi_shard_hook)(void *user, int32_t what, int32_t piece, int32_t async);
int gi_hmcb_set_shard(gi_hmcb *h, int64_t n_total, const double *dobs_c_host, double *gdata_dev,
                      int32_t npieces, double *red_dev, gi_shard_hook hook, void *user);

/* Peer-memory exchange (the default of row-sharded batches on one NVLink / NVSwitch node): instead of
 * calling back into a collective library at the exchange points, the ranks map each other's memory
 * (CUDA IPC) and the kernels themselves move the data:
 *   - the adjoint contraction's epilogue stores every 256-column tile into the staging block of the
 *     rank that OWNS those columns (reduce-scatter fused into the contraction, transfer spread over
 *     the whole pass); the owner's fused update sums the `world` staged partials in rank order and
 *     updates only its own column slice (the replicated update of the hook path shrinks by `world`);
 *   - the updated slice is pushed into every peer's position buffer by the copy engines (all-gather),
 *     nearest reader first, and the next forward contraction starts at once on the own slice: its
 *     CTAs poll a per-source-rank epoch flag before touching another rank's columns;
 *   - per-chain scalars (sum d, sum r^2, Um, K) go through a slot table written by one small kernel
 *     that adds the slots in rank order, so every rank holds identical bits (identical Metropolis
 *     decisions) without any collective call.
 * Protocol: every rank creates a gi_peer with the SAME byte count (>= gi_hmcb_peer_bytes), exchanges
 * the 64-byte handles of gi_peer_export by any host-side means, calls gi_peer_connect with all of them
 * ([world][64], own entry ignored), then gi_hmcb_set_peer; a host barrier must separate set_peer from
 * the first sampler call.  All ranks must issue the same sequence of sampler calls with identical
 * draws.  After set_peer the gradient the handle keeps (gi_hmcb_get_misfit's grad_host) is valid on
 * the owned columns only (gi_hmcb_owned_columns). */
typedef struct gi_peer gi_peer;
int gi_peer_create(int32_t rank, int32_t world, int64_t bytes, gi_peer **out);
int gi_peer_export(gi_peer *p, void *handle64);
int gi_peer_connect(gi_peer *p, const void *handles);
int gi_peer_destroy(gi_peer *p);
/* bytes this rank has stored into its peers' memory so far (NVLink traffic accounting) */
int64_t gi_peer_bytes_sent(const gi_peer *p);
/* in-place sum over the ranks of n doubles (8..512, a multiple of 8) on the device, identical bits on
 * every rank: the scalar exchange above, exposed for set-up reductions and tests */
int gi_peer_allreduce_small(gi_peer *p, double *vec_dev, int32_t n, void *stream);
/* host-only helper: the column slices, col[0..world]; slice q = [col[q], col[q+1]) is a whole number of
 * the adjoint kernel's 256-column strips (the last one ends at ld) and may be empty */
int gi_peer_columns(int64_t ld, int32_t world, int64_t *col);
int64_t gi_hmcb_peer_bytes(const gi_hmcb *h, int32_t world);
int gi_hmcb_set_peer(gi_hmcb *h, gi_peer *peer, int64_t n_total, const double *dobs_c_host);
int gi_hmcb_owned_columns(const gi_hmcb *h, int64_t *lo, int64_t *hi);

/* Streaming mode: every chain runs its own sequence of proposals back to back.  All chains share
 * each gradient evaluation (one "batch step"), and a chain that ends a trajectory -- Metropolis test,
 * commit -- opens its next one inside the same step, so no chain idles while others finish longer
 * trajectories (hmc.py:295-300 per chain, with each chain's own L sequence).  The host feeds the draws
 * of up to GI_STREAM_QUEUE_DEPTH proposals per chain ahead of time (L, u = rand(), p0 = randn(M)*Sigma, in the
 * reference's RNG order) and collects one record per finished proposal.
 *   begin    -- chains keep their current state (gi_hmcb_set_state); queues are emptied
 *   feed     -- enqueue one proposal for `chain`; GI_ERR_BUSY if the chain's queue is full
 *   runway   -- batch steps that can run before some started/fed chain runs out of queued work
 *   advance  -- run up to nsteps batch steps (fewer if every chain runs dry or the record buffer
 *               fills); returns the records of the proposals that finished, in completion order;
 *               x_host (optional, [max_records][M], ideally pinned) receives the chain's current
 *               position after each finished proposal (the accepted sample when accept == 1). */
typedef struct {
    int32_t chain, accept, L, reserved;
    int64_t seq;                /* index of the proposal within its chain (0, 1, ...) */
    double U, U_data, U_model;  /* of the state the chain is in AFTER the proposal */
    double Hcur, Hnew;
} gi_stream_record;
int gi_hmcb_stream_begin(gi_hmcb *h, double dt);
int gi_hmcb_stream_feed(gi_hmcb *h, int32_t chain, int32_t L, double u, const double *p0_host);
/* same as feed with the momentum draw already on the device (e.g. received by a broadcast) */
int gi_hmcb_stream_feed_dev(gi_hmcb *h, int32_t chain, int32_t L, double u, const double *p0_dev);
int gi_hmcb_stream_runway(gi_hmcb *h, int32_t *steps);
/* the host has fed the last proposal of `chain`: runway no longer stops where this chain runs dry */
int gi_hmcb_stream_close_chain(gi_hmcb *h, int32_t chain);
int gi_hmcb_stream_advance(gi_hmcb *h, int32_t nsteps, gi_stream_record *records, int32_t max_records,
                           int32_t *nrecords, int32_t *steps_done, double *x_host);
/* advance in two halves: _begin queues the kernels of up to nsteps batch steps and returns at once (the
 * schedule does not depend on the Metropolis outcomes), _end waits for them and hands out the records.
 * Between the two the host may feed the queue slots the scheduled steps have freed
 * (gi_hmcb_stream_queue_space), so that drawing and staging the next proposals overlaps the device work
 * instead of leaving the device idle between calls. */
int gi_hmcb_stream_advance_begin(gi_hmcb *h, int32_t nsteps, int32_t max_records, double *x_host);
int gi_hmcb_stream_advance_end(gi_hmcb *h, gi_stream_record *records, int32_t max_records,
                               int32_t *nrecords, int32_t *steps_done);
int gi_hmcb_stream_queue_space(gi_hmcb *h, int32_t chain, int32_t *space);
/* drop the queued, not yet started proposals of `chain` (it has reached its sample count); the one in
 * flight still finishes and is recorded */
int gi_hmcb_stream_cancel(gi_hmcb *h, int32_t chain, int32_t *dropped);
/* benchmark helper: nsteps leapfrog steps of every chain, no Metropolis; p0_dev is [Cp][ld] or NULL */
int gi_hmcb_leapfrog_steps(gi_hmcb *h, const double *p0_dev, int32_t nsteps, double dt);
int64_t gi_hmcb_launch_count(const gi_hmcb *h);
int32_t gi_hmcb_padded_chains(const gi_hmcb *h);

/* ---- host helper: the reference's random stream ---------------------------------------------- */
/* out_host[i] = randn()_i * scale for the next n normals of a numpy legacy RandomState (MT19937 +
 * polar method with cached second deviate), bit for bit what `rs.randn(n) * scale` returns, continuing
 * from -- and writing back -- the generator state `rs.get_state()` exposes (key[624], pos, has_gauss,
 * cached_gaussian).  The samplers draw momentum in the reference's RNG order (hmc.py:95); this is
 * ~2x numpy's pace and releases the GIL for the whole vector.  HOST pointers only. */
int gi_legacy_randn_scaled(uint32_t *key624, int32_t *pos, int32_t *has_gauss, double *cached_gauss,
                           int64_t n, double scale, double *out_host);

/* release store / acquire load of one control word of the draw ring that the ranks of a node share
 * through mapped host memory (payload before flag on every host architecture, not only x86).  HOST
 * pointers only. */
void gi_ring_store_release(int64_t *word, int64_t value);
int64_t gi_ring_load_acquire(const int64_t *word);

/* ---- wavelet-compressed forward (compressor1D/3D.py) -------------------------------------- */
/* level-2 db4 periodization DWT of a length-n vector packed like pywt.coeffs_to_array:
 * out = [cA2 | cD2 | cD1], ncoef = len(out) returned through *ncoef (may be called with
 * out_dev == NULL to query the size). */
int gi_dwt_db4_l2_1d(const double *x_dev, int64_t n, double *out_dev, int64_t *ncoef, void *stream);
/* 3-D variant on a (nz, ny, nx) C-ordered volume, Mallat packing of pywt.coeffs_to_array. */
int gi_dwt_db4_l2_3d(const double *x_dev, int32_t nz, int32_t ny, int32_t nx, double *out_dev,
                     int32_t out_shape[3], void *stream);
/* batched variants: `batch` vectors/volumes (<= 65535) with strides x_bs / out_bs in doubles
 * (used to transform many kernel rows at once, compressor*.kernelcompressor) */
int gi_dwt_db4_l2_1d_batch(const double *x_dev, int64_t batch, int64_t x_bs, int64_t n,
                           double *out_dev, int64_t out_bs, int64_t *ncoef, void *stream);
int gi_dwt_db4_l2_3d_batch(const double *x_dev, int64_t batch, int64_t x_bs, int32_t nz, int32_t ny,
                           int32_t nx, double *out_dev, int64_t out_bs, int32_t out_shape[3],
                           void *stream);
/* dense coefficient rows -> CSR (compressor*.kernelcompressor: entries with |c| < thr are zeroed,
 * scipy's csr_matrix drops the zeros).  counts[row] = kept entries; after an exclusive scan into
 * indptr (row offsets relative to this block of rows), gi_csr_fill writes indices/data in ascending
 * column order. */
int gi_csr_count(const double *dense_dev, int64_t nrows, int64_t ncols, int64_t row_stride, double thr,
                 int64_t *counts_dev, void *stream);
int gi_csr_fill(const double *dense_dev, int64_t nrows, int64_t ncols, int64_t row_stride, double thr,
                const int64_t *indptr_dev, int32_t *indices_dev, double *data_dev, void *stream);
/* y = A x for a CSR matrix (int64 indptr[nrows+1], int32 indices, f64 data) */
int gi_csr_spmv(const int64_t *indptr_dev, const int32_t *indices_dev, const double *data_dev,
                int64_t nrows, const double *x_dev, double *y_dev, void *stream);

/* ---- regularised conjugate gradient + bootstrap (inversion/reginv.py) ----- */
/* One handle runs ncols independent CG problems in lockstep over one weighted kernel Aw:
 *   GI_CG_REGINV    <- ConjugateGradient.CG, reginv.py:357-491 (data/model terms :248-355)
 *   GI_CG_BOOTSTRAP <- BootStrap.CG, reginv.py:631-713 (MS with beta^2, no prior, :588-629); the row
 *                      resampling of BootStrap.BSCG (:733-739) enters as per-replicate row
 *                      multiplicities rowweight[c][l] = how often row l was drawn, so Aw is neither
 *                      gathered nor copied and all replicates share every pass over it.
 * ncols == 1 uses the GEMV passes, ncols 2..64 the FP64 tensor-core contractions (ld % 32 == 0).
 * Three passes over Aw per iteration; scalars (alpha, mu, kstep, norms) never leave the device. */
#define GI_CG_REGINV 0
#define GI_CG_BOOTSTRAP 1
typedef struct gi_cg gi_cg;
typedef struct {
    int64_t N, M, ld;
    int32_t ncols;    /* columns (bootstrap replicates) solved together, 1..64 */
    int32_t variant;  /* GI_CG_* */
    gi_reg_params reg; /* reg_kind, nz/ny/nx, beta; alpha/constraint/log_factor are ignored */
    double q;          /* decay of the regularisation factor (reginv.py:400, 640) */
    double stop_tol;   /* REGINV: stop when data/N < tol (0.001, :486); BOOTSTRAP: data < tol (0.1, :694) */
    double rhomin, rhomax; /* bounds on the un-weighted model (reginv.py:434-437) */
} gi_cg_config;
/* G_dev (N x ld weighted kernel), wm/wminv/wmsq_dev (ld-vectors, device) stay owned by the caller.
 * dobs_host [N]; mwapr_host [M] = Wm @ apriorModel or NULL (zero); rowweight_host [ncols][N] or NULL. */
int gi_cg_create(const gi_cg_config *cfg, const double *G_dev, const double *dobs_host,
                 const double *wm_dev, const double *wminv_dev, const double *wmsq_dev,
                 const double *mwapr_host, const double *rowweight_host, void *stream, gi_cg **out);
int gi_cg_destroy(gi_cg *h);
/* Row-sharded CG (Aw partitioned by observation rows over GPUs): the handle is created
 * over this rank's rows (cfg.N = local rows; dobs / rowweight = local rows) and told the global
 * row count; at its exchange points it calls `hook` on its stream's behalf:
 *   what 0: sum-reduce gt_dev[0 : Cp*ld]  (the adjoint output Aw^T (w r) of every column)
 *   what 1: sum-reduce red_dev[0 : Cp]    (sum w q^2 or sum w r^2 per column)
 * in place over every rank, ordered after the work already queued on the handle's stream.  Both
 * buffers are caller-owned (Cp = ncols rounded up to 8, or 1).  Every other quantity is replicated
 * and stays bitwise identical on all ranks (NCCL returns the same bits everywhere). */
/* Wavelet-compressed data terms (reginv.py:250-253, 261-264): d = Awcp @ DWT(mw) replaces Aw @ mw in
 * data(mw) / data_gfun(mw); Aw @ Iw of the step length and Aw^T stay dense, as in the reference.
 * Same arguments as gi_hmc_set_wavelet; single-column, unsharded handles. */
int gi_cg_set_wavelet(gi_cg *h, int32_t kind, int32_t nz, int32_t ny, int32_t nx,
                      const int64_t *indptr_dev, const int32_t *indices_dev, const double *data_dev,
                      int64_t ncoef);
typedef int (*gi_cg_hook)(void *user, int32_t what);
int gi_cg_set_shard(gi_cg *h, int64_t n_total, double *gt_dev, double *red_dev, gi_cg_hook hook,
                    void *user);
/* Run up to maxk iterations from mw0_host [M] = Wm @ initialModel (every column starts there).
 * Outputs (each optional): iters[ncols] = iterations executed (= entries of regul); regul
 * [ncols][maxk]; data_misfit / model_misfit [ncols][maxk] in the reference's list order (REGINV:
 * entry 0 is the start point, `iters` entries; BOOTSTRAP: entry i belongs to iteration i+1). */
int gi_cg_run(gi_cg *h, const double *mw0_host, int32_t maxk, int32_t *iters_host, double *regul_host,
              double *data_misfit_host, double *model_misfit_host);
/* model_inv = WmInv mw [ncols][M]; data_inv = A model_inv [ncols][N] (reginv.py:489-490); mw [ncols][M] */
int gi_cg_get_result(gi_cg *h, double *model_host, double *data_host, double *mw_host);
int64_t gi_cg_launch_count(const gi_cg *h);

/* ---- on-device sample sink ------------------------------------------------ */
/* Replaces the text sink of inversion/hmc.py:241-249 (model.dat: one "%.8f" row of M numbers per
 * accepted sample) and the post-processing of example/uniformgrid/plot_uniform.py:44-135 (np.mean /
 * np.std over the last `last` rows, forward data of both): every sampler handle that has a sink
 * attached adds m = scale .* mw of each ACCEPTED proposal to a Welford accumulator right after the
 * commit, gated on the device by the Metropolis flag and by the window below -- no device->host
 * traffic.  nslots accumulators (one per chain, <= 64) plus pooled statistics over all of them. */
typedef struct gi_stats gi_stats;
int gi_stats_create(int64_t M, int64_t ld, int32_t nslots, gi_stats **out);
int gi_stats_destroy(gi_stats *s);
int gi_stats_reset(gi_stats *s);
/* per slot: ignore the first `skip` accepted samples, then accumulate `take` (<= 0: no limit);
 * the reference's ndraws / nsamples (hmc.py:295, 318) or nsamples - last / last of the plot script */
int gi_stats_window(gi_stats *s, int64_t skip, int64_t take);
/* offer one accepted sample given as a device vector (ld entries); scale_dev may be NULL */
int gi_stats_add(gi_stats *s, int32_t slot, const double *mw_dev, const double *scale_dev, void *stream);
/* mean and population standard deviation (np.std, ddof = 0) of a slot, slot = -1: pooled over all
 * slots.  count = samples accumulated, seen = accepted samples offered.  The _dev variant writes
 * device vectors of ld entries, optionally scaled (scale = Wm gives the weighted models whose forward
 * data Aw (Wm m) are plot_uniform.py:112-114's dpre_mean / dpre_std). */
int gi_stats_result(gi_stats *s, int32_t slot, double *mean_host, double *std_host, int64_t *count,
                    int64_t *seen, void *stream);
int gi_stats_result_dev(gi_stats *s, int32_t slot, const double *scale_dev, double *mean_dev,
                        double *std_dev, int64_t *count, int64_t *seen, void *stream);
int64_t gi_stats_launch_count(const gi_stats *s);
/* attach (or detach with stats = NULL) a sink: gi_hmc adds to `slot`, gi_hmcb chain c to slot c;
 * scale_dev (device, ld entries, e.g. WmInv; caller-owned) maps mw to the recorded model */
int gi_hmc_attach_stats(gi_hmc *h, gi_stats *stats, int32_t slot, const double *scale_dev);
int gi_hmcb_attach_stats(gi_hmcb *h, gi_stats *stats, const double *scale_dev);

#ifdef __cplusplus
}
#endif
#endif /* GRAVINV_B200_H */
