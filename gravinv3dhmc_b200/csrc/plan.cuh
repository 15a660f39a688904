// plan.cuh -- definitions shared by leapfrog.cu (single chain, GEMV) and batched.cu (C chains, DMMA).
#pragma once
#include <math.h>

#include "common.cuh"

namespace gi { struct PeerLaunch; }

struct gi_plan {
    int64_t nrows, M, ld;
    int32_t nchains;
    int fwd_R;
    int64_t fwd_chunk, fwd_nchunks, fwd_rowblocks;
    int64_t adj_rows, adj_nchunks, adj_strips;
    int64_t upd_blocks;
    double *fwd_part, *adj_part, *blockpart;
    unsigned int *counter;
    int64_t workspace_bytes;
    // batched (nchains > 1) tiling, see batched.cu
    int b_nt;                 // n-tiles of 8 chains
    int64_t b_C;              // padded chain count = 8 * b_nt
    int64_t b_npad;           // nrows rounded up to the adjoint stage (16)
    int64_t b_rowblocks, b_kchunk, b_nkc, b_strips;
    double *b_part;           // [b_nkc][b_C][nrows] forward partials
    double *b_blockpart;      // [b_C][upd_blocks][3]
    double *b_scratch_sums;   // [b_C][8]
    unsigned int *b_counter;  // [b_C]
    int64_t b_gp_ldk, b_gp_piece_stride;  // piece-major adjoint output (0 = plain [b_C][ld])
    gi::PeerLaunch *b_peer;   // peer-memory exchange of a row-sharded batch (peer.cuh), or nullptr
};


namespace gi {

// ---------------------------------------------------------------------------------------------
// fused M-vector pass: gradient assembly + regulariser + leapfrog update + clamp
// ---------------------------------------------------------------------------------------------
constexpr int kUpdThreads = 256;
constexpr int kUpdVec = 2;  // groups of 4 consecutive elements per thread

struct UpdateArgs {
    // gradient source: either grad_in (full gradient, e.g. cached at the current state) or
    // 2*sum_k gpart[k] + alpha*dR
    const double *grad_in;
    const double *gpart;
    int64_t gparts;  // number of partial vectors (stride ld)
    const double *x_in, *mw_in, *mwapr, *wmsq, *low, *high;
    double *p, *x_out, *mw_out, *grad_out;
    double pcoef, dt;
    int advance;
    int64_t M, ld;
    gi_reg_params reg;
    double *blockpart;      // [gridDim.x][3]
    unsigned int *counter;  // last-block ticket
    double *sums;           // [2] = Um, [3] = K after update, [4] = K before update
    int save_k0;            // also park K before the update in sums[5] (opening half step)
    const double *p_in;     // momentum is read from here (written to p) when non-null
    // piece-major gradient partials (row-sharded batches all-reduce the adjoint output piece by
    // piece): element j of a chain lives at gpart[(j / gp_ldk) * gp_piece_stride + j % gp_ldk];
    // gp_ldk == 0 means the plain [ld] layout
    int64_t gp_ldk, gp_piece_stride;
    // peer mode (row-sharded batch, peer.cuh): this launch updates the columns [col0, col1) only
    // (col1 == 0: all of them) and the data gradient is the sum of gp_nsrc staged partials, one per
    // source rank, each a [chains][gp_pitch] block (stride gp_src_stride) whose column 0 is matrix column col0:
    // gpart[k * gp_src_stride + (j - col0)] (the chain offset is already in gpart)
    int64_t col0, col1;
    int64_t gp_nsrc, gp_src_stride, gp_pitch;
};

int check_reg(const gi_reg_params *reg, int64_t M);
// batched.cu
int batched_plan_init(gi_plan *p);
void batched_plan_free(gi_plan *p);
// wait_epoch != 0 (peer mode): X is being completed by the peers' pushes of that epoch
int launch_gemm_fwd(gi_plan *p, const double *G, const double *X, cudaStream_t s,
                    unsigned long long wait_epoch = 0);
int launch_gemm_adj(gi_plan *p, const double *G, const double *R, double *out, cudaStream_t s,
                    int piece = 0, int npieces = 1);
int launch_misfit_batched(gi_plan *p, int mode, int64_t n_total, double *d, const double *fix,
                          const double *dobs_c, double *r, double *sums, cudaStream_t s);
int launch_update_batched(gi_plan *p, const gi_reg_params *reg, const double *grad_in,
                          const double *gdata, const double *x_in, const double *mw_in,
                          const double *mwapr, const double *wmsq, const double *low,
                          const double *high, double *pm, double *x_out, double *mw_out,
                          double *grad_out, double dt, const int32_t *L_dev, int step,
                          int uniform_mode, double *sums, cudaStream_t s);

__device__ __forceinline__ double reg_delta(const UpdateArgs &a, int64_t j) {
    return a.mw_in[j] - a.mwapr[j];
}

// body shared by the single-chain and the batched kernels; `copy_x`: when not advancing still
// write x_out = x_in (keeps the ping-pong buffers of frozen / finishing chains consistent).
// A thread owns groups of 4 consecutive elements: every stream (gradient partials, x, p, bounds,
// prior) moves as 256-bit loads/stores, which keeps enough bytes in flight to run this M-vector
// pass near HBM speed (ncu: 1.9 -> ~5 TB/s at M = 2^20, 64 chains); all vectors are padded to ld
// (a multiple of 4) so a group never crosses the end of a buffer.
__device__ __forceinline__ void update_body(const UpdateArgs &a, bool copy_x) {
    __shared__ double scratch[32];
    __shared__ bool is_last;
    double um_t = 0.0, k_after_t = 0.0, k_before_t = 0.0;
    const bool mandatory = a.reg.constraint == GI_CONSTRAINT_MANDATORY;
    const int64_t jend = a.col1 ? min(a.col1, a.M) : a.M;
    for (int v = 0; v < kUpdVec; ++v) {
        const int64_t j0 = a.col0 + ((int64_t)blockIdx.x * kUpdVec + v) * (kUpdThreads * 4) + threadIdx.x * 4;
        if (j0 >= jend) continue;
        const int nvalid = (int)min((int64_t)4, jend - j0);
        double grad[4], mwv[4], pv[4], xin[4];
        ldg4(a.mw_in + j0, mwv[0], mwv[1], mwv[2], mwv[3]);
        {
            const double *ps = (a.p_in ? a.p_in : a.p) + j0;
            ldg4_cg(ps, pv[0], pv[1], pv[2], pv[3]);
        }
        if (a.grad_in) {
            ldg4(a.grad_in + j0, grad[0], grad[1], grad[2], grad[3]);
        } else {
            double gd[4] = {0.0, 0.0, 0.0, 0.0};
            if (a.gp_nsrc) {
                for (int64_t k = 0; k < a.gp_nsrc; ++k) {  // rank order: the same bits whoever owns the slice
                    double t0, t1, t2, t3;
                    ldg_stream4(a.gpart + k * a.gp_src_stride + (j0 - a.col0), t0, t1, t2, t3);
                    gd[0] += t0; gd[1] += t1; gd[2] += t2; gd[3] += t3;
                }
            } else if (a.gp_ldk) {
                const int64_t pc = j0 / a.gp_ldk;  // groups of 4 never straddle a piece
                ldg_stream4(a.gpart + pc * a.gp_piece_stride + (j0 - pc * a.gp_ldk), gd[0], gd[1], gd[2],
                            gd[3]);
            } else {
                for (int64_t k = 0; k < a.gparts; ++k) {
                    double t0, t1, t2, t3;
                    ldg_stream4(a.gpart + k * a.ld + j0, t0, t1, t2, t3);
                    gd[0] += t0; gd[1] += t1; gd[2] += t2; gd[3] += t3;
                }
            }
            double apr[4], w[4] = {0.0, 0.0, 0.0, 0.0};
            ldg4(a.mwapr + j0, apr[0], apr[1], apr[2], apr[3]);
            if (a.reg.reg_kind == GI_REG_MS) ldg4(a.wmsq + j0, w[0], w[1], w[2], w[3]);
            const double beta = a.reg.beta;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (e >= nvalid) { grad[e] = 0.0; continue; }
                const int64_t j = j0 + e;
                const double dl = mwv[e] - apr[e];
                double um = 0.0, gm = 0.0;
                switch (a.reg.reg_kind) {
                    case GI_REG_DAMPING:  // potential.py:775-784
                        um = dl * dl;
                        gm = 2.0 * dl;
                        break;
                    case GI_REG_MS: {  // potential.py:719-736
                        const double sq = dl * dl, den = sq + beta;
                        um = (w[e] * sq) / den;
                        gm = ((2.0 * beta) * w[e] * dl) / (den * den);
                        break;
                    }
                    case GI_REG_SMOOTHNESS:  // potential.py:786-796, D = fd3d (forward differences)
                    case GI_REG_TV: {        // potential.py:798-810
                        const int nx = a.reg.nx, ny = a.reg.ny, nz = a.reg.nz;
                        const int64_t nxy = (int64_t)nx * ny;
                        const int k = (int)(j / nxy);
                        const int rem = (int)(j - (int64_t)k * nxy);
                        const int jy = rem / nx, ix = rem - jy * nx;
                        const bool tv = a.reg.reg_kind == GI_REG_TV;
                        // forward neighbours: rows of D owned by this cell  t = dl - d_next
                        // backward neighbours: rows of D owned by the previous cell  t = d_prev - dl
                        const int64_t offs[3] = {1, nx, nxy};
                        const bool has_f[3] = {ix + 1 < nx, jy + 1 < ny, k + 1 < nz};
                        const bool has_b[3] = {ix > 0, jy > 0, k > 0};
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            if (has_f[d]) {
                                const double t = dl - reg_delta(a, j + offs[d]);
                                if (tv) {
                                    const double sq = sqrt(t * t + beta);
                                    um += sq;
                                    gm += t / sq;
                                } else {
                                    um += t * t;
                                    gm += 2.0 * t;
                                }
                            }
                            if (has_b[d]) {
                                const double t = reg_delta(a, j - offs[d]) - dl;
                                if (tv) gm -= t / sqrt(t * t + beta);
                                else gm -= 2.0 * t;
                            }
                        }
                        break;
                    }
                    default: break;
                }
                um_t += um;
                grad[e] = 2.0 * gd[e] + a.reg.alpha * gm;  // potential.py:708 (2 Aw^T r), :843
            }
        }
        if (a.grad_out) {
            if (nvalid == 4) stg4(a.grad_out + j0, grad[0], grad[1], grad[2], grad[3]);
            else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (e < nvalid) a.grad_out[j0 + e] = grad[e];
            }
        }
        if (a.advance || copy_x) ldg4(a.x_in + j0, xin[0], xin[1], xin[2], xin[3]);
        double hi[4], lo[4], xo[4], mwo[4];
        if (a.advance) {
            ldg4(a.high + j0, hi[0], hi[1], hi[2], hi[3]);
            ldg4(a.low + j0, lo[0], lo[1], lo[2], lo[3]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (e >= nvalid) { xo[e] = 0.0; mwo[e] = 0.0; continue; }
            double p = pv[e];
            k_before_t += p * p;
            p = __dsub_rn(p, __dmul_rn(a.pcoef, grad[e]));  // hmc.py:114,150,152
            if (a.advance) {
                double x = __dadd_rn(xin[e], __dmul_rn(a.dt, p));  // hmc.py:118
                double mw = x;
                if (mandatory) {
                    // hmc.py:135-141 clamp and flip (the while loop runs once)
                    if (x > hi[e]) { x = hi[e]; p = -p; }
                    else if (x < lo[e]) { x = lo[e]; p = -p; }
                    mw = x;
                } else {
                    // potential.py:819-820  mw = (low + high*e**(f x)) / (1 + e**(f x))
                    const double ex = pow(2.718281828459045, a.reg.log_factor * x);
                    mw = (lo[e] + hi[e] * ex) / (1.0 + ex);
                }
                xo[e] = x;
                mwo[e] = mw;
            } else {
                xo[e] = xin[e];
                mwo[e] = mwv[e];
            }
            pv[e] = p;
            k_after_t += p * p;
        }
        if (nvalid == 4) {
            stg4(a.p + j0, pv[0], pv[1], pv[2], pv[3]);
            if (a.advance || copy_x) {
                stg4(a.x_out + j0, xo[0], xo[1], xo[2], xo[3]);
                if (a.mw_out != a.x_out) stg4(a.mw_out + j0, mwo[0], mwo[1], mwo[2], mwo[3]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (e >= nvalid) continue;
                a.p[j0 + e] = pv[e];
                if (a.advance || copy_x) {
                    a.x_out[j0 + e] = xo[e];
                    if (a.mw_out != a.x_out) a.mw_out[j0 + e] = mwo[e];
                }
            }
        }
    }
    // deterministic grid reduction: per-CTA partials, last CTA sums them in index order
    const double um = block_sum(um_t, scratch);
    const double k_after = block_sum(k_after_t, scratch);
    const double k_before = block_sum(k_before_t, scratch);
    if (threadIdx.x == 0) {
        a.blockpart[3 * (int64_t)blockIdx.x + 0] = um;
        a.blockpart[3 * (int64_t)blockIdx.x + 1] = k_after;
        a.blockpart[3 * (int64_t)blockIdx.x + 2] = k_before;
        __threadfence();
        const unsigned int t = atomicAdd(a.counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double s0 = 0, s1 = 0, s2 = 0;
        for (int64_t b = threadIdx.x; b < gridDim.x; b += kUpdThreads) {
            s0 += __ldcg(a.blockpart + 3 * b + 0);
            s1 += __ldcg(a.blockpart + 3 * b + 1);
            s2 += __ldcg(a.blockpart + 3 * b + 2);
        }
        s0 = block_sum(s0, scratch);
        s1 = block_sum(s1, scratch);
        s2 = block_sum(s2, scratch);
        if (threadIdx.x == 0) {
            if (!a.grad_in) a.sums[2] = s0;
            a.sums[3] = 0.5 * s1;  // hmc.py:44-50 with the identity inverse mass
            a.sums[4] = 0.5 * s2;
            if (a.save_k0) a.sums[5] = 0.5 * s2;
            *a.counter = 0;
        }
    }
}

static __global__ void __launch_bounds__(kUpdThreads) update_kernel(UpdateArgs a) { update_body(a, false); }

// ---------------------------------------------------------------------------------------------
// Metropolis test + state commit (hmc.py:156-173)
// ---------------------------------------------------------------------------------------------
struct DevState {
    double U, Ud, Um;  // current state
    gi_hmc_result res;
    double u;  // uniform draw for the next test
};

static __global__ void metropolis_kernel(DevState *st, const double *__restrict__ sums, double alpha, int L,
                                  int force_accept) {
    const double Ud = sums[1], Um = sums[2], Knew = sums[3], K0 = sums[4];
    const double Unew = Ud + alpha * Um;  // potential.py:842
    const double Hcur = K0 + st->U, Hnew = Knew + Unew;
    // hmc.py:167  Hnew < Hcur or u < exp(-(Hnew - Hcur))
    const bool acc = force_accept || (Hnew < Hcur) || (st->u < exp(-(Hnew - Hcur)));
    if (acc) { st->U = Unew; st->Ud = Ud; st->Um = Um; }
    st->res.accept = acc ? 1 : 0;
    st->res.L = L;
    st->res.U = st->U; st->res.U_data = st->Ud; st->res.U_model = st->Um;
    st->res.Hcur = Hcur; st->res.Hnew = Hnew;
    st->res.Unew = Unew; st->res.Unew_data = Ud; st->res.Unew_model = Um;
}

static __global__ void commit_kernel(const DevState *__restrict__ st, int64_t M, int64_t N,
                              const double *__restrict__ x, const double *__restrict__ mw,
                              const double *__restrict__ g, const double *__restrict__ d,
                              double *__restrict__ x_cur, double *__restrict__ mw_cur,
                              double *__restrict__ g_cur, double *__restrict__ d_cur) {
    if (!st->res.accept) return;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < M) {
        x_cur[j] = x[j];
        g_cur[j] = g[j];
        if (mw_cur != x_cur) mw_cur[j] = mw[j];
    }
    if (j < N) d_cur[j] = d[j];
}

// ---------------------------------------------------------------------------------------------
// device RNG for throughput runs: Philox4x32-10 + Box-Muller (not numpy-compatible)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    // uniform in (0, 1): 53 random bits, never exactly 0
    const uint64_t v = (((uint64_t)a << 32) | b) >> 11;
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

static __global__ void philox_normal_kernel(uint64_t seed, uint64_t counter, double sigma, int64_t M,
                                     double *__restrict__ p, DevState *st) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // 2 normals per thread
    uint32_t c[4] = {(uint32_t)t, (uint32_t)(t >> 32), (uint32_t)counter, (uint32_t)(counter >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u1 = u53(c[0], c[1]), u2 = u53(c[2], c[3]);
    const double rad = sqrt(-2.0 * log(u1));
    double s, co;
    sincospi(2.0 * u2, &s, &co);
    const int64_t j = 2 * t;
    if (j < M) p[j] = rad * co * sigma;
    if (j + 1 < M) p[j + 1] = rad * s * sigma;
    if (t == 0 && st) {
        uint32_t c2[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, (uint32_t)counter, (uint32_t)(counter >> 32)};
        philox4x32_10(c2, (uint32_t)seed, (uint32_t)(seed >> 32));
        st->u = u53(c2[0], c2[1]);
    }
}

}  // namespace gi
