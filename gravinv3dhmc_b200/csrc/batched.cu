// batched.cu -- C independent chains batched as columns: the two big passes become dense FP64
// contractions on the tensor cores (DMMA m8n8k4), streaming the kernel matrix once per pass for
// ALL chains.
//
//   forward   D[c][l]  = sum_k Aw[l][k] * X[c][k]        (gemm_fwd_kernel)
//   adjoint   Gt[c][k] = sum_l Aw[l][k] * R[c][l]        (gemm_adj_kernel)
//
// Replaces, for a batch of chains, the same reference lines as leapfrog.cu
// (inversion/potential.py:688-717 data_all, :812-845 misfit_and_grad; inversion/hmc.py:85-177
// _leapfrog); the reference runs one OS process per chain (example/*/run_main.sh:18
// `mpiexec -n 2`), each streaming its own copy of Aw.
//
// Layouts: every per-chain vector is chain-major -- X, P, grad: [C][ld]; D: [C][nrows];
// R: [C][npad] (npad = nrows rounded up to 16, padding zero).  C is padded to 8*NT, NT in 1..8.
//
// Tiling (B200: 148 SMs, 227 KB smem/CTA, FP64 pipe 37.1 TFLOP/s measured for DFMA and DMMA alike,
// profiles/r01_fp64_peak_probe.txt -- at C = 64 both passes are FP64-pipe bound, 16 flop/B):
//   fwd: CTA = 128 rows x one k-chunk, 8 warps x (16 rows x 8*NT chains); 4-stage cp.async ring of
//        [128 rows x 32 voxels] of Aw + [C x 32] of X per stage, XOR-swizzled 16-B chunks so every
//        fragment read is a conflict-free LDS.128; consecutive CTAs share the k-chunk, so X is read
//        from HBM once and served from L2 afterwards; partial[kc][c][row] summed in fixed order.
//   adj: CTA = 256-voxel strip x ALL rows (no split, no partials), 8 warps x (32 voxels x 8*NT
//        chains); 4-stage ring of [16 rows x 256 voxels] of Aw + [C x 16] of R.
// The MMA k-slot <-> memory index assignment is permuted (a sum is order-free as long as A and B
// agree) so that each thread's fragment elements are adjacent in shared memory.
// Deterministic: fixed tile -> slot mapping, no floating-point atomics.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "peer.cuh"
#include "plan.cuh"
#include "sink.cuh"

namespace gi {

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void lds128(const void *smem, double &a, double &b) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(s));
}
__device__ __forceinline__ double lds64(const void *smem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    double a;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(s));
    return a;
}
// D(8x8) += A(8x4, row) * B(4x8, col); lane = 4*g + t holds A[g][t], B[t][g], C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- mbarrier helpers (per-stage full/empty barriers instead of a CTA-wide __syncthreads) --------
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
// arrival that fires once every cp.async this thread has issued so far has landed
__device__ __forceinline__ void mbar_arrive_on_copies(uint64_t *bar) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, int parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
}

#ifndef GI_MBAR
#define GI_MBAR 1
#endif
constexpr int kMbarBytes = 128;  // full[kStages] + empty[kStages] behind the stage ring

constexpr int kGemmThreads = 256;
#ifndef GI_STAGES
#define GI_STAGES 4
#endif
#ifndef GI_FWDK
#define GI_FWDK 32
#endif
#ifndef GI_ADJK
#define GI_ADJK 16
#endif
constexpr int kStages = GI_STAGES;
// tiles in flight ahead of the one being consumed (one less with mbarriers: a stage is refilled two
// iterations after its last read, so a warp never waits for the others)
constexpr int kPrefetch = GI_MBAR ? kStages - 2 : kStages - 1;
constexpr int kFwdRows = 128;  // rows per CTA
constexpr int kFwdK = GI_FWDK;  // voxels per stage (32 or 64)
constexpr int kAdjCols = 256;  // voxels per CTA
constexpr int kAdjK = GI_ADJK;  // rows per stage (16 or 32)

template <int NT>
__host__ __device__ constexpr int fwd_stage_bytes() { return kFwdRows * kFwdK * 8 + 8 * NT * kFwdK * 8; }
template <int NT>
__host__ __device__ constexpr int adj_stage_bytes() { return kAdjK * kAdjCols * 8 + 8 * NT * kAdjK * 8; }

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// WC = warps along the chain dimension: the CTA has 8*WC warps, warp (wr, wc) owns rows
// 16*wr..16*wr+15 and the n-tiles wc*NT/WC .. (wc+1)*NT/WC-1.  WC = 2 doubles the warps per
// scheduler (4 instead of 2), which hides the DMMA issue latency (ncu: stall_wait 39 % at WC = 1).
template <int NT, int WC>
__global__ void __launch_bounds__(kGemmThreads * WC, 1)
gemm_fwd_kernel(const double *__restrict__ G, int64_t ld, const double *__restrict__ X, int64_t nrows,
                int64_t kchunk, int64_t rowblocks, double *__restrict__ part, int64_t nkc, int64_t kc_rot,
                PeerWait pw, PeerMap pm) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int T = kGemmThreads * WC;
    constexpr int C = 8 * NT;
    constexpr int NTW = NT / WC;
    constexpr int GB = kFwdRows * kFwdK * 8;
    constexpr int STAGE = fwd_stage_bytes<NT>();
    constexpr int CPR = kFwdK / 2;          // 16-B chunks per tile row
    constexpr int PITCH = kFwdK * 8;        // bytes per tile row
    constexpr int RPP = T / CPR;            // tile rows covered by one pass of the CTA's threads
    constexpr int GU = kFwdRows / RPP;      // Aw chunks per thread per stage
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int wr = warp & 7, wc = warp >> 3;
    const int64_t tile = blockIdx.x;
    // peer mode: the k-chunks inside this rank's own column slice come first, the others in the
    // order their X slices arrive (kc_rot = first chunk that starts inside the own slice)
    const int64_t kslot = tile / rowblocks, rb = tile - kslot * rowblocks;
    const int64_t kc = (kslot + kc_rot) % nkc;
    const int64_t r0 = rb * kFwdRows;
    const int64_t c0 = kc * kchunk, c1 = min(c0 + kchunk, ld);
    const int ntiles = (int)((c1 - c0) / kFwdK);
    if (pw.epoch) {
        // the X columns of the other ranks' slices are pushed into this buffer by their copy engines;
        // flag[q] >= epoch says slice q of this evaluation's positions has landed (peer.cuh)
        if (tid == 0) {
            for (int q = 0; q < pm.nranks; ++q) {
                if (q == pm.me || pm.col[q + 1] <= pm.col[q] || pm.col[q] >= c1 || pm.col[q + 1] <= c0) continue;
                unsigned long long v, spins = 0;
                do {
                    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(pw.flag + q) : "memory");
                    if (++spins > (1ull << 27)) __trap();
                } while (v < pw.epoch);
            }
        }
        __syncthreads();
    }

    // this thread's copy slots: Aw rows tid/CPR + RPP*u, 16-B chunk tid%CPR (swizzled by row parity)
    const int lrow = tid / CPR, lch = tid % CPR;
    const double *gsrc[GU];
#pragma unroll
    for (int u = 0; u < GU; ++u)
        gsrc[u] = G + min(r0 + lrow + RPP * u, nrows - 1) * ld + c0 + 2 * lch;
    const int gdst = lrow * PITCH + ((lch ^ ((lrow & 1) << 2)) << 4);  // + RPP*u rows: parity unchanged

    constexpr int XU = (C + RPP - 1) / RPP;  // X chunks per thread per stage
    const double *xsrc = X + (int64_t)lrow * ld + c0 + 2 * lch;  // chain lrow; + RPP*u chains
    const int xdst = GB + gdst;

    auto load_stage = [&](int s, int kt) {
        unsigned char *base = smem + s * STAGE;
        const int64_t koff = (int64_t)kt * kFwdK;
#pragma unroll
        for (int u = 0; u < GU; ++u) cp_async16(base + gdst + u * RPP * PITCH, gsrc[u] + koff);
#pragma unroll
        for (int u = 0; u < XU; ++u) {
            if (C % RPP == 0 || lrow + RPP * u < C)
                cp_async16(base + xdst + u * RPP * PITCH, xsrc + (int64_t)u * RPP * ld + koff);
        }
    };

    double acc[2][NTW][2];
#pragma unroll
    for (int rm = 0; rm < 2; ++rm)
#pragma unroll
        for (int j = 0; j < NTW; ++j) acc[rm][j][0] = acc[rm][j][1] = 0.0;

#if GI_MBAR
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + kStages * STAGE), *empty = full + kStages;
    if (tid == 0)
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, T); mbar_init(empty + s, T / 32); }
    __syncthreads();
#endif
#pragma unroll
    for (int s = 0; s < kPrefetch; ++s) {
        if (s < ntiles) {
            load_stage(s, s);
#if GI_MBAR
            mbar_arrive_on_copies(full + s);
#endif
        }
        cp_async_commit();
    }
    const int sw = (g & 1) << 2;  // rows 16wr + 8rm + g and chains 8j + g have the parity of g
    for (int kt = 0; kt < ntiles; ++kt) {
#if GI_MBAR
        mbar_wait(full + kt % kStages, (kt / kStages) & 1);
#else
        cp_async_wait<kStages - 2>();
        __syncthreads();
#endif
        const unsigned char *gs = smem + (kt % kStages) * STAGE;
        const unsigned char *xs = gs + GB + wc * NTW * 8 * PITCH;
#pragma unroll
        for (int q = 0; q < kFwdK / 16; ++q) {
            const int ch_lo = ((8 * q + t) ^ sw) << 4, ch_hi = ((8 * q + 4 + t) ^ sw) << 4;
            double a[2][4];
#pragma unroll
            for (int rm = 0; rm < 2; ++rm) {
                const unsigned char *rowp = gs + (wr * 16 + rm * 8 + g) * PITCH;
                lds128(rowp + ch_lo, a[rm][0], a[rm][1]);
                lds128(rowp + ch_hi, a[rm][2], a[rm][3]);
            }
#pragma unroll
            for (int j = 0; j < NTW; ++j) {
                const unsigned char *rowp = xs + (8 * j + g) * PITCH;
                double b[4];
                lds128(rowp + ch_lo, b[0], b[1]);
                lds128(rowp + ch_hi, b[2], b[3]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    dmma884(acc[0][j][0], acc[0][j][1], a[0][i], b[i]);
                    dmma884(acc[1][j][0], acc[1][j][1], a[1][i], b[i]);
                }
            }
            if (q == 0) {
                // refill a stage every warp has finished with; issued behind the first k-group
                // so the tensor pipe has work while the copies are set up
                const int tl = kt + kPrefetch;
                if (tl < ntiles) {
#if GI_MBAR
                    if (tl >= kStages) mbar_wait(empty + tl % kStages, ((tl - kStages) / kStages) & 1);
                    load_stage(tl % kStages, tl);
                    mbar_arrive_on_copies(full + tl % kStages);
#else
                    load_stage(tl % kStages, tl);
#endif
                }
                cp_async_commit();
            }
        }
#if GI_MBAR
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + kt % kStages);
#endif
    }
    cp_async_wait<0>();
#pragma unroll
    for (int rm = 0; rm < 2; ++rm) {
        const int64_t row = r0 + wr * 16 + rm * 8 + g;
        if (row < nrows) {
#pragma unroll
            for (int j = 0; j < NTW; ++j) {
                const int64_t chain = 8 * (wc * NTW + j) + 2 * t;
                part[(kc * C + chain) * nrows + row] = acc[rm][j][0];
                part[(kc * C + chain + 1) * nrows + row] = acc[rm][j][1];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// adjoint
// ---------------------------------------------------------------------------------------------
template <int NT, int WC>
__global__ void __launch_bounds__(kGemmThreads * WC, 1)
gemm_adj_kernel(const double *__restrict__ G, int64_t ld, const double *__restrict__ R, int64_t npad,
                int64_t nrows, double *__restrict__ out, int64_t strip0, int64_t out_ld,
                int64_t out_col0, PeerOut po, PeerMap pm) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int T = kGemmThreads * WC;
    constexpr int C = 8 * NT;
    constexpr int NTW = NT / WC;
    constexpr int GB = kAdjK * kAdjCols * 8;
    constexpr int STAGE = adj_stage_bytes<NT>();
    constexpr int GU = kAdjK * 128 / T;  // Aw chunks per thread per stage
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int wv = warp & 7, wc = warp >> 3;  // voxel group, chain half
    // strips [strip0, strip0 + gridDim.x) of the matrix; the output of this launch is a [C][out_ld]
    // block whose column 0 is matrix column out_col0 (one "piece" of the piece-major layout)
    const int64_t v0 = (strip0 + blockIdx.x) * kAdjCols;
    const int ntiles = (int)(npad / kAdjK);

    // copy slots: Aw rows (tid>>7) + (T/128)u of the stage, 16-B chunk tid&127 of the 256-voxel strip.
    // All addressing is hoisted: per stage a thread only bumps one pointer by 16 rows (the row
    // clamp is needed in the last, partial stage only), so the per-stage integer work that every
    // warp executes right after the barrier -- while no DMMA is in flight -- stays minimal.
    const int lrow = tid >> 7, lcv = tid & 127;
    const int64_t col = (v0 + 2 * lcv < ld) ? v0 + 2 * lcv : 0;  // strip tail: any valid address
    constexpr int RCP = kAdjK / 2;     // 16-B chunks per R tile row
    constexpr int RPITCH = kAdjK * 8;  // bytes per R tile row
    constexpr int RCH = T / RCP;       // chains covered by one pass of the CTA's threads
    const int rchain = tid / RCP, rec = tid % RCP;
    const int full_tiles = (int)(nrows / kAdjK);  // stages whose 16 rows all exist
    const double *gp = G + (int64_t)lrow * ld + col;  // row lrow of stage 0; + (T/128)u rows; + 16 rows per stage
    const int64_t gstep = (int64_t)(T / 128) * ld;
    int gdst[GU];
#pragma unroll
    for (int u = 0; u < GU; ++u) {
        const int row = lrow + (T / 128) * u;
        gdst[u] = row * 2048 + ((lcv ^ (2 * (row & 3))) << 4);
    }
    const double *rp = R + (int64_t)rchain * npad + 2 * rec;
    const int rdst = GB + rchain * RPITCH + ((rec ^ (2 * (rchain & 3))) << 4);

    auto load_stage = [&](int s, int ot) {
        unsigned char *base = smem + s * STAGE;
        const double *src = gp + (int64_t)ot * kAdjK * ld;
        if (ot < full_tiles) {
#pragma unroll
            for (int u = 0; u < GU; ++u) cp_async16(base + gdst[u], src + u * gstep);
        } else {
            const int64_t o0 = (int64_t)ot * kAdjK;
#pragma unroll
            for (int u = 0; u < GU; ++u)
                cp_async16(base + gdst[u],
                           G + min(o0 + lrow + (T / 128) * u, nrows - 1) * ld + col);
        }
#pragma unroll
        for (int u = 0; u < (C + RCH - 1) / RCH; ++u) {
            if (C % RCH == 0 || rchain + RCH * u < C)
                cp_async16(base + rdst + u * RCH * RPITCH,  // chain & 3 unchanged by + RCH
                           rp + (int64_t)u * RCH * npad + (int64_t)ot * kAdjK);
        }
    };

    double acc[4][NTW][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NTW; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#if GI_MBAR
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + kStages * STAGE), *empty = full + kStages;
    if (tid == 0)
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, T); mbar_init(empty + s, T / 32); }
    __syncthreads();
#endif
#pragma unroll
    for (int s = 0; s < kPrefetch; ++s) {
        if (s < ntiles) {
            load_stage(s, s);
#if GI_MBAR
            mbar_arrive_on_copies(full + s);
#endif
        }
        cp_async_commit();
    }
    const int a_lo = ((16 * wv + g) ^ (2 * t)) << 4, a_hi = ((16 * wv + 8 + g) ^ (2 * t)) << 4;
    for (int ot = 0; ot < ntiles; ++ot) {
#if GI_MBAR
        mbar_wait(full + ot % kStages, (ot / kStages) & 1);
#else
        cp_async_wait<kStages - 2>();
        __syncthreads();
#endif
        const unsigned char *gs = smem + (ot % kStages) * STAGE;
        const unsigned char *rs = gs + GB + wc * NTW * 8 * RPITCH;
#pragma unroll
        for (int kq = 0; kq < kAdjK / 4; ++kq) {
            const unsigned char *rowp = gs + (4 * kq + t) * 2048;
            double a[4];
            lds128(rowp + a_lo, a[0], a[1]);  // voxels 32wv + 2g + {0,1}
            lds128(rowp + a_hi, a[2], a[3]);  // voxels 32wv + 16 + 2g + {0,1}
            const int e = ((4 * kq + t) ^ (4 * (g & 3))) << 3;
#pragma unroll
            for (int j = 0; j < NTW; ++j) {
                const double b = lds64(rs + (8 * j + g) * RPITCH + e);
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma884(acc[i][j][0], acc[i][j][1], a[i], b);
            }
            if (kq == 0) {
                // refill a stage every warp has finished with; issued here, behind the first
                // k-group, so that the tensor pipe already has work while the copies are set up
                const int tl = ot + kPrefetch;
                if (tl < ntiles) {
#if GI_MBAR
                    // the stage was last read for tile tl - kStages (two iterations ago): its
                    // "empty" barrier has long completed unless some warp lags a whole stage
                    if (tl >= kStages) mbar_wait(empty + tl % kStages, ((tl - kStages) / kStages) & 1);
                    load_stage(tl % kStages, tl);
                    mbar_arrive_on_copies(full + tl % kStages);
#else
                    load_stage(tl % kStages, tl);
#endif
                }
                cp_async_commit();
            }
        }
#if GI_MBAR
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + ot % kStages);
#endif
    }
    cp_async_wait<0>();
    if (pm.nranks) {
        // peer mode: this strip's tile goes straight into its OWNER rank's staging block over NVLink
        // ([source rank][chain][owned columns]; a strip never straddles two slices) -- the
        // reduce-scatter of the gradient partials, spread over the whole contraction
        int o = 0;
        while (o + 1 < pm.nranks && v0 >= pm.col[o + 1]) ++o;
        out = po.stage[o] + (int64_t)pm.me * po.src_stride;
        out_ld = po.ldp;
        out_col0 = pm.col[o];
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int64_t v = v0 + 32 * wv + 16 * h + 2 * g;
        if (v < ld) {
#pragma unroll
            for (int j = 0; j < NTW; ++j)
#pragma unroll
                for (int e2 = 0; e2 < 2; ++e2) {
                    double *dst = out + (int64_t)(8 * (wc * NTW + j) + 2 * t + e2) * out_ld + (v - out_col0);
                    asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(dst), "d"(acc[2 * h][j][e2]),
                                 "d"(acc[2 * h + 1][j][e2])
                                 : "memory");
                }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// per-chain data misfit (potential.py:699-706), one CTA per chain
//  mode 0: d = sum_k part[k]; s0 = sum(d + fix); r = (d + fix - s0/n_total) - dobs_c; s1 = sum r^2
//  mode 1: d = sum_k part[k]; s0 only         (row-sharded: s0 is all-reduced before mode 2)
//  mode 2: r, s1 from d and sums[0]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
misfit_batched_kernel(int mode, const double *__restrict__ part, int64_t nkc, int64_t C, int64_t nrows,
                      int64_t npad, int64_t n_total, double *__restrict__ d,
                      const double *__restrict__ fix, const double *__restrict__ dobs_c,
                      double *__restrict__ r, double *__restrict__ sums) {
    __shared__ double scratch[32];
    const int64_t chain = blockIdx.x;
    double *dc = d + chain * nrows, *rc = r + chain * npad, *sc = sums + chain * 8;
    double s0 = 0.0;
    if (mode != 2) {
        for (int64_t row = threadIdx.x; row < nrows; row += 1024) {
            double tsum = 0.0;
            for (int64_t k = 0; k < nkc; ++k) tsum += part[(k * C + chain) * nrows + row];
            dc[row] = tsum;
            s0 += fix ? tsum + fix[row] : tsum;
        }
        s0 = block_sum(s0, scratch);
        if (threadIdx.x == 0) sc[0] = s0;
        if (mode == 1) return;
    } else {
        s0 = sc[0];
    }
    const double mean = n_total > 0 ? s0 / (double)n_total : 0.0;  // n_total <= 0: no mean removal
    double s1 = 0.0;
    for (int64_t row = threadIdx.x; row < npad; row += 1024) {
        double rr = 0.0;
        if (row < nrows) {
            const double tv = dc[row];
            const double dinv = fix ? tv + fix[row] : tv;
            rr = (dinv - mean) - dobs_c[row];
        }
        rc[row] = rr;
        s1 += rr * rr;
    }
    s1 = block_sum(s1, scratch);
    if (threadIdx.x == 0) sc[1] = s1;
}

// ---------------------------------------------------------------------------------------------
// batched fused update: grid (ceil(M/256), C).  Per chain c and trajectory step `step`:
//   step <  L[c]: p -= dt*grad, x += dt*p, clamp         (hmc.py:118-150)
//   step == L[c]: p -= dt/2*grad, grad_out written       (hmc.py:152)
//   step >  L[c]: frozen (the chain's trajectory is over; x copied, sums recomputed identically)
//   step == 0   : the opening half step from the cached gradient (hmc.py:104-118)
//   L[c] == 0   : inactive chain (frozen for the whole proposal)
// ---------------------------------------------------------------------------------------------
struct BatchCtl {
    const int32_t *L;  // [C] or nullptr (uniform: every chain takes `uniform_mode`)
    int32_t step;
    int32_t uniform_mode;  // 0 full step, 1 final half step, 2 frozen, 3 opening half step
    int64_t vec_stride;    // ld
    int64_t nblocks;       // gridDim.x
    // streaming sampler: explicit per-chain modes (4 = leave the chain untouched) and, for mode 3,
    // the queue slot holding the chain's next momentum draw
    int32_t use_modes;
    signed char modes[64];
    signed char slots[64];
    const double *qp;      // [GI_STREAM_QUEUE_DEPTH][C][ld] momentum queue
    int64_t qp_slot_stride;
};

__global__ void __launch_bounds__(kUpdThreads) update_batched_kernel(UpdateArgs a, BatchCtl ctl) {
    const int64_t c = blockIdx.y;
    int mode = ctl.uniform_mode;
    if (ctl.use_modes) {
        mode = ctl.modes[c];
        if (mode == 4) return;
        if (mode == 3 && ctl.qp) a.p_in = ctl.qp + ctl.slots[c] * ctl.qp_slot_stride + c * ctl.vec_stride;
    } else if (ctl.L) {
        const int32_t Lc = ctl.L[c];
        if (Lc == 0) mode = 2;  // inactive chain: nothing moves
        else if (ctl.step == 0) mode = 3;
        else mode = (ctl.step < Lc) ? 0 : (ctl.step == Lc ? 1 : 2);
    }
    const int64_t off = c * ctl.vec_stride;
    if (a.grad_in) a.grad_in += off;
    if (a.gpart) a.gpart += a.gp_nsrc ? c * a.gp_pitch : (a.gp_ldk ? c * a.gp_ldk : off);
    a.x_in += off;
    a.mw_in += off;
    a.p += off;
    if (a.x_out) a.x_out += off;
    if (a.mw_out) a.mw_out += off;
    if (a.grad_out) a.grad_out += off;
    a.blockpart += c * 3 * ctl.nblocks;
    a.counter += c;
    a.sums += c * 8;
    const double dt = a.dt;
    if (mode == 0) { a.pcoef = dt; a.advance = 1; a.grad_out = nullptr; }
    else if (mode == 1) { a.pcoef = 0.5 * dt; a.advance = 0; }
    else if (mode == 2) { a.pcoef = 0.0; a.advance = 0; a.grad_out = nullptr; }
    else { a.pcoef = 0.5 * dt; a.advance = 1; a.grad_out = nullptr; a.save_k0 = 1; }
    update_body(a, true);
}

__global__ void transform_batched_kernel(const double *__restrict__ x, const double *__restrict__ low,
                                         const double *__restrict__ high, double log_factor, int64_t M,
                                         int64_t ld, double *__restrict__ mw) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const int64_t o = (int64_t)blockIdx.y * ld + j;
    const double ex = pow(2.718281828459045, log_factor * x[o]);
    mw[o] = (low[j] + high[j] * ex) / (1.0 + ex);
}

// Metropolis per chain (hmc.py:156-173); K0 of the opening half step is parked in sums[5]
__global__ void metropolis_batched_kernel(DevState *st, const double *__restrict__ sums, double alpha,
                                          const int32_t *__restrict__ L, int C, int force_accept) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double *s = sums + 8 * c;
    DevState *sc = st + c;
    if (L && L[c] == 0 && !force_accept) {  // inactive chain: state untouched, nothing to commit
        sc->res.accept = 0;
        sc->res.L = 0;
        return;
    }
    const double Ud = s[1], Um = s[2], Knew = s[3], K0 = force_accept ? 0.0 : s[5];
    const double Unew = Ud + alpha * Um;
    const double Hcur = K0 + sc->U, Hnew = Knew + Unew;
    const bool acc = force_accept || (Hnew < Hcur) || (sc->u < exp(-(Hnew - Hcur)));
    if (acc) { sc->U = Unew; sc->Ud = Ud; sc->Um = Um; }
    sc->res.accept = acc ? 1 : 0;
    sc->res.L = L ? L[c] : 0;
    sc->res.U = sc->U; sc->res.U_data = sc->Ud; sc->res.U_model = sc->Um;
    sc->res.Hcur = Hcur; sc->res.Hnew = Hnew;
    sc->res.Unew = Unew; sc->res.Unew_data = Ud; sc->res.Unew_model = Um;
}

__global__ void commit_batched_kernel(const DevState *__restrict__ st, int64_t M, int64_t N, int64_t ld,
                                      const double *__restrict__ x, const double *__restrict__ mw,
                                      const double *__restrict__ gnew, const double *__restrict__ d,
                                      double *__restrict__ x_cur, double *__restrict__ mw_cur,
                                      double *__restrict__ g_cur, double *__restrict__ d_cur) {
    const int64_t c = blockIdx.y;
    if (!st[c].res.accept) return;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < M) {
        const int64_t o = c * ld + j;
        x_cur[o] = x[o];
        g_cur[o] = gnew[o];
        if (mw_cur != x_cur) mw_cur[o] = mw[o];
    }
    if (j < N) d_cur[c * N + j] = d[c * N + j];
}

// device draws for chain c: momentum from Philox key (seed + c), like the reference's seed + myrank
__global__ void philox_normal_batched_kernel(uint64_t seed, uint64_t counter, double sigma, int64_t M,
                                             int64_t ld, double *__restrict__ p, DevState *st) {
    const int64_t c = blockIdx.y;
    const uint64_t sd = seed + (uint64_t)c;
    const int64_t tt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cw[4] = {(uint32_t)tt, (uint32_t)(tt >> 32), (uint32_t)counter, (uint32_t)(counter >> 32)};
    philox4x32_10(cw, (uint32_t)sd, (uint32_t)(sd >> 32));
    const double u1 = u53(cw[0], cw[1]), u2 = u53(cw[2], cw[3]);
    const double rad = sqrt(-2.0 * log(u1));
    double s, co;
    sincospi(2.0 * u2, &s, &co);
    const int64_t j = 2 * tt;
    double *pc = p + c * ld;
    if (j < M) pc[j] = rad * co * sigma;
    if (j + 1 < M) pc[j + 1] = rad * s * sigma;
    if (tt == 0) {
        uint32_t c2[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, (uint32_t)counter, (uint32_t)(counter >> 32)};
        philox4x32_10(c2, (uint32_t)sd, (uint32_t)(sd >> 32));
        st[c].u = u53(c2[0], c2[1]);
    }
}

// ---- streaming sampler: every chain runs its proposals back to back (no idling) ----------------
struct StreamCtl {
    signed char fin[64];   // 1: the chain ended a trajectory this step (Metropolis + commit)
    signed char start[64]; // 1: the chain opens a new trajectory this step (u_next is its uniform)
    int32_t rec[64];       // record slot of a finishing chain
    int32_t L[64];         // trajectory length of a finishing chain (for the record)
    double u_next[64];
};

__global__ void stream_finish_kernel(DevState *st, const double *__restrict__ sums, double alpha,
                                     StreamCtl ctl, gi_stream_record *__restrict__ records,
                                     int64_t seq_base, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    DevState *sc = st + c;
    if (ctl.fin[c]) {
        const double *s = sums + 8 * c;
        const double Ud = s[1], Um = s[2], Knew = s[3], K0 = s[5];
        const double Unew = Ud + alpha * Um;
        const double Hcur = K0 + sc->U, Hnew = Knew + Unew;
        const bool acc = (Hnew < Hcur) || (sc->u < exp(-(Hnew - Hcur)));  // hmc.py:167
        if (acc) { sc->U = Unew; sc->Ud = Ud; sc->Um = Um; }
        sc->res.accept = acc ? 1 : 0;
        gi_stream_record *r = records + ctl.rec[c];
        r->chain = c; r->accept = acc ? 1 : 0; r->L = ctl.L[c]; r->reserved = 0;
        r->seq = seq_base;
        r->U = sc->U; r->U_data = sc->Ud; r->U_model = sc->Um; r->Hcur = Hcur; r->Hnew = Hnew;
    }
    if (ctl.start[c]) sc->u = ctl.u_next[c];
}

__global__ void commit_stream_kernel(const DevState *__restrict__ st, StreamCtl ctl, int64_t M, int64_t N,
                                     int64_t ld, const double *__restrict__ x,
                                     const double *__restrict__ mw, const double *__restrict__ gnew,
                                     const double *__restrict__ d, double *__restrict__ x_cur,
                                     double *__restrict__ mw_cur, double *__restrict__ g_cur,
                                     double *__restrict__ d_cur) {
    const int64_t c = blockIdx.y;
    if (!ctl.fin[c] || !st[c].res.accept) return;
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < M) {
        const int64_t o = c * ld + j;
        x_cur[o] = x[o];
        g_cur[o] = gnew[o];
        if (mw_cur != x_cur) mw_cur[o] = mw[o];
    }
    if (j < N) d_cur[c * N + j] = d[c * N + j];
}

// sums[c][col] <-> red[off + c] (contiguous buffers for the scalar all-reduces of the sharded mode)
__global__ void pack_sums_kernel(const double *__restrict__ sums, double *__restrict__ red, int C,
                                 int col, int off, int unpack) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (unpack) const_cast<double *>(sums)[8 * c + col] = red[off + c];
    else red[off + c] = sums[8 * c + col];
}

__global__ void set_u_kernel(DevState *st, const double *__restrict__ u, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) st[c].u = u[c];
}

}  // namespace gi

using namespace gi;

// =============================================================================================
// batched plan pieces (declared in plan.cuh, used by gi_plan_* in leapfrog.cu)
// =============================================================================================
static int pick_nt(int nchains) { return (nchains + 7) / 8; }  // chains padded to a multiple of 8

// warps along the chain dimension per n-tile count (see gemm_fwd_kernel)
#ifndef GI_WC_FWD
#define GI_WC_FWD 2
#endif
#ifndef GI_WC_ADJ
#define GI_WC_ADJ 1
#endif
template <int NT> struct WarpSplit {
    static constexpr int fwd = (NT >= 4 && NT % 2 == 0) ? GI_WC_FWD : 1;
    static constexpr int adj = (NT >= 4 && NT % 2 == 0) ? GI_WC_ADJ : 1;
};

template <int NT>
static int set_smem_attrs() {
    GI_CUDA(cudaFuncSetAttribute(gemm_fwd_kernel<NT, WarpSplit<NT>::fwd>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kStages * fwd_stage_bytes<NT>() + kMbarBytes));
    GI_CUDA(cudaFuncSetAttribute(gemm_adj_kernel<NT, WarpSplit<NT>::adj>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kStages * adj_stage_bytes<NT>() + kMbarBytes));
    return GI_OK;
}

int gi::batched_plan_init(gi_plan *p) {
    GI_REQUIRE(p->nchains >= 2 && p->nchains <= 64, "gi_plan_create: 1..64 chains per plan");
    GI_REQUIRE(p->ld % 32 == 0, "gi_plan_create: batched plans need ld %% 32 == 0");
    p->b_nt = pick_nt(p->nchains);
    p->b_C = 8 * p->b_nt;
    p->b_npad = ceil_div(p->nrows, kAdjK) * kAdjK;
    p->b_rowblocks = ceil_div(p->nrows, kFwdRows);
    // k-chunks: whole waves of one CTA per SM (every CTA does the same work, so a partial last wave
    // is pure loss), at least ~8 of them, and as few chunks as that allows -- the partials
    // [nkc][C][nrows] are re-read by the misfit kernel; chunk a multiple of the 32-voxel stage
    const int sms = sm_count();
    const int64_t max_nkc = std::max<int64_t>(1, ceil_div(p->ld, 8 * kFwdK));  // >= 8 stages per CTA
    int64_t nkc = 1;
    double best = -1.0;
    for (int64_t cand = 1; cand <= std::min<int64_t>(max_nkc, 2048); ++cand) {
        const int64_t tiles = cand * p->b_rowblocks;
        const int64_t waves = ceil_div(tiles, sms);
        const double eff = (double)tiles / (double)(waves * sms);  // fill of the last wave
        // prefer >= 8 waves, then the best fill; ties go to fewer chunks
        const double score = eff - (waves < 8 ? 0.5 * (8 - waves) / 8.0 : 0.0) - 1e-4 * cand;
        if (score > best) { best = score; nkc = cand; }
    }
    p->b_kchunk = ceil_div(ceil_div(p->ld, nkc), kFwdK) * kFwdK;
    p->b_nkc = ceil_div(p->ld, p->b_kchunk);
    p->b_strips = ceil_div(p->ld, kAdjCols);
    const size_t b_part = sizeof(double) * p->b_nkc * p->b_C * p->nrows;
    const size_t b_blk = sizeof(double) * 3 * p->upd_blocks * p->b_C;
    cudaError_t e = cudaMalloc(&p->b_part, b_part);
    if (e == cudaSuccess) e = cudaMalloc(&p->b_blockpart, b_blk);
    if (e == cudaSuccess) e = cudaMalloc(&p->b_scratch_sums, sizeof(double) * 8 * p->b_C);
    if (e == cudaSuccess) e = cudaMalloc(&p->b_counter, sizeof(unsigned int) * p->b_C);
    if (e == cudaSuccess) e = cudaMemset(p->b_counter, 0, sizeof(unsigned int) * p->b_C);
    if (e != cudaSuccess) return cuda_fail(e, "batched plan workspace", __FILE__, __LINE__);
    p->workspace_bytes += (int64_t)(b_part + b_blk);
    switch (p->b_nt) {
        case 1: return set_smem_attrs<1>();
        case 2: return set_smem_attrs<2>();
        case 3: return set_smem_attrs<3>();
        case 4: return set_smem_attrs<4>();
        case 5: return set_smem_attrs<5>();
        case 6: return set_smem_attrs<6>();
        case 7: return set_smem_attrs<7>();
        default: return set_smem_attrs<8>();
    }
}

void gi::batched_plan_free(gi_plan *p) {
    cudaFree(p->b_part);
    cudaFree(p->b_blockpart);
    cudaFree(p->b_scratch_sums);
    cudaFree(p->b_counter);
}

int gi::launch_gemm_fwd(gi_plan *p, const double *G, const double *X, cudaStream_t s,
                        unsigned long long wait_epoch) {
    const unsigned grid = (unsigned)(p->b_nkc * p->b_rowblocks);
    PeerWait pw;
    PeerMap pm;
    memset(&pw, 0, sizeof(pw));
    memset(&pm, 0, sizeof(pm));
    int64_t kc_rot = 0;
    if (p->b_peer) {
        pm = p->b_peer->map;
        kc_rot = p->b_peer->kc_rot;
        pw.flag = p->b_peer->xflag;
        pw.epoch = wait_epoch;
    }
#define GI_FWD(NT)                                                                             \
    gemm_fwd_kernel<NT, WarpSplit<NT>::fwd>                                                    \
        <<<grid, kGemmThreads * WarpSplit<NT>::fwd, kStages * fwd_stage_bytes<NT>() + kMbarBytes, s>>>(     \
        G, p->ld, X, p->nrows, p->b_kchunk, p->b_rowblocks, p->b_part, p->b_nkc, kc_rot, pw, pm)
    switch (p->b_nt) {
        case 1: GI_FWD(1); break;
        case 2: GI_FWD(2); break;
        case 3: GI_FWD(3); break;
        case 4: GI_FWD(4); break;
        case 5: GI_FWD(5); break;
        case 6: GI_FWD(6); break;
        case 7: GI_FWD(7); break;
        default: GI_FWD(8); break;
    }
#undef GI_FWD
    GI_LAUNCH_CHECK();
    return GI_OK;
}

// npieces > 1: `out` is piece-major [npieces][C][ld / npieces]; this launch fills piece `piece`
int gi::launch_gemm_adj(gi_plan *p, const double *G, const double *R, double *out, cudaStream_t s,
                        int piece, int npieces) {
    const int64_t ldk = p->ld / npieces;
    const int64_t strips = npieces == 1 ? p->b_strips : ldk / kAdjCols;
    const int64_t strip0 = piece * strips;
    const unsigned grid = (unsigned)strips;
    double *outp = out ? out + (int64_t)piece * p->b_C * ldk : nullptr;
    const int64_t col0 = piece * ldk;
    PeerOut po;
    PeerMap pm;
    memset(&po, 0, sizeof(po));
    memset(&pm, 0, sizeof(pm));
    if (p->b_peer) {  // the epilogue stores every strip's tile into its owner's staging block
        po = p->b_peer->out;
        pm = p->b_peer->map;
    }
#define GI_ADJ(NT)                                                                             \
    gemm_adj_kernel<NT, WarpSplit<NT>::adj>                                                    \
        <<<grid, kGemmThreads * WarpSplit<NT>::adj, kStages * adj_stage_bytes<NT>() + kMbarBytes, s>>>(     \
        G, p->ld, R, p->b_npad, p->nrows, outp, strip0, ldk, col0, po, pm)
    switch (p->b_nt) {
        case 1: GI_ADJ(1); break;
        case 2: GI_ADJ(2); break;
        case 3: GI_ADJ(3); break;
        case 4: GI_ADJ(4); break;
        case 5: GI_ADJ(5); break;
        case 6: GI_ADJ(6); break;
        case 7: GI_ADJ(7); break;
        default: GI_ADJ(8); break;
    }
#undef GI_ADJ
    GI_LAUNCH_CHECK();
    return GI_OK;
}

// peer mode: restrict an update launch to this rank's column slice and point its data gradient at the
// staged partials of all source ranks; returns the number of CTAs along x (0: empty slice)
static int64_t peer_slice(const gi_plan *p, UpdateArgs &a, bool from_partials) {
    if (!p->b_peer) return p->upd_blocks;
    const PeerLaunch &pl = *p->b_peer;
    a.col0 = pl.map.col[pl.map.me];
    a.col1 = pl.map.col[pl.map.me + 1];
    const int64_t hi = std::min<int64_t>(a.col1, p->M);
    if (hi <= a.col0) return 0;
    if (from_partials) {
        a.gpart = pl.stage_local;
        a.gp_nsrc = pl.map.nranks;
        a.gp_src_stride = pl.out.src_stride;
        a.gp_pitch = pl.out.ldp;
        a.gp_ldk = 0;
    }
    return ceil_div(hi - a.col0, (int64_t)kUpdThreads * kUpdVec * 4);
}

int gi::launch_misfit_batched(gi_plan *p, int mode, int64_t n_total, double *d, const double *fix,
                              const double *dobs_c, double *r, double *sums, cudaStream_t s) {
    misfit_batched_kernel<<<(unsigned)p->b_C, 1024, 0, s>>>(mode, p->b_part, p->b_nkc, p->b_C, p->nrows,
                                                            p->b_npad, n_total, d, fix, dobs_c, r, sums);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

int gi::launch_update_batched(gi_plan *p, const gi_reg_params *reg, const double *grad_in,
                              const double *gdata, const double *x_in, const double *mw_in,
                              const double *mwapr, const double *wmsq, const double *low,
                              const double *high, double *pm, double *x_out, double *mw_out,
                              double *grad_out, double dt, const int32_t *L_dev, int step,
                              int uniform_mode, double *sums, cudaStream_t s) {
    UpdateArgs a;
    memset(&a, 0, sizeof(a));
    a.grad_in = grad_in; a.gpart = gdata; a.gparts = 1;
    a.x_in = x_in; a.mw_in = mw_in; a.mwapr = mwapr; a.wmsq = wmsq; a.low = low; a.high = high;
    a.p = pm; a.x_out = x_out; a.mw_out = mw_out; a.grad_out = grad_out;
    a.dt = dt; a.M = p->M; a.ld = p->ld; a.reg = *reg;
    a.blockpart = p->b_blockpart; a.counter = p->b_counter; a.sums = sums;
    if (gdata) { a.gp_ldk = p->b_gp_ldk; a.gp_piece_stride = p->b_gp_piece_stride; }
    BatchCtl ctl;
    memset(&ctl, 0, sizeof(ctl));
    ctl.L = L_dev; ctl.step = step; ctl.uniform_mode = uniform_mode; ctl.vec_stride = p->ld;
    ctl.nblocks = p->upd_blocks;
    const int64_t blocks = peer_slice(p, a, gdata != nullptr);
    if (blocks == 0) return GI_OK;  // this rank owns no columns
    dim3 grid((unsigned)blocks, (unsigned)p->b_C);
    update_batched_kernel<<<grid, kUpdThreads, 0, s>>>(a, ctl);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

// =============================================================================================
// building blocks for C chains (row-sharded driver) -- same roles as gi_gemv_fwd & co.
// =============================================================================================
extern "C" int gi_plan_batch_info(const gi_plan *p, int32_t *padded_chains, int64_t *padded_rows) {
    GI_REQUIRE(p && p->nchains > 1, "gi_plan_batch_info: not a batched plan");
    if (padded_chains) *padded_chains = (int32_t)p->b_C;
    if (padded_rows) *padded_rows = p->b_npad;
    return GI_OK;
}

extern "C" int gi_gemm_fwd(gi_plan *p, const double *G, const double *X, double *D, void *stream) {
    GI_REQUIRE(p && G && X && D, "gi_gemm_fwd: null pointer");
    GI_REQUIRE(p->nchains > 1, "gi_gemm_fwd: plan was created for one chain (use gi_gemv_fwd)");
    int rc = launch_gemm_fwd(p, G, X, (cudaStream_t)stream);
    if (rc) return rc;
    // mode 1 with a scratch sums row would also do; D = sum of partials is all that is needed here
    return launch_misfit_batched(p, 1, p->nrows, D, nullptr, nullptr, nullptr, p->b_scratch_sums,
                                 (cudaStream_t)stream);
}

extern "C" int gi_data_sum_batched(gi_plan *p, const double *D, const double *fix, double *sums,
                                   void *stream) {
    GI_REQUIRE(p && D && sums && p->nchains > 1, "gi_data_sum_batched: bad argument");
    // s0[c] = sum_l (D[c][l] + fix[l]) -- reuse mode 1 on a single "partial" (D itself)
    misfit_batched_kernel<<<(unsigned)p->b_C, 1024, 0, (cudaStream_t)stream>>>(
        1, D, 1, p->b_C, p->nrows, p->b_npad, p->nrows, const_cast<double *>(D), fix, nullptr, nullptr,
        sums);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_residual_batched(gi_plan *p, const double *D, const double *fix,
                                   const double *dobs_c, int64_t n_total, double *R, double *sums,
                                   void *stream) {
    GI_REQUIRE(p && D && dobs_c && R && sums && p->nchains > 1,
               "gi_residual_batched: bad argument");
    return launch_misfit_batched(p, 2, n_total, const_cast<double *>(D), fix, dobs_c, R, sums,
                                 (cudaStream_t)stream);
}

extern "C" int gi_gemm_adj(gi_plan *p, const double *G, const double *R, double *Gt, void *stream) {
    GI_REQUIRE(p && G && R && Gt && p->nchains > 1, "gi_gemm_adj: bad argument");
    return launch_gemm_adj(p, G, R, Gt, (cudaStream_t)stream);
}

extern "C" int gi_update_batched(gi_plan *p, const gi_reg_params *reg, const double *grad_in_dev,
                                 const double *gdata_dev, const double *x_in, const double *mw_in,
                                 const double *mwapr, const double *wmsq, const double *low,
                                 const double *high, double *pm, double *x_out, double *mw_out,
                                 double *grad_out, double dt, const int32_t *L_dev, int32_t step,
                                 int32_t uniform_mode, double *sums, void *stream) {
    GI_REQUIRE(p && (gdata_dev || grad_in_dev) && x_in && mw_in && mwapr && pm && sums && x_out &&
                   mw_out && low && high && p->nchains > 1,
               "gi_update_batched: bad argument");
    GI_REQUIRE(x_out != x_in, "gi_update_batched: x_in and x_out may not alias");
    GI_REQUIRE(uniform_mode >= 0 && uniform_mode <= 3, "gi_update_batched: bad mode");
    int rc = check_reg(reg, p->M);
    if (rc) return rc;
    GI_REQUIRE(reg->reg_kind != GI_REG_MS || wmsq, "gi_update_batched: MS needs wmsq");
    return launch_update_batched(p, reg, grad_in_dev, gdata_dev, x_in, mw_in, mwapr, wmsq, low, high,
                                 pm, x_out, mw_out, grad_out, dt, L_dev, step, uniform_mode, sums,
                                 (cudaStream_t)stream);
}

// =============================================================================================
// single-GPU device-resident batched sampler
// =============================================================================================
struct gi_hmcb {
    gi_hmc_config cfg;
    int32_t nchains, C;  // user chains, padded chains
    gi_plan *plan;
    const double *G;
    cudaStream_t stream;
    double *x_cur, *mw_cur, *g_cur, *xa, *xb, *mwa, *mwb, *p, *gnew, *gdata;  // [C][ld]
    double *low, *high, *mwapr, *wmsq;                                          // [ld]
    double *d_cur, *d, *r, *dobs_c, *fix;  // [C][N], [C][N], [C][npad], [N], [N]
    double *sums, *u_dev;                  // [C][8], [C]
    int32_t *L_dev;
    DevState *st, *st_host;
    bool has_state;
    int64_t launches;
    // streaming sampler
    struct ChainQ {
        int L_cur, pos, qn, qhead;
        int closed;  // the host has no more proposals for this chain: it may run dry (runway ignores it)
        int qL[GI_STREAM_QUEUE_DEPTH];
        double qu[GI_STREAM_QUEUE_DEPTH];
        int64_t seq;
    } cq[64];
    bool streaming;
    double stream_dt;
    double *qp;  // [GI_STREAM_QUEUE_DEPTH][C][ld] queued momentum draws
    gi_stream_record *rec_dev, *rec_host;
    int64_t rec_cap;
    double *s_xin, *s_xout, *s_mwin, *s_mwout;
    cudaStream_t copy_stream;        // device->host copies of the per-record positions
    cudaEvent_t ev_commit, ev_copied;
    bool copies_pending;
    // row-sharded mode (gi_hmcb_set_shard)
    int64_t n_total;
    gi_shard_hook hook;
    void *hook_user;
    double *g_ext, *red;  // caller-owned piece-major gradient buffer and [2*C] scalar buffer
    int npieces;
    // on-device sample sink (gi_hmcb_attach_stats): chain c -> slot c
    gi_stats *stats;
    const double *stats_scale;
    // peer-memory exchange (gi_hmcb_set_peer, peer.cuh): replaces the hook path
    gi_peer *peer;
    PeerLaunch pl;
    double *sums_part;                 // [C][8] this rank's partial Um / K sums of the last update
    unsigned long long *xepoch_src;    // [8] local ring: the epoch value the copy engines send as X flag
    unsigned long long xepoch;         // pushes so far; the forward pass waits for flag >= epoch
    unsigned long long wait_xa, wait_xb;  // epoch that completes the positions held in xa / xb (0: local)
    cudaStream_t push_stream;
    cudaEvent_t ev_upd, ev_push;
    bool push_pending;
    bool adv_pending;      // gi_hmcb_stream_advance_begin was called, its records are still on the device
    int adv_nrec, adv_done;
    double *own_xa, *own_xb, *own_mwa, *own_mwb;  // the handle's own buffers, replaced by symmetric ones
    int64_t peer_steps;
};

static void hmcb_free(gi_hmcb *h) {
    if (!h) return;
    const bool logc = h->cfg.reg.constraint == GI_CONSTRAINT_LOGARITHMIC;
    double *bufs[] = {h->x_cur, h->g_cur, h->p, h->gnew, h->gdata, h->low, h->high,
                      h->mwapr, h->wmsq, h->d_cur, h->d, h->r, h->dobs_c, h->fix, h->sums, h->u_dev};
    for (double *b : bufs) cudaFree(b);
    if (!h->peer) { cudaFree(h->xa); cudaFree(h->xb); }
    if (logc) {
        cudaFree(h->mw_cur);
        if (!h->peer) { cudaFree(h->mwa); cudaFree(h->mwb); }
    }
    cudaFree(h->L_dev);
    cudaFree(h->st);
    cudaFree(h->qp);
    cudaFree(h->rec_dev);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->ev_commit) cudaEventDestroy(h->ev_commit);
    if (h->ev_copied) cudaEventDestroy(h->ev_copied);
    if (h->rec_host) cudaFreeHost(h->rec_host);
    if (h->st_host) cudaFreeHost(h->st_host);
    if (h->peer) {  // xa / xb (mwa / mwb) live in the caller's symmetric buffer; free the originals
        cudaFree(h->own_xa); cudaFree(h->own_xb);
        if (logc) { cudaFree(h->own_mwa); cudaFree(h->own_mwb); }
        cudaFree(h->sums_part);
        cudaFree(h->xepoch_src);
        if (h->push_stream) cudaStreamDestroy(h->push_stream);
        if (h->ev_upd) cudaEventDestroy(h->ev_upd);
        if (h->ev_push) cudaEventDestroy(h->ev_push);
    }
    gi_plan_destroy(h->plan);
    delete h;
}

#define HB_CUDA(h, call)                                             \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) {                                    \
            hmcb_free(h);                                            \
            return gi::cuda_fail(e__, #call, __FILE__, __LINE__);    \
        }                                                            \
    } while (0)

extern "C" int gi_hmcb_create(const gi_hmc_config *cfg, int32_t nchains, const double *G,
                              const double *dobs_host, const double *gravfix_host,
                              const double *low_host, const double *high_host,
                              const double *mwapr_host, const double *wmsq_host, void *stream,
                              gi_hmcb **out) {
    GI_REQUIRE(cfg && G && dobs_host && low_host && high_host && mwapr_host && out,
               "gi_hmcb_create: null pointer");
    GI_REQUIRE(cfg->N > 0 && cfg->M > 0 && cfg->ld >= cfg->M && cfg->ld % 32 == 0,
               "gi_hmcb_create: bad shape (ld must be a multiple of 32)");
    GI_REQUIRE(nchains >= 2 && nchains <= 64, "gi_hmcb_create: 2..64 chains per batch");
    int rc = check_reg(&cfg->reg, cfg->M);
    if (rc) return rc;
    GI_REQUIRE(!cfg->fixed || gravfix_host, "gi_hmcb_create: fixed=True needs grav_fix");
    gi_hmcb *h = new gi_hmcb();
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->G = G;
    h->nchains = nchains;
    h->n_total = cfg->nocenter ? 0 : cfg->N;
    h->npieces = 1;
    h->stream = (cudaStream_t)stream;
    rc = gi_plan_create(cfg->N, cfg->M, cfg->ld, nchains, &h->plan);
    if (rc) { delete h; return rc; }
    const int64_t C = h->C = h->plan->b_C, ld = cfg->ld, N = cfg->N, npad = h->plan->b_npad;
    const bool logc = cfg->reg.constraint == GI_CONSTRAINT_LOGARITHMIC;
    const size_t bm = sizeof(double) * ld, bcm = bm * C, bn = sizeof(double) * N;
    double **cmv[] = {&h->x_cur, &h->g_cur, &h->xa, &h->xb, &h->p, &h->gnew, &h->gdata};
    for (double **b : cmv) {
        HB_CUDA(h, cudaMalloc(b, bcm));
        HB_CUDA(h, cudaMemsetAsync(*b, 0, bcm, h->stream));
    }
    if (logc) {
        double **lv[] = {&h->mw_cur, &h->mwa, &h->mwb};
        for (double **b : lv) {
            HB_CUDA(h, cudaMalloc(b, bcm));
            HB_CUDA(h, cudaMemsetAsync(*b, 0, bcm, h->stream));
        }
    } else {
        h->mw_cur = h->x_cur; h->mwa = h->xa; h->mwb = h->xb;
    }
    double **mv[] = {&h->low, &h->high, &h->mwapr, &h->wmsq};
    for (double **b : mv) {
        HB_CUDA(h, cudaMalloc(b, bm));
        HB_CUDA(h, cudaMemsetAsync(*b, 0, bm, h->stream));
    }
    HB_CUDA(h, cudaMalloc(&h->d_cur, bn * C));
    HB_CUDA(h, cudaMalloc(&h->d, bn * C));
    HB_CUDA(h, cudaMalloc(&h->r, sizeof(double) * npad * C));
    HB_CUDA(h, cudaMalloc(&h->dobs_c, bn));
    HB_CUDA(h, cudaMalloc(&h->fix, bn));
    HB_CUDA(h, cudaMemsetAsync(h->d_cur, 0, bn * C, h->stream));
    HB_CUDA(h, cudaMemsetAsync(h->d, 0, bn * C, h->stream));
    HB_CUDA(h, cudaMemsetAsync(h->r, 0, sizeof(double) * npad * C, h->stream));
    HB_CUDA(h, cudaMemsetAsync(h->fix, 0, bn, h->stream));
    HB_CUDA(h, cudaMalloc(&h->sums, sizeof(double) * 8 * C));
    HB_CUDA(h, cudaMemsetAsync(h->sums, 0, sizeof(double) * 8 * C, h->stream));
    HB_CUDA(h, cudaMalloc(&h->u_dev, sizeof(double) * C));
    HB_CUDA(h, cudaMalloc(&h->L_dev, sizeof(int32_t) * C));
    HB_CUDA(h, cudaMemsetAsync(h->L_dev, 0, sizeof(int32_t) * C, h->stream));
    HB_CUDA(h, cudaMalloc(&h->st, sizeof(DevState) * C));
    HB_CUDA(h, cudaMemsetAsync(h->st, 0, sizeof(DevState) * C, h->stream));
    HB_CUDA(h, cudaMallocHost(&h->st_host, sizeof(DevState) * C));
    const size_t vm = sizeof(double) * cfg->M;
    HB_CUDA(h, cudaMemcpyAsync(h->low, low_host, vm, cudaMemcpyHostToDevice, h->stream));
    HB_CUDA(h, cudaMemcpyAsync(h->high, high_host, vm, cudaMemcpyHostToDevice, h->stream));
    HB_CUDA(h, cudaMemcpyAsync(h->mwapr, mwapr_host, vm, cudaMemcpyHostToDevice, h->stream));
    if (wmsq_host)
        HB_CUDA(h, cudaMemcpyAsync(h->wmsq, wmsq_host, vm, cudaMemcpyHostToDevice, h->stream));
    {
        double *tmp = new double[N];
        long double acc = 0.0L;
        for (int64_t i = 0; i < N; ++i) acc += dobs_host[i];
        const double mean = cfg->nocenter ? 0.0 : (double)(acc / (long double)N);
        for (int64_t i = 0; i < N; ++i) tmp[i] = dobs_host[i] - mean;
        cudaError_t e = cudaMemcpyAsync(h->dobs_c, tmp, bn, cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        delete[] tmp;
        HB_CUDA(h, e);
    }
    if (cfg->fixed)
        HB_CUDA(h, cudaMemcpyAsync(h->fix, gravfix_host, bn, cudaMemcpyHostToDevice, h->stream));
    HB_CUDA(h, cudaStreamSynchronize(h->stream));
    *out = h;
    return GI_OK;
}

extern "C" int gi_hmcb_destroy(gi_hmcb *h) {
    if (h) cudaStreamSynchronize(h->stream);
    hmcb_free(h);
    return GI_OK;
}

extern "C" int gi_hmcb_set_reg(gi_hmcb *h, const gi_reg_params *reg) {
    GI_REQUIRE(h && reg, "gi_hmcb_set_reg: null pointer");
    int rc = check_reg(reg, h->cfg.M);
    if (rc) return rc;
    GI_REQUIRE(reg->constraint == h->cfg.reg.constraint,
               "gi_hmcb_set_reg: the constraint is fixed at creation");
    h->cfg.reg = *reg;
    h->has_state = false;
    return GI_OK;
}

// ---- peer mode helpers (peer.cuh) ----------------------------------------------------------------
// epoch of the push that completes the positions held in `buf` (0: complete locally)
static unsigned long long hb_wait_epoch(const gi_hmcb *h, const double *buf) {
    if (!h->peer) return 0;
    if (buf == h->xa || buf == h->mwa) return h->wait_xa;
    if (buf == h->xb || buf == h->mwb) return h->wait_xb;
    return 0;
}

// before an update launch: the copy engines must be done reading the buffer it is about to rewrite
static int hb_pre_update(gi_hmcb *h) {
    if (h->peer && h->push_pending) {
        GI_CUDA(cudaStreamWaitEvent(h->stream, h->ev_push, 0));
        h->push_pending = false;
    }
    return GI_OK;
}

// after an update launch: sum the slices' partial Um / K over the ranks (same bits everywhere) and
// leave the next push's epoch where the copy engines will read it
static int hb_post_update(gi_hmcb *h) {
    if (!h->peer) return GI_OK;
    const unsigned long long next = h->peer->xepoch + 1ull;
    h->launches += 1;
    return peer_scalars(h->peer, h->sums_part, h->sums, (int)h->C, 2, 4, h->xepoch_src + (next & 7ull), next,
                        h->stream);
}

// all-gather of the freshly updated positions: this rank's column slice of `xbuf` (and of `mwbuf`
// under the logarithmic constraint) goes to every peer's copy of the buffer on the copy engines of
// a side stream, nearest reader first, each followed by the 8-byte epoch flag the peer's forward
// kernel polls; the forward pass of this rank starts at once on its own slice
static int hb_push(gi_hmcb *h, double *xbuf, double *mwbuf) {
    if (!h->peer) return GI_OK;
    gi_peer *pr = h->peer;
    pr->xepoch += 1ull;
    const unsigned long long ep = pr->xepoch;
    const int P = pr->world, me = pr->rank;
    const int64_t lo = h->pl.map.col[me], w = h->pl.map.col[me + 1] - lo;
    const size_t pitch = sizeof(double) * h->cfg.ld;
    if (P > 1) {
        GI_CUDA(cudaEventRecord(h->ev_upd, h->stream));
        GI_CUDA(cudaStreamWaitEvent(h->push_stream, h->ev_upd, 0));
        for (int k = 1; k < P; ++k) {
            const int q = (me - k + P) % P;  // rank me - 1 reads slice `me` right after its own
            if (w > 0) {
                double *dst = reinterpret_cast<double *>(pr->base[q] + (reinterpret_cast<unsigned char *>(xbuf) - pr->base[me]));
                GI_CUDA(cudaMemcpy2DAsync(dst + lo, pitch, xbuf + lo, pitch, sizeof(double) * w, (size_t)h->C,
                                          cudaMemcpyDeviceToDevice, h->push_stream));
                pr->nvlink_bytes += (int64_t)sizeof(double) * w * h->C;
                if (mwbuf != xbuf) {
                    dst = reinterpret_cast<double *>(pr->base[q] + (reinterpret_cast<unsigned char *>(mwbuf) - pr->base[me]));
                    GI_CUDA(cudaMemcpy2DAsync(dst + lo, pitch, mwbuf + lo, pitch, sizeof(double) * w, (size_t)h->C,
                                              cudaMemcpyDeviceToDevice, h->push_stream));
                    pr->nvlink_bytes += (int64_t)sizeof(double) * w * h->C;
                }
            }
            unsigned long long *flag = reinterpret_cast<unsigned long long *>(pr->base[q] + kPeerFlagX) + me;
            GI_CUDA(cudaMemcpyAsync(flag, h->xepoch_src + (ep & 7ull), sizeof(unsigned long long),
                                    cudaMemcpyDeviceToDevice, h->push_stream));
        }
        GI_CUDA(cudaEventRecord(h->ev_push, h->push_stream));
        h->push_pending = true;
    }
    if (xbuf == h->xa) h->wait_xa = ep;
    else h->wait_xb = ep;
    return GI_OK;
}

// d, r, sums[.][0..1] and the gradient partials for the positions mw_in; in row-sharded mode the two
// exchange steps go through the caller's hook and the adjoint output is reduced piece by piece
// while the next piece is being computed -- or, in peer mode, through peer memory: the scalars by
// the slot-table kernel, the gradient partials by the adjoint kernel's own epilogue
static int hb_data_pass(gi_hmcb *h, const double *mw_in) {
    gi_plan *p = h->plan;
    cudaStream_t s = h->stream;
    const double *fix = h->cfg.fixed ? h->fix : nullptr;
    int rc = launch_gemm_fwd(p, h->G, mw_in, s, hb_wait_epoch(h, mw_in));
    if (rc) return rc;
    if (h->peer) {
        const int Cc = (int)h->C;
        rc = launch_misfit_batched(p, 1, h->n_total, h->d, fix, h->dobs_c, h->r, h->sums, s);
        if (!rc) rc = peer_scalars(h->peer, h->sums, h->sums, Cc, 0, 1, nullptr, 0, s);  // sum d -> mean
        if (!rc) rc = launch_misfit_batched(p, 2, h->n_total, h->d, fix, h->dobs_c, h->r, h->sums, s);
        if (!rc) rc = launch_gemm_adj(p, h->G, h->r, nullptr, s);  // tiles land in their owners' staging blocks
        // sum r^2 -- and the barrier of the reduce-scatter: a rank's flag follows its contraction
        if (!rc) rc = peer_scalars(h->peer, h->sums, h->sums, Cc, 1, 1, nullptr, 0, s);
        h->peer->nvlink_bytes += (int64_t)sizeof(double) * h->C * (h->cfg.ld - (h->pl.map.col[h->pl.map.me + 1] - h->pl.map.col[h->pl.map.me]));
        h->launches += 6;
        h->peer_steps += 1;
        return rc;
    }
    if (!h->hook) {
        rc = launch_misfit_batched(p, 0, h->n_total, h->d, fix, h->dobs_c, h->r, h->sums, s);
        if (!rc) rc = launch_gemm_adj(p, h->G, h->r, h->gdata, s);
        h->launches += 3;
        return rc;
    }
    const int C = (int)h->C;
    rc = launch_misfit_batched(p, 1, h->n_total, h->d, fix, h->dobs_c, h->r, h->sums, s);
    if (rc) return rc;
    pack_sums_kernel<<<1, 64, 0, s>>>(h->sums, h->red, C, 0, 0, 0);
    GI_LAUNCH_CHECK();
    GI_REQUIRE(h->hook(h->hook_user, 0, 0, 0) == 0, "shard hook failed (data sums)");
    pack_sums_kernel<<<1, 64, 0, s>>>(h->sums, h->red, C, 0, 0, 1);
    GI_LAUNCH_CHECK();
    rc = launch_misfit_batched(p, 2, h->n_total, h->d, fix, h->dobs_c, h->r, h->sums, s);
    if (rc) return rc;
    pack_sums_kernel<<<1, 64, 0, s>>>(h->sums, h->red, C, 1, C, 0);
    GI_LAUNCH_CHECK();
    GI_REQUIRE(h->hook(h->hook_user, 2, 0, 1) == 0, "shard hook failed (residual norms)");
    for (int pc = 0; pc < h->npieces; ++pc) {
        rc = launch_gemm_adj(p, h->G, h->r, h->g_ext, s, pc, h->npieces);
        if (rc) return rc;
        GI_REQUIRE(h->hook(h->hook_user, 1, pc, 1) == 0, "shard hook failed (gradient piece)");
    }
    GI_REQUIRE(h->hook(h->hook_user, 3, 0, 0) == 0, "shard hook failed (wait)");
    pack_sums_kernel<<<1, 64, 0, s>>>(h->sums, h->red, C, 1, C, 1);
    GI_LAUNCH_CHECK();
    h->launches += 7 + h->npieces;
    return GI_OK;
}

extern "C" int gi_hmcb_set_shard(gi_hmcb *h, int64_t n_total, const double *dobs_c_host,
                                 double *gdata_dev, int32_t npieces, double *red_dev,
                                 gi_shard_hook hook, void *user) {
    GI_REQUIRE(h && hook && gdata_dev && red_dev && dobs_c_host, "gi_hmcb_set_shard: null pointer");
    GI_REQUIRE(n_total >= h->cfg.N, "gi_hmcb_set_shard: n_total is the GLOBAL observation count");
    GI_REQUIRE(!h->cfg.nocenter, "gi_hmcb_set_shard: the joint data term is a single-GPU path");
    GI_REQUIRE(npieces >= 1 && h->cfg.ld % npieces == 0 && (npieces == 1 || (h->cfg.ld / npieces) % kAdjCols == 0),
               "gi_hmcb_set_shard: ld / npieces must be a multiple of 256");
    GI_CUDA(cudaMemcpyAsync(h->dobs_c, dobs_c_host, sizeof(double) * h->cfg.N, cudaMemcpyHostToDevice,
                            h->stream));
    GI_CUDA(cudaStreamSynchronize(h->stream));
    h->n_total = n_total;
    h->hook = hook;
    h->hook_user = user;
    h->g_ext = gdata_dev;
    h->red = red_dev;
    h->npieces = npieces;
    h->plan->b_gp_ldk = h->cfg.ld / npieces;
    h->plan->b_gp_piece_stride = h->C * (h->cfg.ld / npieces);
    h->has_state = false;
    return GI_OK;
}

// ---- peer mode set-up ------------------------------------------------------------------------------
static void peer_columns(int64_t ld, int world, int64_t *col) {
    const int64_t nstrips = ceil_div(ld, (int64_t)kAdjCols);
    for (int q = 0; q <= world; ++q) col[q] = std::min<int64_t>(ld, kAdjCols * ((q * nstrips) / world));
    col[world] = ld;
}

static int64_t peer_pitch(int64_t ld, int world) {
    int64_t col[kPeerMax + 1], w = 0;
    peer_columns(ld, world, col);
    for (int q = 0; q < world; ++q) w = std::max(w, col[q + 1] - col[q]);
    return std::max<int64_t>(w, 4);
}

// host-only: the column slices of a peer-mode batch, col[0..world] (slice q = [col[q], col[q+1]))
extern "C" int gi_peer_columns(int64_t ld, int32_t world, int64_t *col) {
    GI_REQUIRE(col && ld > 0 && world >= 1 && world <= kPeerMax, "gi_peer_columns: bad argument");
    peer_columns(ld, world, col);
    return GI_OK;
}

extern "C" int64_t gi_hmcb_peer_bytes(const gi_hmcb *h, int32_t world) {
    if (!h || world < 1 || world > kPeerMax) return -1;
    const bool logc = h->cfg.reg.constraint == GI_CONSTRAINT_LOGARITHMIC;
    const int64_t ldp = peer_pitch(h->cfg.ld, world);
    return kPeerCtlBytes + (int64_t)sizeof(double) * h->C * (world * ldp + (logc ? 4 : 2) * h->cfg.ld);
}

extern "C" int gi_hmcb_owned_columns(const gi_hmcb *h, int64_t *lo, int64_t *hi) {
    GI_REQUIRE(h && lo && hi, "gi_hmcb_owned_columns: null pointer");
    *lo = h->peer ? h->pl.map.col[h->pl.map.me] : 0;
    *hi = h->peer ? std::min<int64_t>(h->pl.map.col[h->pl.map.me + 1], h->cfg.M) : h->cfg.M;
    return GI_OK;
}

extern "C" int gi_hmcb_set_peer(gi_hmcb *h, gi_peer *peer, int64_t n_total, const double *dobs_c_host) {
    GI_REQUIRE(h && peer && dobs_c_host, "gi_hmcb_set_peer: null pointer");
    GI_REQUIRE(!h->hook && !h->peer, "gi_hmcb_set_peer: the handle already has an exchange path");
    GI_REQUIRE(peer->connected, "gi_hmcb_set_peer: call gi_peer_connect first");
    GI_REQUIRE(n_total >= h->cfg.N, "gi_hmcb_set_peer: n_total is the GLOBAL observation count");
    if (h->cfg.nocenter) n_total = 0;  // no mean removal: nothing global about the residual
    GI_REQUIRE(peer->bytes >= gi_hmcb_peer_bytes(h, peer->world),
               "gi_hmcb_set_peer: the symmetric buffer is smaller than gi_hmcb_peer_bytes()");
    const bool logc = h->cfg.reg.constraint == GI_CONSTRAINT_LOGARITHMIC;
    const int64_t ld = h->cfg.ld, C = h->C;
    const int P = peer->world, me = peer->rank;
    cudaStream_t s = h->stream;
    PeerLaunch &pl = h->pl;
    memset(&pl, 0, sizeof(pl));
    pl.map.nranks = P;
    pl.map.me = me;
    peer_columns(ld, P, pl.map.col);
    const int64_t ldp = peer_pitch(ld, P);
    // symmetric layout: control block | staging [P][C][ldp] | xa [C][ld] | xb | (mwa | mwb)
    const int64_t off_stage = kPeerCtlBytes;
    const int64_t off_xa = off_stage + (int64_t)sizeof(double) * P * C * ldp;
    const int64_t bx = (int64_t)sizeof(double) * C * ld;
    for (int q = 0; q < P; ++q) pl.out.stage[q] = reinterpret_cast<double *>(peer->base[q] + off_stage);
    pl.out.src_stride = C * ldp;
    pl.out.ldp = ldp;
    pl.stage_local = pl.out.stage[me];
    pl.xflag = reinterpret_cast<const unsigned long long *>(peer->base[me] + kPeerFlagX);
    // own slice first: the first k-chunk that starts inside it
    pl.kc_rot = ceil_div(pl.map.col[me], h->plan->b_kchunk) % h->plan->b_nkc;
    GI_CUDA(cudaMemsetAsync(peer->base[me] + off_stage, 0, (size_t)(peer->bytes - off_stage), s));
    GI_CUDA(cudaMalloc(&h->sums_part, sizeof(double) * 8 * C));
    GI_CUDA(cudaMemsetAsync(h->sums_part, 0, sizeof(double) * 8 * C, s));
    GI_CUDA(cudaMalloc(&h->xepoch_src, sizeof(unsigned long long) * 8));
    GI_CUDA(cudaMemsetAsync(h->xepoch_src, 0, sizeof(unsigned long long) * 8, s));
    GI_CUDA(cudaStreamCreateWithFlags(&h->push_stream, cudaStreamNonBlocking));
    GI_CUDA(cudaEventCreateWithFlags(&h->ev_upd, cudaEventDisableTiming));
    GI_CUDA(cudaEventCreateWithFlags(&h->ev_push, cudaEventDisableTiming));
    h->own_xa = h->xa; h->own_xb = h->xb; h->own_mwa = h->mwa; h->own_mwb = h->mwb;
    h->xa = reinterpret_cast<double *>(peer->base[me] + off_xa);
    h->xb = reinterpret_cast<double *>(peer->base[me] + off_xa + bx);
    if (logc) {
        h->mwa = reinterpret_cast<double *>(peer->base[me] + off_xa + 2 * bx);
        h->mwb = reinterpret_cast<double *>(peer->base[me] + off_xa + 3 * bx);
    } else {
        h->mwa = h->xa; h->mwb = h->xb;
    }
    GI_CUDA(cudaMemcpyAsync(h->dobs_c, dobs_c_host, sizeof(double) * h->cfg.N, cudaMemcpyHostToDevice, s));
    GI_CUDA(cudaStreamSynchronize(s));
    h->n_total = n_total;
    h->peer = peer;
    h->plan->b_peer = &h->pl;
    h->wait_xa = h->wait_xb = 0;
    h->push_pending = false;
    h->has_state = false;
    return GI_OK;
}

// one batched misfit_and_grad at (x_in, mw_in) + the fused update
static int hb_grad_eval_and_update(gi_hmcb *h, const double *x_in, const double *mw_in, double *x_out,
                                   double *mw_out, double *grad_out, double dt, const int32_t *L_dev,
                                   int step, int uniform_mode) {
    gi_plan *p = h->plan;
    cudaStream_t s = h->stream;
    int rc = hb_data_pass(h, mw_in);
    if (!rc) rc = hb_pre_update(h);
    if (rc) return rc;
    const double *gsrc = h->peer ? h->pl.stage_local : (h->hook ? h->g_ext : h->gdata);
    rc = launch_update_batched(p, &h->cfg.reg, nullptr, gsrc, x_in, mw_in,
                               h->mwapr, h->wmsq, h->low, h->high, h->p, x_out, mw_out, grad_out, dt,
                               L_dev, step, uniform_mode, h->peer ? h->sums_part : h->sums, s);
    h->launches += 1;
    if (!rc) rc = hb_post_update(h);
    return rc;
}

extern "C" int gi_hmcb_set_state(gi_hmcb *h, const double *x_host) {
    GI_REQUIRE(h && x_host, "gi_hmcb_set_state: null pointer");
    cudaStream_t s = h->stream;
    const int64_t M = h->cfg.M, ld = h->cfg.ld, C = h->C;
    GI_CUDA(cudaMemsetAsync(h->x_cur, 0, sizeof(double) * ld * C, s));
    GI_CUDA(cudaMemcpy2DAsync(h->x_cur, sizeof(double) * ld, x_host, sizeof(double) * M,
                              sizeof(double) * M, h->nchains, cudaMemcpyHostToDevice, s));
    if (h->cfg.reg.constraint == GI_CONSTRAINT_LOGARITHMIC) {
        dim3 grid((unsigned)ceil_div(M, 256), (unsigned)C);
        transform_batched_kernel<<<grid, 256, 0, s>>>(h->x_cur, h->low, h->high, h->cfg.reg.log_factor,
                                                      M, ld, h->mw_cur);
        GI_LAUNCH_CHECK();
        h->launches += 1;
    }
    GI_CUDA(cudaMemsetAsync(h->p, 0, sizeof(double) * ld * C, s));
    // gradient at the start state: "final half step" mode with p = 0 writes grad_out = g_cur;
    // x_out is a scratch copy
    int rc = hb_grad_eval_and_update(h, h->x_cur, h->mw_cur, h->xa, h->mwa, h->g_cur, 0.0, nullptr, 0, 1);
    if (rc) return rc;
    metropolis_batched_kernel<<<1, 64, 0, s>>>(h->st, h->sums, h->cfg.reg.alpha, nullptr, (int)C, 1);
    GI_LAUNCH_CHECK();
    GI_CUDA(cudaMemcpyAsync(h->d_cur, h->d, sizeof(double) * h->cfg.N * C, cudaMemcpyDeviceToDevice, s));
    h->launches += 1;
    GI_CUDA(cudaStreamSynchronize(s));
    h->has_state = true;
    return GI_OK;
}

extern "C" int gi_hmcb_get_state(gi_hmcb *h, double *x_host, double *d_host, double *mw_host) {
    GI_REQUIRE(h && h->has_state, "gi_hmcb_get_state: no state set");
    cudaStream_t s = h->stream;
    const int64_t M = h->cfg.M, ld = h->cfg.ld, N = h->cfg.N;
    if (x_host)
        GI_CUDA(cudaMemcpy2DAsync(x_host, sizeof(double) * M, h->x_cur, sizeof(double) * ld,
                                  sizeof(double) * M, h->nchains, cudaMemcpyDeviceToHost, s));
    if (mw_host)
        GI_CUDA(cudaMemcpy2DAsync(mw_host, sizeof(double) * M, h->mw_cur, sizeof(double) * ld,
                                  sizeof(double) * M, h->nchains, cudaMemcpyDeviceToHost, s));
    if (d_host)
        GI_CUDA(cudaMemcpyAsync(d_host, h->d_cur, sizeof(double) * N * h->nchains,
                                cudaMemcpyDeviceToHost, s));
    GI_CUDA(cudaStreamSynchronize(s));
    return GI_OK;
}

extern "C" int gi_hmcb_get_misfit(gi_hmcb *h, double *U, double *Ud, double *Um, double *grad_host) {
    GI_REQUIRE(h && h->has_state, "gi_hmcb_get_misfit: no state set");
    cudaStream_t s = h->stream;
    GI_CUDA(cudaMemcpyAsync(h->st_host, h->st, sizeof(DevState) * h->C, cudaMemcpyDeviceToHost, s));
    if (grad_host)
        GI_CUDA(cudaMemcpy2DAsync(grad_host, sizeof(double) * h->cfg.M, h->g_cur,
                                  sizeof(double) * h->cfg.ld, sizeof(double) * h->cfg.M, h->nchains,
                                  cudaMemcpyDeviceToHost, s));
    GI_CUDA(cudaStreamSynchronize(s));
    for (int c = 0; c < h->nchains; ++c) {
        if (U) U[c] = h->st_host[c].U;
        if (Ud) Ud[c] = h->st_host[c].Ud;
        if (Um) Um[c] = h->st_host[c].Um;
    }
    return GI_OK;
}

// trajectories of all chains from their current states; momenta already in h->p, L in h->L_dev
static int hb_run(gi_hmcb *h, int32_t Lmax, double dt, const int32_t *L_dev, gi_hmc_result *results,
                  double *trace_x_host, double *trace_U_host, bool metropolis) {
    cudaStream_t s = h->stream;
    gi_plan *p = h->plan;
    const int64_t M = h->cfg.M, ld = h->cfg.ld, C = h->C, N = h->cfg.N;
    const int nc = h->nchains;
    const bool logc = h->cfg.reg.constraint == GI_CONSTRAINT_LOGARITHMIC;
    const double alpha = h->cfg.reg.alpha;
    double *hs = nullptr;
    if (trace_U_host) hs = new double[8 * C];
    auto trace = [&](int i, const double *xbuf) -> int {
        if (trace_x_host)
            GI_CUDA(cudaMemcpy2DAsync(trace_x_host + (int64_t)i * nc * M, sizeof(double) * M, xbuf,
                                      sizeof(double) * ld, sizeof(double) * M, nc,
                                      cudaMemcpyDeviceToHost, s));
        if (trace_U_host) {
            if (i == 0) {
                GI_CUDA(cudaMemcpyAsync(h->st_host, h->st, sizeof(DevState) * C, cudaMemcpyDeviceToHost, s));
                GI_CUDA(cudaStreamSynchronize(s));
                for (int c = 0; c < nc; ++c) trace_U_host[c] = h->st_host[c].U;
            } else {
                GI_CUDA(cudaMemcpyAsync(hs, h->sums, sizeof(double) * 8 * C, cudaMemcpyDeviceToHost, s));
                GI_CUDA(cudaStreamSynchronize(s));
                for (int c = 0; c < nc; ++c)
                    trace_U_host[(int64_t)i * nc + c] = hs[8 * c + 1] + alpha * hs[8 * c + 2];
            }
        }
        return GI_OK;
    };
    int rc = trace(0, h->x_cur);
    // opening half step from the cached gradients (hmc.py:104-118); K0 -> sums[c][5]
    if (!rc) rc = hb_pre_update(h);
    if (!rc)
        rc = launch_update_batched(p, &h->cfg.reg, h->g_cur, nullptr, h->x_cur, h->mw_cur, h->mwapr,
                                   h->wmsq, h->low, h->high, h->p, h->xa, h->mwa, nullptr, dt, L_dev,
                                   0, 3, h->peer ? h->sums_part : h->sums, s);
    h->launches += 1;
    if (!rc) rc = hb_post_update(h);
    if (!rc) rc = hb_push(h, h->xa, h->mwa);
    double *xin = h->xa, *xout = h->xb, *mwin = h->mwa, *mwout = h->mwb;
    for (int i = 1; i <= Lmax && !rc; ++i) {
        rc = hb_grad_eval_and_update(h, xin, mwin, xout, mwout, h->gnew, dt, L_dev, i,
                                     metropolis ? 0 : 0);
        if (!rc) rc = hb_push(h, xout, mwout);
        if (!rc) rc = trace(i, xin);
        double *tp = xin; xin = xout; xout = tp;
        if (logc) { tp = mwin; mwin = mwout; mwout = tp; }
        else { mwin = xin; mwout = xout; }
    }
    delete[] hs;
    // peer mode: the positions the commit copies were completed by the last push
    if (!rc && h->peer) rc = peer_wait_x(h->peer, h->pl.map, hb_wait_epoch(h, xin), s);
    if (rc || !metropolis) return rc;
    metropolis_batched_kernel<<<1, 64, 0, s>>>(h->st, h->sums, alpha, L_dev, (int)C, 0);
    GI_LAUNCH_CHECK();
    const int64_t n = std::max(M, N);
    dim3 grid((unsigned)ceil_div(n, 256), (unsigned)C);
    commit_batched_kernel<<<grid, 256, 0, s>>>(h->st, M, N, ld, xin, mwin, h->gnew, h->d, h->x_cur,
                                               h->mw_cur, h->g_cur, h->d_cur);
    GI_LAUNCH_CHECK();
    h->launches += 2;
    if (h->stats) {  // accepted models go to the sink without leaving the device
        StatsMap map;
        memset(&map, 0, sizeof(map));
        for (int c = 0; c < nc; ++c) map.fin[c] = 1;  // chains that sat out have accept == 0
        rc = stats_add_chains(h->stats, map, h->st, h->mw_cur, ld, nc, 0, h->stats_scale, s);
        if (rc) return rc;
        h->launches += 2;
    }
    GI_CUDA(cudaMemcpyAsync(h->st_host, h->st, sizeof(DevState) * C, cudaMemcpyDeviceToHost, s));
    GI_CUDA(cudaStreamSynchronize(s));
    if (results)
        for (int c = 0; c < nc; ++c) results[c] = h->st_host[c].res;
    return GI_OK;
}

extern "C" int gi_hmcb_attach_stats(gi_hmcb *h, gi_stats *stats, const double *scale_dev) {
    GI_REQUIRE(h, "gi_hmcb_attach_stats: null handle");
    h->stats = stats;
    h->stats_scale = scale_dev;
    return GI_OK;
}

static int hb_set_L(gi_hmcb *h, const int32_t *L_host, int32_t *Lmax) {
    int32_t tmp[64];
    *Lmax = 0;
    for (int c = 0; c < h->C; ++c) {
        tmp[c] = c < h->nchains ? L_host[c] : 0;
        GI_REQUIRE(tmp[c] >= 0, "gi_hmcb_propose: L must be >= 0 (0 = chain sits this proposal out)");
        *Lmax = std::max(*Lmax, tmp[c]);
    }
    GI_CUDA(cudaMemcpyAsync(h->L_dev, tmp, sizeof(int32_t) * h->C, cudaMemcpyHostToDevice, h->stream));
    GI_CUDA(cudaStreamSynchronize(h->stream));  // tmp is a stack buffer
    return GI_OK;
}

extern "C" int gi_hmcb_propose(gi_hmcb *h, const double *p0_host, const int32_t *L_host, double dt,
                               const double *u_host, gi_hmc_result *results, double *trace_x_host,
                               double *trace_U_host) {
    GI_REQUIRE(h && p0_host && L_host && u_host && results, "gi_hmcb_propose: null pointer");
    GI_REQUIRE(h->has_state, "gi_hmcb_propose: call gi_hmcb_set_state first");
    cudaStream_t s = h->stream;
    int32_t Lmax = 0;
    int rc = hb_set_L(h, L_host, &Lmax);
    if (rc) return rc;
    GI_CUDA(cudaMemsetAsync(h->p, 0, sizeof(double) * h->cfg.ld * h->C, s));
    GI_CUDA(cudaMemcpy2DAsync(h->p, sizeof(double) * h->cfg.ld, p0_host, sizeof(double) * h->cfg.M,
                              sizeof(double) * h->cfg.M, h->nchains, cudaMemcpyHostToDevice, s));
    GI_CUDA(cudaMemsetAsync(h->u_dev, 0, sizeof(double) * h->C, s));
    GI_CUDA(cudaMemcpyAsync(h->u_dev, u_host, sizeof(double) * h->nchains, cudaMemcpyHostToDevice, s));
    set_u_kernel<<<1, 64, 0, s>>>(h->st, h->u_dev, (int)h->C);
    GI_LAUNCH_CHECK();
    h->launches += 1;
    return hb_run(h, Lmax, dt, h->L_dev, results, trace_x_host, trace_U_host, true);
}

extern "C" int gi_hmcb_propose_philox(gi_hmcb *h, uint64_t seed, uint64_t counter, double sigma,
                                      const int32_t *L_host, double dt, gi_hmc_result *results) {
    GI_REQUIRE(h && L_host && results, "gi_hmcb_propose_philox: null pointer");
    GI_REQUIRE(h->has_state, "gi_hmcb_propose_philox: call gi_hmcb_set_state first");
    int32_t Lmax = 0;
    int rc = hb_set_L(h, L_host, &Lmax);
    if (rc) return rc;
    const int64_t pairs = ceil_div(h->cfg.M, 2);
    dim3 grid((unsigned)ceil_div(pairs, 256), (unsigned)h->C);
    philox_normal_batched_kernel<<<grid, 256, 0, h->stream>>>(seed, counter, sigma, h->cfg.M, h->cfg.ld,
                                                             h->p, h->st);
    GI_LAUNCH_CHECK();
    h->launches += 1;
    return hb_run(h, Lmax, dt, h->L_dev, results, nullptr, nullptr, true);
}

extern "C" int gi_hmcb_leapfrog_steps(gi_hmcb *h, const double *p0_dev, int32_t nsteps, double dt) {
    GI_REQUIRE(h, "gi_hmcb_leapfrog_steps: null handle");
    GI_REQUIRE(h->has_state, "gi_hmcb_leapfrog_steps: call gi_hmcb_set_state first");
    GI_REQUIRE(nsteps >= 1, "gi_hmcb_leapfrog_steps: nsteps must be >= 1");
    cudaStream_t s = h->stream;
    if (p0_dev)
        GI_CUDA(cudaMemcpyAsync(h->p, p0_dev, sizeof(double) * h->cfg.ld * h->C,
                                cudaMemcpyDeviceToDevice, s));
    else
        GI_CUDA(cudaMemsetAsync(h->p, 0, sizeof(double) * h->cfg.ld * h->C, s));
    // every chain takes nsteps full steps: L = nsteps + 1 is never reached
    int32_t tmp[64];
    for (int c = 0; c < 64; ++c) tmp[c] = nsteps + 1;
    GI_CUDA(cudaMemcpyAsync(h->L_dev, tmp, sizeof(int32_t) * h->C, cudaMemcpyHostToDevice, s));
    GI_CUDA(cudaStreamSynchronize(s));
    return hb_run(h, nsteps, dt, h->L_dev, nullptr, nullptr, nullptr, false);
}

// =============================================================================================
// streaming sampler (see include/gravinv_b200.h)
// =============================================================================================
// Wait for a stream WITHOUT spinning: cudaStreamSynchronize busy-polls by default, which takes a whole
// host core for the ~65 ms a streaming call lasts -- cores the draw workers need (every proposal's
// momentum is 2^20 MT19937 normals generated on the host in the reference's RNG order).
static cudaError_t sync_sleeping(cudaStream_t s) {
    static thread_local cudaEvent_t ev = nullptr;
    static const bool spin = getenv("GI_SLEEPING_SYNC") && getenv("GI_SLEEPING_SYNC")[0] == '0';
    if (spin) return cudaStreamSynchronize(s);
    cudaError_t e = cudaSuccess;
    if (!ev) e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ev, s);
    if (e == cudaSuccess) e = cudaEventSynchronize(ev);
    return e;
}

extern "C" int gi_hmcb_stream_begin(gi_hmcb *h, double dt) {
    GI_REQUIRE(h, "gi_hmcb_stream_begin: null handle");
    GI_REQUIRE(h->has_state, "gi_hmcb_stream_begin: call gi_hmcb_set_state first");
    const size_t bq = sizeof(double) * GI_STREAM_QUEUE_DEPTH * h->C * h->cfg.ld;
    if (!h->qp) {
        GI_CUDA(cudaMalloc(&h->qp, bq));
        GI_CUDA(cudaMemsetAsync(h->qp, 0, bq, h->stream));
    }
    if (!h->rec_dev) {
        h->rec_cap = 4096;
        GI_CUDA(cudaMalloc(&h->rec_dev, sizeof(gi_stream_record) * h->rec_cap));
        GI_CUDA(cudaMallocHost(&h->rec_host, sizeof(gi_stream_record) * h->rec_cap));
    }
    if (!h->copy_stream) {
        GI_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        GI_CUDA(cudaEventCreateWithFlags(&h->ev_commit, cudaEventDisableTiming));
        GI_CUDA(cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming));
    }
    h->copies_pending = false;
    h->adv_pending = false;
    memset(h->cq, 0, sizeof(h->cq));
    h->streaming = true;
    h->stream_dt = dt;
    h->s_xin = h->xa; h->s_xout = h->xb; h->s_mwin = h->mwa; h->s_mwout = h->mwb;
    return GI_OK;
}

static int stream_feed(gi_hmcb *h, int32_t chain, int32_t L, double u, const double *p0,
                       cudaMemcpyKind kind) {
    GI_REQUIRE(h && h->streaming && p0, "gi_hmcb_stream_feed: call gi_hmcb_stream_begin first");
    GI_REQUIRE(chain >= 0 && chain < h->nchains && L >= 1, "gi_hmcb_stream_feed: bad chain or L");
    gi_hmcb::ChainQ &q = h->cq[chain];
    if (q.qn >= GI_STREAM_QUEUE_DEPTH) {
        set_error("gi_hmcb_stream_feed: chain %d already has %d proposals queued", chain, GI_STREAM_QUEUE_DEPTH);
        return GI_ERR_BUSY;
    }
    const int slot = (q.qhead + q.qn) % GI_STREAM_QUEUE_DEPTH;
    // same stream as the kernels: the copy is ordered after the step that consumed this slot
    GI_CUDA(cudaMemcpyAsync(h->qp + ((int64_t)slot * h->C + chain) * h->cfg.ld, p0,
                            sizeof(double) * h->cfg.M, kind, h->stream));
    q.qL[slot] = L;
    q.qu[slot] = u;
    q.qn += 1;
    return GI_OK;
}

extern "C" int gi_hmcb_stream_feed(gi_hmcb *h, int32_t chain, int32_t L, double u,
                                   const double *p0_host) {
    return stream_feed(h, chain, L, u, p0_host, cudaMemcpyHostToDevice);
}

extern "C" int gi_hmcb_stream_feed_dev(gi_hmcb *h, int32_t chain, int32_t L, double u,
                                       const double *p0_dev) {
    return stream_feed(h, chain, L, u, p0_dev, cudaMemcpyDeviceToDevice);
}

extern "C" int gi_hmcb_stream_runway(gi_hmcb *h, int32_t *steps) {
    GI_REQUIRE(h && h->streaming && steps, "gi_hmcb_stream_runway: not streaming");
    // steps until a chain the host still has proposals for would run dry; closed chains are allowed to
    // (they park), so they only count when nothing else is left
    int best = -1, longest = 0;
    for (int c = 0; c < h->nchains; ++c) {
        const gi_hmcb::ChainQ &q = h->cq[c];
        if (q.L_cur == 0 && q.qn == 0) continue;  // parked chain
        int rem = q.L_cur > 0 ? q.L_cur - q.pos : 0;
        for (int k = 0; k < q.qn; ++k) rem += q.qL[(q.qhead + k) % GI_STREAM_QUEUE_DEPTH];
        longest = std::max(longest, rem);
        if (q.closed) continue;
        if (best < 0 || rem < best) best = rem;
    }
    *steps = best < 0 ? longest : best;
    return GI_OK;
}

// the host has fed the last proposal of `chain`: the chain may run dry without ending a call early
extern "C" int gi_hmcb_stream_close_chain(gi_hmcb *h, int32_t chain) {
    GI_REQUIRE(h && h->streaming && chain >= 0 && chain < h->nchains, "gi_hmcb_stream_close_chain: bad argument");
    h->cq[chain].closed = 1;
    return GI_OK;
}

static int launch_update_modes(gi_hmcb *h, const double *grad_in, const double *gdata,
                               const double *x_in, const double *mw_in, double *x_out, double *mw_out,
                               double *grad_out, const signed char *modes, const signed char *slots) {
    gi_plan *p = h->plan;
    UpdateArgs a;
    memset(&a, 0, sizeof(a));
    a.grad_in = grad_in; a.gpart = gdata; a.gparts = 1;
    a.x_in = x_in; a.mw_in = mw_in; a.mwapr = h->mwapr; a.wmsq = h->wmsq; a.low = h->low;
    a.high = h->high; a.p = h->p; a.x_out = x_out; a.mw_out = mw_out; a.grad_out = grad_out;
    a.dt = h->stream_dt; a.M = p->M; a.ld = p->ld; a.reg = h->cfg.reg;
    a.blockpart = p->b_blockpart; a.counter = p->b_counter; a.sums = h->peer ? h->sums_part : h->sums;
    if (gdata) { a.gp_ldk = p->b_gp_ldk; a.gp_piece_stride = p->b_gp_piece_stride; }
    BatchCtl ctl;
    memset(&ctl, 0, sizeof(ctl));
    ctl.vec_stride = p->ld; ctl.nblocks = p->upd_blocks; ctl.use_modes = 1;
    memcpy(ctl.modes, modes, 64);
    if (slots) {
        memcpy(ctl.slots, slots, 64);
        ctl.qp = h->qp;
        ctl.qp_slot_stride = (int64_t)h->C * p->ld;
    }
    const int64_t blocks = peer_slice(p, a, gdata != nullptr);
    if (blocks == 0) return GI_OK;
    dim3 grid((unsigned)blocks, (unsigned)p->b_C);
    update_batched_kernel<<<grid, kUpdThreads, 0, h->stream>>>(a, ctl);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

// queue the kernels of up to nsteps batch steps; the finished proposals' records stay on the device
// until stream_collect
static int stream_enqueue(gi_hmcb *h, int32_t nsteps, int32_t max_records, double *x_host) {
    GI_REQUIRE(h && h->streaming, "gi_hmcb_stream_advance: not streaming");
    GI_REQUIRE(nsteps >= 0 && max_records >= 0, "gi_hmcb_stream_advance: bad argument");
    GI_REQUIRE(!h->adv_pending, "gi_hmcb_stream_advance: the previous call's records were not collected");
    cudaStream_t s = h->stream;
    const int64_t M = h->cfg.M, ld = h->cfg.ld, N = h->cfg.N;
    const int C = (int)h->C;
    const bool logc = h->cfg.reg.constraint == GI_CONSTRAINT_LOGARITHMIC;
    const int cap = (int)std::min<int64_t>(max_records, h->rec_cap);
    int nrec = 0, done = 0;
    while (true) {
        signed char modeA[64], modeB[64], slotB[64];
        StreamCtl sc;
        memset(&sc, 0, sizeof(sc));
        bool any_active = false, any_fin = false, any_start = false;
        int fins = 0;
        for (int c = 0; c < 64; ++c) {
            modeA[c] = 2;  // idle / padding chain: frozen
            modeB[c] = 4;
            slotB[c] = 0;
            if (c >= h->nchains) continue;
            const gi_hmcb::ChainQ &q = h->cq[c];
            if (q.L_cur > 0) {
                any_active = true;
                modeA[c] = (q.pos + 1 < q.L_cur) ? 0 : 1;
                fins += modeA[c] == 1;
            }
        }
        if (any_active && (done >= nsteps || nrec + fins > cap)) break;
        for (int c = 0; c < h->nchains; ++c) {
            const gi_hmcb::ChainQ &q = h->cq[c];
            const bool fin = modeA[c] == 1, idle = q.L_cur == 0;
            if (fin) {
                any_fin = true;
                sc.fin[c] = 1;
                sc.rec[c] = nrec;
                sc.L[c] = q.L_cur;
                h->rec_host[nrec].seq = q.seq;  // host-side fields, merged after the copy back
                h->rec_host[nrec].chain = c;
                nrec += 1;
            }
            if ((fin || idle) && q.qn > 0) {
                any_start = true;
                sc.start[c] = 1;
                sc.u_next[c] = q.qu[q.qhead];
                modeB[c] = 3;
                slotB[c] = (signed char)q.qhead;
            }
        }
        if (!any_active && !any_start) break;  // every chain ran dry
        int rc = GI_OK;
        if (any_active) {
            rc = hb_data_pass(h, h->s_mwin);
            if (!rc) rc = hb_pre_update(h);
            if (!rc)
                rc = launch_update_modes(h, nullptr,
                                         h->peer ? h->pl.stage_local : (h->hook ? h->g_ext : h->gdata), h->s_xin,
                                         h->s_mwin, h->s_xout, h->s_mwout, h->gnew, modeA, nullptr);
            h->launches += 1;
            if (!rc) rc = hb_post_update(h);
            if (rc) return rc;
        }
        if (any_fin || any_start) {
            stream_finish_kernel<<<1, 64, 0, s>>>(h->st, h->sums, h->cfg.reg.alpha, sc, h->rec_dev, 0, C);
            GI_LAUNCH_CHECK();
            h->launches += 1;
        }
        if (any_fin) {
            const int64_t n = std::max(M, N);
            dim3 grid((unsigned)ceil_div(n, 256), (unsigned)C);
            // x_cur is about to change: the positions still being copied out must be gone first
            if (h->copies_pending) GI_CUDA(cudaStreamWaitEvent(s, h->ev_copied, 0));
            commit_stream_kernel<<<grid, 256, 0, s>>>(h->st, sc, M, N, ld, h->s_xin, h->s_mwin, h->gnew,
                                                      h->d, h->x_cur, h->mw_cur, h->g_cur, h->d_cur);
            GI_LAUNCH_CHECK();
            h->launches += 1;
            if (h->stats) {
                StatsMap map;
                memcpy(map.fin, sc.fin, sizeof(map.fin));
                rc = stats_add_chains(h->stats, map, h->st, h->mw_cur, ld, h->nchains, 0,
                                      h->stats_scale, s);
                if (rc) return rc;
                h->launches += 2;
            }
            if (x_host) GI_CUDA(cudaEventRecord(h->ev_commit, s));
        }
        if (any_start) {
            // opening half step of the next trajectory from the (possibly just committed) state
            rc = hb_pre_update(h);
            if (!rc)
                rc = launch_update_modes(h, h->g_cur, nullptr, h->x_cur, h->mw_cur, h->s_xout, h->s_mwout,
                                         nullptr, modeB, slotB);
            if (!rc) rc = hb_post_update(h);
            if (rc) return rc;
            h->launches += 1;
        }
        // peer mode: every rank's slice of the new positions goes to the other ranks (hidden under
        // the next step's forward pass)
        rc = hb_push(h, h->s_xout, h->s_mwout);
        if (rc) return rc;
        if (any_fin && x_host) {
            // the finished chains' positions go to the host on their own stream, under the next steps'
            // contractions -- and, in peer mode, BEHIND this step's slice pushes: both are copy-engine
            // work, and a 40 MB PCIe read queued ahead of the NVLink pushes would hold back every
            // peer's forward pass (measured: 1.5 ms per step at 8 GPUs)
            GI_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_commit, 0));
            if (h->peer && h->push_pending) GI_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_push, 0));
            for (int c = 0; c < h->nchains; ++c)
                if (sc.fin[c])
                    GI_CUDA(cudaMemcpyAsync(x_host + (int64_t)sc.rec[c] * M, h->x_cur + c * ld,
                                            sizeof(double) * M, cudaMemcpyDeviceToHost, h->copy_stream));
            GI_CUDA(cudaEventRecord(h->ev_copied, h->copy_stream));
            h->copies_pending = true;
        }
        for (int c = 0; c < h->nchains; ++c) {
            gi_hmcb::ChainQ &q = h->cq[c];
            if (modeA[c] == 0) q.pos += 1;
            if (sc.fin[c]) { q.L_cur = 0; q.pos = 0; q.seq += 1; }
            if (sc.start[c]) {
                q.L_cur = q.qL[q.qhead];
                q.pos = 0;
                q.qn -= 1;
                q.qhead = (q.qhead + 1) % GI_STREAM_QUEUE_DEPTH;
            }
        }
        std::swap(h->s_xin, h->s_xout);
        if (logc) std::swap(h->s_mwin, h->s_mwout);
        else { h->s_mwin = h->s_xin; h->s_mwout = h->s_xout; }
        if (any_active) done += 1;
    }
    h->adv_pending = true;
    h->adv_nrec = nrec;
    h->adv_done = done;
    return GI_OK;
}

// wait for the queued steps and hand out their records
static int stream_collect(gi_hmcb *h, gi_stream_record *records, int32_t max_records, int32_t *nrecords,
                          int32_t *steps_done) {
    GI_REQUIRE(h && h->streaming && nrecords && h->adv_pending, "gi_hmcb_stream_advance_end: nothing pending");
    cudaStream_t s = h->stream;
    const int nrec = h->adv_nrec, done = h->adv_done;
    GI_REQUIRE(nrec <= max_records && (records || nrec == 0), "gi_hmcb_stream_advance_end: record buffer too small");
    h->adv_pending = false;
    if (nrec > 0) {
        // device fields (accept, U, H) come back through a staging copy, host fields are kept
        gi_stream_record *tmp = new gi_stream_record[nrec];
        cudaError_t e = cudaMemcpyAsync(tmp, h->rec_dev, sizeof(gi_stream_record) * nrec,
                                        cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = sync_sleeping(s);
        if (e != cudaSuccess) { delete[] tmp; return cuda_fail(e, "records", __FILE__, __LINE__); }
        for (int i = 0; i < nrec; ++i) {
            records[i] = tmp[i];
            records[i].seq = h->rec_host[i].seq;
            records[i].chain = h->rec_host[i].chain;
        }
        delete[] tmp;
    } else {
        GI_CUDA(sync_sleeping(s));
    }
    if (h->copies_pending) {
        GI_CUDA(sync_sleeping(h->copy_stream));
        h->copies_pending = false;
    }
    *nrecords = nrec;
    if (steps_done) *steps_done = done;
    return GI_OK;
}

extern "C" int gi_hmcb_stream_advance(gi_hmcb *h, int32_t nsteps, gi_stream_record *records,
                                      int32_t max_records, int32_t *nrecords, int32_t *steps_done,
                                      double *x_host) {
    GI_REQUIRE(nrecords && (records || max_records == 0), "gi_hmcb_stream_advance: bad argument");
    int rc = stream_enqueue(h, nsteps, max_records, x_host);
    if (rc) return rc;
    return stream_collect(h, records, max_records, nrecords, steps_done);
}

extern "C" int gi_hmcb_stream_advance_begin(gi_hmcb *h, int32_t nsteps, int32_t max_records, double *x_host) {
    return stream_enqueue(h, nsteps, max_records, x_host);
}

extern "C" int gi_hmcb_stream_advance_end(gi_hmcb *h, gi_stream_record *records, int32_t max_records,
                                          int32_t *nrecords, int32_t *steps_done) {
    return stream_collect(h, records, max_records, nrecords, steps_done);
}

// drop the proposals queued for `chain` that have not started (the chain has all the samples it needs)
extern "C" int gi_hmcb_stream_cancel(gi_hmcb *h, int32_t chain, int32_t *dropped) {
    GI_REQUIRE(h && h->streaming && chain >= 0 && chain < h->nchains, "gi_hmcb_stream_cancel: bad argument");
    GI_REQUIRE(!h->adv_pending, "gi_hmcb_stream_cancel: collect the pending records first");
    if (dropped) *dropped = h->cq[chain].qn;
    h->cq[chain].qn = 0;
    return GI_OK;
}

extern "C" int gi_hmcb_stream_queue_space(gi_hmcb *h, int32_t chain, int32_t *space) {
    GI_REQUIRE(h && h->streaming && space && chain >= 0 && chain < h->nchains,
               "gi_hmcb_stream_queue_space: bad argument");
    *space = GI_STREAM_QUEUE_DEPTH - h->cq[chain].qn;
    return GI_OK;
}

extern "C" int64_t gi_hmcb_launch_count(const gi_hmcb *h) { return h ? h->launches : 0; }
extern "C" int32_t gi_hmcb_padded_chains(const gi_hmcb *h) { return h ? h->C : 0; }
