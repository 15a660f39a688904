// fused.cu -- ONE pass over the weighted kernel per gradient evaluation of a single chain:
//     d = Aw x          (inversion/potential.py:698  dpre = np.dot(self.Aw, mw))
//     g = Aw^T r        (potential.py:708            2 * np.dot(self.Aw.T, r), factor 2 applied later)
//     r = (d + fix - mean(d + fix)) - dobs_c         (potential.py:699-706)
// The two GEMV passes of gemv_fwd_kernel / gemv_adj_kernel each stream Aw from HBM (2 x 8 N M bytes
// per evaluation).  r depends on ALL of d through the mean, but linearly:  Aw^T r = Aw^T e - mean * s
// with e = d + fix - dobs_c (row-local) and s = Aw^T 1 (precomputed once), so row i's contribution to
// the adjoint can be added as soon as d_i is complete -- while the row is still ON CHIP.
//
// Persistent kernel, one CTA per SM (cooperative launch => co-resident).  CTA b owns the column strip
// [b W, (b+1) W) for every row: its slice of x and of the adjoint accumulator live in registers.
// Rows arrive in a 4-slot shared-memory ring filled by 1-D TMA bulk copies
// (cp.async.bulk ... mbarrier::complete_tx), two to three rows (56 KB each at M = 2^20) in flight per SM.
//   iteration i, row i     : partial dot of the strip with x (four accumulator chains), published as two
//                            self-validating 8-byte words {value bits 0..31 | tag, bits 32..63 | tag}
//                            (no fence, no flag; each word is single-copy atomic, both must carry the tag);
//   iteration i, row i - 2 : every CTA has polled the 148 pairs of that row (the poll is issued an
//                            iteration early, its L2 round trip hides under the arithmetic), sums them in
//                            a fixed order (identical bits on every CTA), forms e and adds e * row-strip
//                            -- held in REGISTERS since iteration i - 1 -- to its accumulator;
//   iteration i, row i - 1 : its strip moves from shared memory into 28 registers per thread and the
//                            slot goes straight back to the TMA unit (row i + 3).
// Synchronisation inside the CTA: one full barrier per iteration; the hand-offs to thread 0 (publish,
// re-arm the TMA) are named barriers on which the other seven warps only arrive.
// No atomics (a same-address counter would serialise 148 L2 atomics per row), no grid-wide barrier:
// a CTA only ever waits for rows published two iterations earlier.  Measured on B200 (4096 x 2^20,
// two-pass kernels 9.5 ms): round 1 (rows waited in the ring, two full barriers, one dot chain)
// 6.8 ms = 0.77 of the measured copy bandwidth on DRAM bytes; retained row in registers 6.5 ms;
// + arrive/sync hand-offs, four dot chains, prefetched fix / dobs 5.41 ms = 6.35 TB/s = 0.97 (c5, 16 384
// rows: 23.6 ms = 0.89).  Tried and slower (round 1): release/acquire flags instead of tagged words
// (14.7 ms: a MEMBAR per row on the critical path), consuming a row one iteration after its publication
// (10.9 ms: every poll misses), two CTAs per SM with half strips (10.3 ms: the hand-off traffic grows
// with the square of the CTA count), a dedicated TMA producer warp with mbarrier slot release (10.5 ms),
// publishing at the end of the iteration (12.3 ms).  DRAM traffic per evaluation is 8 N M bytes instead
// of 16 N M (ncu: 34.38 GB read at 4096 x 2^20); the result is deterministic (fixed strip ownership and
// summation order).  Rounding differs from the two-pass form at the 1e-16 * |mean s| / |Aw^T r| level
// (~1e-14), far inside the 1e-9 parity bar of the trajectories.
#include <algorithm>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace gi {
namespace {

constexpr int kFT = 256;    // threads per CTA
constexpr int kFS = 4;      // ring slots
constexpr int kFLag = 2;    // rows between publishing a partial and consuming the complete row
constexpr int kFU = 7;      // double4 chunks per thread at most -> strip width <= 7168 columns
constexpr int kFCtas = 1;   // CTAs per SM (2 with half strips was slower: the hand-off costs ~ P^2)
constexpr int kFQ = 1;      // partials polled per thread -> up to 256 CTAs
constexpr unsigned long long kSpinLimit = 1ull << 26;

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void lds2(unsigned a, double &x0, double &x1) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(a));
}

struct FusedArgs {
    const double *G;
    int64_t ld, nrows, W;
    const double *x, *dobs_c, *fix, *s;
    double inv_n;
    unsigned long long *part;     // [nrows][P][2] {value lo32 | tag << 32, value hi32 | tag << 32}, tag = epoch_base + row + 1 (mod 2^32)
    unsigned long long epoch_base;
    double *mean_io;  // [2] mean of the previous / this evaluation (ping-pong by launch parity)
    int mean_slot;
    int center;       // 1: r = e - mean(e) (potential.py:706); 0: r = e (reginv.py:256, no mean removal)
    double *d_out, *g_out;
};

__global__ void __launch_bounds__(kFT, 1) fused_pass_kernel(FusedArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int P = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t W = a.W, c0 = (int64_t)b * W;
    const int64_t rem = a.ld - c0;
    const int64_t Wb = rem <= 0 ? 0 : (rem < W ? rem : W);
    const unsigned bytes = (unsigned)(Wb * 8);
    double *ring = reinterpret_cast<double *>(smem);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem + (size_t)kFS * W * 8);
    double *scratch = reinterpret_cast<double *>(full + kFS);
    const unsigned ring_a = smem_addr(ring), full_a = smem_addr(full);

    // this thread's slice of x and of the adjoint accumulator.  Group u = 1024 strip columns; the
    // thread owns doubles {2 tid, 2 tid + 1} of the group's first 512 and of its second 512, so a
    // warp's 16-byte shared-memory loads cover 512 contiguous bytes (conflict free; 32 contiguous
    // bytes per thread would be 2-way bank conflicts -- 44 % of the wavefronts in the ncu capture)
    double xv[kFU][4], ga[kFU][4], rr[kFU][4];  // rr: the strip of the row whose d_j arrives next iteration
#pragma unroll
    for (int u = 0; u < kFU; ++u) {
        const int64_t o0 = 1024LL * u + 2 * tid, o1 = o0 + 512;
        rr[u][0] = rr[u][1] = rr[u][2] = rr[u][3] = 0.0;
        xv[u][0] = xv[u][1] = xv[u][2] = xv[u][3] = 0.0;
        if (o0 < Wb) { xv[u][0] = __ldg(a.x + c0 + o0); xv[u][1] = __ldg(a.x + c0 + o0 + 1); }
        if (o1 < Wb) { xv[u][2] = __ldg(a.x + c0 + o1); xv[u][3] = __ldg(a.x + c0 + o1 + 1); }
        ga[u][0] = ga[u][1] = ga[u][2] = ga[u][3] = 0.0;
    }
    if (tid == 0) {
        for (int s = 0; s < kFS; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full_a + 8 * s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int64_t row) {  // thread 0: fill slot row % kFS with the strip of `row`
        const int slot = (int)(row % kFS);
        const unsigned bar = full_a + 8 * slot;
        if (bytes) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    ring_a + (unsigned)(slot * W * 8)),
                "l"(a.G + row * a.ld + c0), "r"(bytes), "r"(bar)
                : "memory");
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        }
    };
    if (tid == 0)
        for (int64_t r = 0; r < kFS && r < a.nrows; ++r) issue(r);

    // Partials travel as TWO self-validating 8-byte words, {value bits 0..31 | tag << 32} and
    // {value bits 32..63 | tag << 32} (the protocol of NCCL's LL): PTX guarantees single-copy atomicity
    // for an aligned 8-byte scalar -- and a vector access is a set of such scalars, with no ordering
    // or 16-byte atomicity between them -- so each word carries its own 32-bit tag and a value is
    // accepted only when BOTH words show the tag of the row.  No fence, no separate flag.  A slot's
    // previous occupant is always the previous launch's word for the same row (every launch rewrites
    // every slot), whose tag differs by nrows + 1 (mod 2^32) != 0, so a stale word can never pass.
    // The poll for row j is issued one iteration early and only CHECKED when the row is consumed, so
    // its L2 round trip hides under the arithmetic of the rows in between.
    auto poll = [&](int64_t row, unsigned long long &w0, unsigned long long &w1) {
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];"
                     : "=l"(w0), "=l"(w1)
                     : "l"(a.part + 2 * (row * P + tid))
                     : "memory");
    };
    // Iteration i: forward dot of row i (published), then the adjoint update with row j = i - 2.
    // Synchronisation inside the CTA (round 2): ONE full barrier per iteration.  The two hand-offs to
    // thread 0 -- "all warps' dot partials are in scratch" (publish) and "all warps are done with the
    // slot" (refill) -- are named barriers on which the other seven warps only ARRIVE, so they run on
    // into the poll check / the next row while warp 0 publishes or re-arms the TMA.
    const double m0 = a.center ? a.mean_io[a.mean_slot] : 0.0;  // last evaluation's mean: e is formed around it, so
    double sd = 0.0;                           // the correction mean' * s below stays small
    unsigned long long pw0 = 0, pw1 = 0;       // prefetched words of the row consumed NEXT iteration
    double *sA = scratch, *sV = scratch + 8;   // sA[8]: dot partials; sV[2][8]: d_j partials (by parity)
    // strip bounds of this thread as group counts: o0 = 1024 u + 2 tid < Wb  <=>  u < n0 (second half: n1)
    const int n0 = (int)max((int64_t)0, min((int64_t)kFU, (Wb - 2 * tid + 1023) / 1024));
    const int n1 = (int)max((int64_t)0, min((int64_t)kFU, (Wb - 512 - 2 * tid + 1023) / 1024));
    double nfix = 0.0, ndobs = 0.0;            // fix[j], dobs_c[j] of the row consumed NEXT iteration
    if (a.nrows > 0 && kFLag == 1) { nfix = a.fix ? __ldg(a.fix) : 0.0; ndobs = __ldg(a.dobs_c); }
    for (int64_t i = 0; i < a.nrows + kFLag; ++i) {
        const int64_t j = i - kFLag;
        unsigned long long cw0 = pw0, cw1 = pw1;  // words of row j, polled during iteration i - 1
        if (tid < P && j + 1 >= 0 && j + 1 < a.nrows) poll(j + 1, pw0, pw1);
        const double cfix = nfix, cdobs = ndobs;
        if (j + 1 >= 0 && j + 1 < a.nrows) {      // off the critical path: an L2 round trip per row otherwise
            nfix = a.fix ? __ldg(a.fix + j + 1) : 0.0;
            ndobs = __ldg(a.dobs_c + j + 1);
        }
        if (i < a.nrows) {
            // ---- phase 1: partial dot product of row i with this CTA's slice of x ------------
            const int slot = (int)(i % kFS);
            const unsigned bar = full_a + 8 * slot, parity = (unsigned)((i / kFS) & 1);
            unsigned done = 0;
            while (!done)
                asm volatile(
                    "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                    : "=r"(done)
                    : "r"(bar), "r"(parity)
                    : "memory");
            const unsigned row_a = ring_a + (unsigned)(slot * W * 8) + 16u * tid;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;  // four chains: the DFMA latency is not serialised 28 deep
#pragma unroll
            for (int u = 0; u < kFU; ++u) {
                double g0, g1;
                if (u < n0) {
                    lds2(row_a + 8192u * u, g0, g1);
                    a0 = fma(g0, xv[u][0], a0);
                    a1 = fma(g1, xv[u][1], a1);
                }
                if (u < n1) {
                    lds2(row_a + 8192u * u + 4096u, g0, g1);
                    a2 = fma(g0, xv[u][2], a2);
                    a3 = fma(g1, xv[u][3], a3);
                }
            }
            const double acc = warp_sum((a0 + a1) + (a2 + a3));
            if (lane == 0) sA[warp] = acc;
            if (warp == 0) {
                asm volatile("barrier.cta.sync 1, %0;" ::"n"(kFT) : "memory");
                if (tid == 0) {
                    double t = 0.0;
                    for (int w = 0; w < kFT / 32; ++w) t += sA[w];
                    const unsigned long long tag = (a.epoch_base + (unsigned long long)i + 1ull) << 32;
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(t);
                    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(a.part + 2 * (i * P + b)),
                                 "l"((bits & 0xffffffffull) | tag), "l"((bits >> 32) | tag)
                                 : "memory");
                }
            } else {
                asm volatile("barrier.cta.arrive 1, %0;" ::"n"(kFT) : "memory");
            }
        }
        double *sv = sV + 8 * (int)(i & 1);
        if (j >= 0) {
            double v = 0.0;
            if (tid < P) {
                // the complete row j: this thread's share is CTA `tid`'s partial
                const unsigned long long target = (a.epoch_base + (unsigned long long)j + 1ull) & 0xffffffffull;
                unsigned long long spins = 0;
                while ((cw0 >> 32) != target || (cw1 >> 32) != target) {
                    poll(j, cw0, cw1);
                    if (++spins > kSpinLimit) __trap();
                }
                v = __longlong_as_double((long long)((cw0 & 0xffffffffull) | (cw1 << 32)));
            }
            v = warp_sum(v);  // d_j in a fixed order (identical bits on every CTA)
            if (lane == 0) sv[warp] = v;
        }
        __syncthreads();  // the one full barrier: d_j's partials are in, last iteration's reads of sv' are done
        if (j >= 0) {
            // ---- phase 2: add e_j * (row j strip, held in registers since last iteration) -------
            double dj = 0.0;
            for (int w = 0; w < kFT / 32; ++w) dj += sv[w];
            const double dinv = (a.fix ? dj + cfix : dj) - m0;
            const double ej = dinv - cdobs;
            sd += dinv;
#pragma unroll
            for (int u = 0; u < kFU; ++u) {
                ga[u][0] = fma(rr[u][0], ej, ga[u][0]);
                ga[u][1] = fma(rr[u][1], ej, ga[u][1]);
                ga[u][2] = fma(rr[u][2], ej, ga[u][2]);
                ga[u][3] = fma(rr[u][3], ej, ga[u][3]);
            }
            if (tid == 0 && b == (int)(j % P)) a.d_out[j] = dj;
        }
        const int64_t k = i - 1;  // the row whose forward dot ran last iteration: shared memory -> registers
        if (k >= 0 && k < a.nrows) {
            const unsigned row_a = ring_a + (unsigned)((int)(k % kFS) * W * 8) + 16u * tid;
#pragma unroll
            for (int u = 0; u < kFU; ++u) {
                if (u < n0) lds2(row_a + 8192u * u, rr[u][0], rr[u][1]);
                if (u < n1) lds2(row_a + 8192u * u + 4096u, rr[u][2], rr[u][3]);
            }
            // every warp is done with slot k % kFS: warp 0 hands it back to the TMA unit (one iteration
            // after the row's dot), the others only arrive
            if (warp == 0) {
                asm volatile("barrier.cta.sync 2, %0;" ::"n"(kFT) : "memory");
                if (tid == 0 && k + kFS < a.nrows) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(k + kFS);
                }
            } else {
                asm volatile("barrier.cta.arrive 2, %0;" ::"n"(kFT) : "memory");
            }
        }
    }
    // g = Aw^T e - mean' * s   (r = e - mean', mean' = mean - m0)
    const double mean = a.center ? sd * a.inv_n : 0.0;
    if (b == 0 && tid == 0 && a.center) a.mean_io[a.mean_slot ^ 1] = m0 + mean;
#pragma unroll
    for (int u = 0; u < kFU; ++u) {
        const int64_t o0 = 1024LL * u + 2 * tid, o1 = o0 + 512;
        if (o0 < Wb) {
            a.g_out[c0 + o0] = ga[u][0] - mean * __ldg(a.s + c0 + o0);
            a.g_out[c0 + o0 + 1] = ga[u][1] - mean * __ldg(a.s + c0 + o0 + 1);
        }
        if (o1 < Wb) {
            a.g_out[c0 + o1] = ga[u][2] - mean * __ldg(a.s + c0 + o1);
            a.g_out[c0 + o1 + 1] = ga[u][3] - mean * __ldg(a.s + c0 + o1 + 1);
        }
    }
}

__global__ void fill_kernel(double *p, int64_t n, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace
}  // namespace gi

using namespace gi;

struct gi_fused {
    int64_t nrows, M, ld, W;
    int P;
    size_t smem;
    const double *G;
    double *s, *mean_io;
    unsigned long long *part, epoch, nlaunch;
    int64_t launches;
};

extern "C" int gi_fused_destroy(gi_fused *f) {
    if (!f) return GI_OK;
    cudaFree(f->s);
    cudaFree(f->part);
    cudaFree(f->mean_io);
    delete f;
    return GI_OK;
}

extern "C" int gi_fused_create(int64_t nrows, int64_t M, int64_t ld, const double *G_dev, void *stream,
                               gi_fused **out) {
    GI_REQUIRE(out && G_dev && nrows > 0 && M > 0 && ld >= M && ld % 4 == 0, "gi_fused_create: bad argument");
    int dev = 0, coop = 0, max_smem = 0;
    GI_CUDA(cudaGetDevice(&dev));
    GI_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    GI_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int P = kFCtas * sm_count();
    const int64_t W = ceil_div(ceil_div(ld, (int64_t)P), 4) * 4;
    const size_t smem = (size_t)kFS * W * 8 + kFS * 8 + 32 * 8;
    GI_REQUIRE(coop, "gi_fused_create: the device does not support cooperative launches");
    GI_REQUIRE(P <= kFT * kFQ, "gi_fused_create: more CTAs than partial slots per thread");
    GI_REQUIRE(W <= 4LL * kFT * kFU && W % 2 == 0 && smem <= (size_t)max_smem,
               "gi_fused_create: the column strip of one SM (%lld columns) does not fit its shared memory",
               (long long)W);
    GI_CUDA(cudaFuncSetAttribute(fused_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_pass_kernel, kFT, smem));
    GI_REQUIRE(per_sm >= kFCtas, "gi_fused_create: the CTAs of one SM do not fit it together");
    gi_fused *f = new gi_fused();
    memset(f, 0, sizeof(*f));
    f->nrows = nrows; f->M = M; f->ld = ld; f->W = W; f->P = P; f->smem = smem; f->G = G_dev;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMalloc(&f->s, sizeof(double) * ld);
    if (e == cudaSuccess) e = cudaMalloc(&f->part, 16 * (size_t)nrows * P);
    if (e == cudaSuccess) e = cudaMemsetAsync(f->part, 0, 16 * (size_t)nrows * P, st);
    if (e == cudaSuccess) e = cudaMalloc(&f->mean_io, 2 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemsetAsync(f->mean_io, 0, 2 * sizeof(double), st);
    if (e != cudaSuccess) {
        gi_fused_destroy(f);
        return cuda_fail(e, "gi_fused_create", __FILE__, __LINE__);
    }
    // s = Aw^T 1 through the deterministic adjoint pass
    gi_plan *plan = nullptr;
    double *ones = nullptr;
    int rc = gi_plan_create(nrows, M, ld, 1, &plan);
    if (!rc) {
        e = cudaMalloc(&ones, sizeof(double) * nrows);
        if (e != cudaSuccess) rc = cuda_fail(e, "gi_fused_create", __FILE__, __LINE__);
    }
    if (!rc) {
        fill_kernel<<<(unsigned)ceil_div(nrows, 256), 256, 0, st>>>(ones, nrows, 1.0);
        rc = gi_gemv_adj(plan, G_dev, ones, f->s, st);
    }
    if (!rc) {
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = cuda_fail(e, "gi_fused_create", __FILE__, __LINE__);
    }
    cudaFree(ones);
    gi_plan_destroy(plan);
    if (rc) {
        gi_fused_destroy(f);
        return rc;
    }
    *out = f;
    return GI_OK;
}

extern "C" int gi_fused_pass(gi_fused *f, const double *x_dev, const double *dobs_c_dev, const double *fix_dev,
                             int32_t center, double *d_dev, double *g_dev, void *stream) {
    GI_REQUIRE(f && x_dev && dobs_c_dev && d_dev && g_dev, "gi_fused_pass: null pointer");
    FusedArgs a;
    a.G = f->G; a.ld = f->ld; a.nrows = f->nrows; a.W = f->W;
    a.x = x_dev; a.dobs_c = dobs_c_dev; a.fix = fix_dev; a.s = f->s;
    a.inv_n = 1.0 / (double)f->nrows;
    a.part = f->part; a.epoch_base = f->epoch;
    a.mean_io = f->mean_io; a.mean_slot = (int)(f->nlaunch & 1ull);
    a.center = center ? 1 : 0;
    if (center) f->nlaunch += 1;
    a.d_out = d_dev; a.g_out = g_dev;
    f->epoch += (unsigned long long)f->nrows + 1ull;
    void *args[] = {&a};
    GI_CUDA(cudaLaunchCooperativeKernel((void *)fused_pass_kernel, dim3((unsigned)f->P), dim3(kFT), args, f->smem,
                                        (cudaStream_t)stream));
    f->launches += 1;
    return GI_OK;
}
