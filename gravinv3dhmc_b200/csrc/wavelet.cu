// wavelet.cu -- wavelet-compressed forward path: level-2 db4 'periodization' DWT (1-D and 3-D,
// packed like pywt.coeffs_to_array) and the CSR matvec Awcp @ coef.
//
// Replaces gravmag/compressor1D.py:17-60 and gravmag/compressor3D.py:17-68, whose arithmetic lives
// in the third-party PyWavelets package (pywt.wavedec / wavedecn / coeffs_to_array; no version is
// pinned by the reference and the package is absent from this image -> parity is pinned only by
// self-consistency).  Conventions restated from pywt:
//   * filters: db4 dec_lo / dec_hi (quadrature mirror), F = 8 taps
//   * mode 'periodization': odd lengths are extended by repeating the last sample (ne = n + 1),
//     output length ne/2, cA[o] = sum_j dec_lo[j] * xe[(2o + F/2 - j) mod ne]
//   * wavedec -> [cA2, cD2, cD1] concatenated; wavedecn transforms axes 0,1,2 in order per level;
//     coeffs_to_array packs Mallat style: level-2 block at the origin, level-1 detail blocks
//     offset by 2*h2 along every 'd' axis (gaps stay zero when shapes do not nest).
#include <stdlib.h>

#include "common.cuh"

namespace gi {

__device__ __constant__ double kDb4Lo[8] = {
    -0.010597401784997278, 0.032883011666982945, 0.030841381835986965, -0.18703481171888114,
    -0.02798376941698385,  0.6308807679295904,   0.7148465705525415,   0.23037781330885523};
__device__ __constant__ double kDb4Hi[8] = {
    -0.23037781330885523, 0.7148465705525415,  -0.6308807679295904,  -0.02798376941698385,
    0.18703481171888114,  0.030841381835986965, -0.032883011666982945, -0.010597401784997278};

// One output pair (approximation, detail) of a single-level DWT along `axis`: element t of the
// (m0,m1,m2) index space, m_axis = no.  Shared by the per-axis kernel and the fused small-volume
// kernel, so both produce the same bits.
__device__ __forceinline__ void dwt_axis_point(const double *src, int d1, int d2, int n0, int n1, int n2,
                                               int axis, double *dst, int e1, int e2, int64_t t) {
    const int n[3] = {n0, n1, n2};
    const int len = n[axis];
    const int ne = len + (len & 1);
    const int no = ne / 2;
    int m[3] = {n0, n1, n2};
    m[axis] = no;
    int idx[3];
    idx[2] = (int)(t % m[2]);
    idx[1] = (int)((t / m[2]) % m[1]);
    idx[0] = (int)(t / ((int64_t)m[2] * m[1]));
    const int o = idx[axis];
    const int64_t istr[3] = {(int64_t)d1 * d2, d2, 1};
    const int64_t ostr[3] = {(int64_t)e1 * e2, e2, 1};
    int64_t base = 0, obase = 0;
    for (int a = 0; a < 3; ++a)
        if (a != axis) {
            base += idx[a] * istr[a];
            obase += idx[a] * ostr[a];
        }
    double ca = 0.0, cd = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int s = (2 * o + 4 - j) % ne;
        if (s < 0) s += ne;
        if (s >= len) s = len - 1;  // odd length: the appended sample repeats the last one
        const double v = src[base + s * istr[axis]];
        ca += kDb4Lo[j] * v;
        cd += kDb4Hi[j] * v;
    }
    dst[obase + o * ostr[axis]] = ca;
    dst[obase + (o + no) * ostr[axis]] = cd;
}

// One single-level DWT along `axis` of a batch of C-ordered (d0,d1,d2) volumes.
// in : [batch][d0][d1][d2] with batch stride in_bs; only the sub-box (n0,n1,n2) is read, with
//      row strides taken from the full dims (d0,d1,d2)
// out: [batch][e0][e1][e2] where e_axis = 2*no (approximation then detail) and e_other = n_other
__global__ void dwt_axis_kernel(const double *__restrict__ in, int64_t in_bs, int d1, int d2, int n0,
                                int n1, int n2, int axis, double *__restrict__ out, int64_t out_bs,
                                int e1, int e2) {
    int m[3] = {n0, n1, n2};
    m[axis] = (m[axis] + 1) / 2;
    const int64_t total = (int64_t)m[0] * m[1] * m[2];
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    dwt_axis_point(in + (int64_t)blockIdx.y * in_bs, d1, d2, n0, n1, n2, axis, out + (int64_t)blockIdx.y * out_bs,
                   e1, e2, t);
}

// The whole level-2 3-D transform of ONE small volume in one CTA (round 2): the six axis passes run
// between shared-memory buffers with a barrier in between and the Mallat packing writes the result,
// instead of 6 kernels + memset + pack + a stream-ordered allocation (ten stream operations, 69 us on
// the 6000-voxel example grid -- more than the CSR matvec they feed).  Same arithmetic per coefficient
// (dwt_axis_point), so the same bits as the per-axis path.
// shared: A[t1sz] | B[t1sz] | C[t1sz / 2], t1sz = 8 h1_0 h1_1 h1_2
__global__ void __launch_bounds__(1024)
dwt3d_l2_small_kernel(const double *__restrict__ x, int nz, int ny, int nx, double *__restrict__ out) {
    extern __shared__ double sm[];
    const int h1[3] = {(nz + 1) / 2, (ny + 1) / 2, (nx + 1) / 2};
    const int h2[3] = {(h1[0] + 1) / 2, (h1[1] + 1) / 2, (h1[2] + 1) / 2};
    const int f[3] = {2 * h2[0] + h1[0], 2 * h2[1] + h1[1], 2 * h2[2] + h1[2]};
    const int64_t t1sz = 8LL * h1[0] * h1[1] * h1[2];
    double *A = sm, *B = sm + t1sz, *Cb = sm + 2 * t1sz;
    auto pass = [&](const double *src, int sd1, int sd2, int a0, int a1, int a2, int axis, double *dst, int e1,
                    int e2) {
        int m[3] = {a0, a1, a2};
        m[axis] = (m[axis] + 1) / 2;
        const int64_t total = (int64_t)m[0] * m[1] * m[2];
        for (int64_t t = threadIdx.x; t < total; t += blockDim.x)
            dwt_axis_point(src, sd1, sd2, a0, a1, a2, axis, dst, e1, e2, t);
        __syncthreads();
    };
    const int o0 = 2 * h1[0], o1 = 2 * h1[1], o2 = 2 * h1[2];
    // level 1 (same buffer dims as dwt3_level): x -> A (o0,ny,nx) -> B (o0,o1,nx) -> A = T1 (o0,o1,o2)
    pass(x, ny, nx, nz, ny, nx, 0, A, ny, nx);
    pass(A, ny, nx, o0, ny, nx, 1, B, o1, nx);
    pass(B, o1, nx, o0, o1, nx, 2, A, o1, o2);
    // level 2 on the 'aaa' corner of T1 (sub-box h1 of dims 2 h1): -> B (p0,h1_1,h1_2) -> C (p0,p1,h1_2) -> B = T2
    const int p0 = 2 * h2[0], p1 = 2 * h2[1], p2 = 2 * h2[2];
    pass(A, o1, o2, h1[0], h1[1], h1[2], 0, B, h1[1], h1[2]);
    pass(B, h1[1], h1[2], p0, h1[1], h1[2], 1, Cb, p1, h1[2]);
    pass(Cb, p1, h1[2], p0, p1, h1[2], 2, B, p1, p2);
    // pack like pack3d_kernel: zeros, the level-2 block at the origin, the level-1 detail blocks
    const int64_t fsz = (int64_t)f[0] * f[1] * f[2];
    for (int64_t t = threadIdx.x; t < fsz; t += blockDim.x) out[t] = 0.0;
    __syncthreads();
    const int64_t n2 = (int64_t)p0 * p1 * p2;
    for (int64_t t = threadIdx.x; t < n2; t += blockDim.x) {
        const int i2 = (int)(t % p2), i1 = (int)((t / p2) % p1), i0 = (int)(t / ((int64_t)p2 * p1));
        out[((int64_t)i0 * f[1] + i1) * f[2] + i2] = B[t];
    }
    for (int64_t q = threadIdx.x; q < t1sz; q += blockDim.x) {
        const int i[3] = {(int)(q / ((int64_t)o2 * o1)), (int)((q / o2) % o1), (int)(q % o2)};
        if (i[0] < h1[0] && i[1] < h1[1] && i[2] < h1[2]) continue;  // 'aaa' went on to level 2
        int o[3];
        for (int a = 0; a < 3; ++a) o[a] = (i[a] < h1[a]) ? i[a] : 2 * h2[a] + (i[a] - h1[a]);
        out[((int64_t)o[0] * f[1] + o[1]) * f[2] + o[2]] = A[q];
    }
}

// the level-2 1-D transform of one short vector in one CTA: x -> T1 = [a1 | d1] (shared) -> out = [a2 | d2 | d1]
__global__ void __launch_bounds__(1024) dwt1d_l2_small_kernel(const double *__restrict__ x, int n, double *__restrict__ out) {
    extern __shared__ double sm[];
    const int h1 = (n + 1) / 2, h2 = (h1 + 1) / 2;
    for (int64_t t = threadIdx.x; t < h1; t += blockDim.x) dwt_axis_point(x, 1, n, 1, 1, n, 2, sm, 1, 2 * h1, t);
    __syncthreads();
    for (int64_t t = threadIdx.x; t < h2; t += blockDim.x) dwt_axis_point(sm, 1, 2 * h1, 1, 1, h1, 2, out, 1, 2 * h2, t);
    for (int t = threadIdx.x; t < h1; t += blockDim.x) out[2 * h2 + t] = sm[h1 + t];
}

// pack level-1 result T1 (dims 2*h1) and level-2 result T2 (dims 2*h2) into F (dims 2*h2 + h1)
__global__ void pack3d_kernel(const double *__restrict__ T1, const double *__restrict__ T2, int h1_0,
                              int h1_1, int h1_2, int h2_0, int h2_1, int h2_2,
                              double *__restrict__ F, int64_t t1_bs, int64_t t2_bs, int64_t f_bs) {
    const int h1[3] = {h1_0, h1_1, h1_2}, h2[3] = {h2_0, h2_1, h2_2};
    const int f[3] = {2 * h2_0 + h1_0, 2 * h2_1 + h1_1, 2 * h2_2 + h1_2};
    const int64_t n1 = 8LL * h1_0 * h1_1 * h1_2, n2 = 8LL * h2_0 * h2_1 * h2_2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double *t1 = T1 + (int64_t)blockIdx.y * t1_bs;
    const double *t2 = T2 + (int64_t)blockIdx.y * t2_bs;
    double *out = F + (int64_t)blockIdx.y * f_bs;
    if (t < n2) {
        const int e[3] = {2 * h2_0, 2 * h2_1, 2 * h2_2};
        const int i2 = (int)(t % e[2]), i1 = (int)((t / e[2]) % e[1]), i0 = (int)(t / ((int64_t)e[2] * e[1]));
        out[((int64_t)i0 * f[1] + i1) * f[2] + i2] = t2[t];
    } else if (t < n2 + n1) {
        const int64_t q = t - n2;
        const int e[3] = {2 * h1_0, 2 * h1_1, 2 * h1_2};
        const int i[3] = {(int)(q / ((int64_t)e[2] * e[1])), (int)((q / e[2]) % e[1]), (int)(q % e[2])};
        if (i[0] < h1[0] && i[1] < h1[1] && i[2] < h1[2]) return;  // 'aaa' went on to level 2
        int o[3];
        for (int a = 0; a < 3; ++a) o[a] = (i[a] < h1[a]) ? i[a] : 2 * h2[a] + (i[a] - h1[a]);
        out[((int64_t)o[0] * f[1] + o[1]) * f[2] + o[2]] = t1[q];
    }
}

// y[row] = sum_k data[k] * x[indices[k]]  -- one warp per row, fixed summation tree
__global__ void csr_spmv_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                const double *__restrict__ data, int64_t nrows,
                                const double *__restrict__ x, double *__restrict__ y) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    double acc = 0.0;
    for (int64_t k = indptr[row] + lane; k < indptr[row + 1]; k += 32)
        acc = fma(data[k], __ldg(x + indices[k]), acc);
    acc = warp_sum(acc);
    if (lane == 0) y[row] = acc;
}

// ---- dense coefficient rows -> CSR (compressor*.kernelcompressor: zero |c| < thr, csr_matrix) ----
// one warp per row; an entry is kept when !(|c| < thr) && c != 0 (scipy drops exact zeros)
__device__ __forceinline__ bool csr_keep(double v, double thr) { return !(fabs(v) < thr) && v != 0.0; }

__global__ void csr_count_kernel(const double *__restrict__ dense, int64_t nrows, int64_t ncols,
                                 int64_t bs, double thr, int64_t *__restrict__ counts) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    const double *p = dense + row * bs;
    int n = 0;
    for (int64_t c = lane; c < ncols; c += 32) n += csr_keep(p[c], thr) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) counts[row] = n;
}

__global__ void csr_fill_kernel(const double *__restrict__ dense, int64_t nrows, int64_t ncols,
                                int64_t bs, double thr, const int64_t *__restrict__ indptr,
                                int32_t *__restrict__ indices, double *__restrict__ data) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nrows) return;
    const double *p = dense + row * bs;
    int64_t pos = indptr[row];
    for (int64_t c0 = 0; c0 < ncols; c0 += 32) {  // ascending column order within the row
        const int64_t c = c0 + lane;
        const double v = c < ncols ? p[c] : 0.0;
        const bool keep = c < ncols && csr_keep(v, thr);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int64_t at = pos + __popc(m & ((1u << lane) - 1u));
            indices[at] = (int32_t)c;
            data[at] = v;
        }
        pos += __popc(m);
    }
}

static inline int half_up(int n) { return (n + 1) / 2; }

// does a working set of `doubles` fit one CTA's shared memory (GI_DWT_FUSED=0 switches the fused
// small-volume kernels off)
static bool small_dwt_smem(int64_t doubles) {
    static const bool off = getenv("GI_DWT_FUSED") && getenv("GI_DWT_FUSED")[0] == '0';
    return !off && doubles * (int64_t)sizeof(double) <= 200 * 1024;
}

// transform the (n0,n1,n2) sub-box of `in` (full dims d0.. ) along all three axes -> out with
// dims (2*no0, 2*no1, 2*no2); tmpA/tmpB are scratch of at least out size
static int dwt3_level(const double *in, int64_t in_bs, int d1, int d2, int n0, int n1, int n2,
                      double *tmpA, double *tmpB, double *out, int batch, cudaStream_t s) {
    const int o0 = 2 * half_up(n0), o1 = 2 * half_up(n1), o2 = 2 * half_up(n2);
    auto launch = [&](const double *src, int64_t sbs, int sd1, int sd2, int a0, int a1, int a2,
                      int axis, double *dst, int e0, int e1, int e2) {
        int m[3] = {a0, a1, a2};
        m[axis] = half_up(m[axis]);
        const int64_t total = (int64_t)m[0] * m[1] * m[2];
        dim3 grid((unsigned)ceil_div(total, 256), (unsigned)batch);
        dwt_axis_kernel<<<grid, 256, 0, s>>>(src, sbs, sd1, sd2, a0, a1, a2, axis, dst,
                                             (int64_t)e0 * e1 * e2, e1, e2);
    };
    // axis 0: (n0,n1,n2) -> (o0,n1,n2)
    launch(in, in_bs, d1, d2, n0, n1, n2, 0, tmpA, o0, n1, n2);
    // axis 1: (o0,n1,n2) -> (o0,o1,n2)
    launch(tmpA, (int64_t)o0 * n1 * n2, n1, n2, o0, n1, n2, 1, tmpB, o0, o1, n2);
    // axis 2: (o0,o1,n2) -> (o0,o1,o2)
    launch(tmpB, (int64_t)o0 * o1 * n2, o1, n2, o0, o1, n2, 2, out, o0, o1, o2);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

}  // namespace gi

using namespace gi;

extern "C" int gi_dwt_db4_l2_3d_batch(const double *x, int64_t batch, int64_t x_bs, int32_t nz,
                                      int32_t ny, int32_t nx, double *out, int64_t out_bs,
                                      int32_t out_shape[3], void *stream) {
    GI_REQUIRE(nz > 0 && ny > 0 && nx > 0 && batch >= 0, "gi_dwt_db4_l2_3d: bad shape");
    const int h1[3] = {half_up(nz), half_up(ny), half_up(nx)};
    const int h2[3] = {half_up(h1[0]), half_up(h1[1]), half_up(h1[2])};
    const int f[3] = {2 * h2[0] + h1[0], 2 * h2[1] + h1[1], 2 * h2[2] + h1[2]};
    if (out_shape) { out_shape[0] = f[0]; out_shape[1] = f[1]; out_shape[2] = f[2]; }
    if (!out || batch == 0) return GI_OK;
    GI_REQUIRE(x, "gi_dwt_db4_l2_3d: null input");
    GI_REQUIRE(batch <= 65535, "gi_dwt_db4_l2_3d: batch too large (<= 65535)");
    const int64_t fsz = (int64_t)f[0] * f[1] * f[2];
    GI_REQUIRE(out_bs >= fsz && x_bs >= (int64_t)nz * ny * nx, "gi_dwt_db4_l2_3d: bad strides");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t t1sz = 8LL * h1[0] * h1[1] * h1[2], t2sz = 8LL * h2[0] * h2[1] * h2[2];
    if (batch == 1 && small_dwt_smem(5 * t1sz / 2 + 1)) {
        // one small volume (the per-evaluation transform of the model): one launch instead of ten
        // stream operations
        const size_t smem = sizeof(double) * (size_t)(5 * t1sz / 2 + 1);
        GI_CUDA(cudaFuncSetAttribute(dwt3d_l2_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dwt3d_l2_small_kernel<<<1, 1024, smem, s>>>(x, nz, ny, nx, out);
        GI_LAUNCH_CHECK();
        return GI_OK;
    }
    double *ws = nullptr;
    GI_CUDA(cudaMallocAsync(&ws, sizeof(double) * batch * (3 * t1sz + t2sz), s));
    double *T1 = ws, *tA = ws + batch * t1sz, *tB = tA + batch * t1sz, *T2 = tB + batch * t1sz;
    int rc = dwt3_level(x, x_bs, ny, nx, nz, ny, nx, tA, tB, T1, (int)batch, s);
    // level 2 on the 'aaa' corner of T1 (dims 2*h1, sub-box h1)
    if (!rc) rc = dwt3_level(T1, t1sz, 2 * h1[1], 2 * h1[2], h1[0], h1[1], h1[2], tA, tB, T2, (int)batch, s);
    if (!rc) {
        cudaMemsetAsync(out, 0, sizeof(double) * out_bs * batch, s);
        dim3 grid((unsigned)ceil_div(t1sz + t2sz, 256), (unsigned)batch);
        pack3d_kernel<<<grid, 256, 0, s>>>(T1, T2, h1[0], h1[1], h1[2], h2[0], h2[1], h2[2], out, t1sz,
                                           t2sz, out_bs);
    }
    cudaFreeAsync(ws, s);
    if (rc) return rc;
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_dwt_db4_l2_3d(const double *x, int32_t nz, int32_t ny, int32_t nx, double *out,
                                int32_t out_shape[3], void *stream) {
    int32_t shp[3];
    int rc = gi_dwt_db4_l2_3d_batch(nullptr, 0, 0, nz, ny, nx, nullptr, 0, shp, stream);
    if (rc) return rc;
    if (out_shape) { out_shape[0] = shp[0]; out_shape[1] = shp[1]; out_shape[2] = shp[2]; }
    if (!out) return GI_OK;
    return gi_dwt_db4_l2_3d_batch(x, 1, (int64_t)nz * ny * nx, nz, ny, nx, out,
                                  (int64_t)shp[0] * shp[1] * shp[2], shp, stream);
}

extern "C" int gi_dwt_db4_l2_1d_batch(const double *x, int64_t batch, int64_t x_bs, int64_t n,
                                      double *out, int64_t out_bs, int64_t *ncoef, void *stream) {
    GI_REQUIRE(n > 0 && n < (1LL << 31) && batch >= 0, "gi_dwt_db4_l2_1d: bad length");
    const int h1 = half_up((int)n), h2 = half_up(h1);
    const int64_t nc = 2LL * h2 + h1;
    if (ncoef) *ncoef = nc;
    if (!out || batch == 0) return GI_OK;
    GI_REQUIRE(x, "gi_dwt_db4_l2_1d: null input");
    GI_REQUIRE(batch <= 65535, "gi_dwt_db4_l2_1d: batch too large (<= 65535)");
    GI_REQUIRE(out_bs >= nc && x_bs >= n, "gi_dwt_db4_l2_1d: bad strides");
    cudaStream_t s = (cudaStream_t)stream;
    if (batch == 1 && small_dwt_smem(2LL * h1)) {
        const size_t smem = sizeof(double) * (size_t)(2 * h1);
        GI_CUDA(cudaFuncSetAttribute(dwt1d_l2_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dwt1d_l2_small_kernel<<<1, 1024, smem, s>>>(x, (int)n, out);
        GI_LAUNCH_CHECK();
        return GI_OK;
    }
    double *T1 = nullptr;
    GI_CUDA(cudaMallocAsync(&T1, sizeof(double) * batch * 2 * h1, s));
    dim3 g1((unsigned)ceil_div(h1, 256), (unsigned)batch), g2((unsigned)ceil_div(h2, 256), (unsigned)batch);
    // level 1: x -> T1 = [a1 | d1]
    dwt_axis_kernel<<<g1, 256, 0, s>>>(x, x_bs, 1, (int)n, 1, 1, (int)n, 2, T1, 2LL * h1, 1, 2 * h1);
    // level 2: a1 -> out[0 : 2*h2] = [a2 | d2]
    dwt_axis_kernel<<<g2, 256, 0, s>>>(T1, 2LL * h1, 1, 2 * h1, 1, 1, h1, 2, out, out_bs, 1, 2 * h2);
    // d1 -> out[2*h2 : 2*h2 + h1]
    cudaMemcpy2DAsync(out + 2 * h2, sizeof(double) * out_bs, T1 + h1, sizeof(double) * 2 * h1,
                      sizeof(double) * h1, batch, cudaMemcpyDeviceToDevice, s);
    cudaFreeAsync(T1, s);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_dwt_db4_l2_1d(const double *x, int64_t n, double *out, int64_t *ncoef,
                                void *stream) {
    int64_t nc = 0;
    int rc = gi_dwt_db4_l2_1d_batch(nullptr, 0, 0, n, nullptr, 0, &nc, stream);
    if (rc) return rc;
    if (ncoef) *ncoef = nc;
    if (!out) return GI_OK;
    return gi_dwt_db4_l2_1d_batch(x, 1, n, n, out, nc, nullptr, stream);
}

extern "C" int gi_csr_spmv(const int64_t *indptr, const int32_t *indices, const double *data,
                           int64_t nrows, const double *x, double *y, void *stream) {
    GI_REQUIRE(nrows >= 0, "gi_csr_spmv: bad shape");
    if (nrows == 0) return GI_OK;
    GI_REQUIRE(indptr && x && y, "gi_csr_spmv: null pointer");
    csr_spmv_kernel<<<(unsigned)ceil_div(nrows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        indptr, indices, data, nrows, x, y);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_csr_count(const double *dense, int64_t nrows, int64_t ncols, int64_t row_stride,
                            double thr, int64_t *counts, void *stream) {
    GI_REQUIRE(nrows >= 0 && ncols >= 0 && row_stride >= ncols, "gi_csr_count: bad shape");
    if (nrows == 0) return GI_OK;
    GI_REQUIRE(dense && counts, "gi_csr_count: null pointer");
    csr_count_kernel<<<(unsigned)ceil_div(nrows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        dense, nrows, ncols, row_stride, thr, counts);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

extern "C" int gi_csr_fill(const double *dense, int64_t nrows, int64_t ncols, int64_t row_stride,
                           double thr, const int64_t *indptr, int32_t *indices, double *data,
                           void *stream) {
    GI_REQUIRE(nrows >= 0 && ncols >= 0 && ncols < (1LL << 31) && row_stride >= ncols,
               "gi_csr_fill: bad shape");
    if (nrows == 0) return GI_OK;
    GI_REQUIRE(dense && indptr && indices && data, "gi_csr_fill: null pointer");
    csr_fill_kernel<<<(unsigned)ceil_div(nrows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        dense, nrows, ncols, row_stride, thr, indptr, indices, data);
    GI_LAUNCH_CHECK();
    return GI_OK;
}
