// DMMA.8x8x4 issue-rate probe with the operand pattern of gemm_adj_kernel: 4 A fragments x NB B
// fragments -> 4*NB accumulators, all distinct registers.  Variants differ in issue order.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_operands dmma_operands.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// ORDER 0: j outer, i inner (kernel order).  1: i outer, j inner.  2: same a and b for all (reuse)
template <int NB, int ORDER>
__global__ void k(double *out, const double *in) {
    double acc[4][NB][2], a[4], b[NB];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = in[threadIdx.x + 32 * i];
#pragma unroll
    for (int j = 0; j < NB; ++j) b[j] = in[threadIdx.x + 32 * (4 + j)];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int it = 0; it < ITERS; ++it) {
        if (ORDER == 0) {
#pragma unroll
            for (int j = 0; j < NB; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        } else if (ORDER == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        } else {
#pragma unroll
            for (int j = 0; j < NB; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma884(acc[i][j][0], acc[i][j][1], a[0], b[0]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NB; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> static double timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize(); cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5 * 1e-3;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out, *in; cudaMalloc(&out, 8 * sms * 1024); cudaMalloc(&in, 8 * 4096); cudaMemset(in, 0, 8 * 4096);
#define RUN(NB, ORDER, THREADS)                                                                  \
    { double t = timeit([&] { k<NB, ORDER><<<sms, THREADS>>>(out, in); });                        \
      printf("NB=%d order=%d threads=%4d : %6.2f TFLOP/s\n", NB, ORDER, THREADS,                   \
             (double)sms * THREADS / 32 * ITERS * 4 * NB * 512 / t / 1e12); }
    RUN(8, 0, 256) RUN(8, 1, 256) RUN(8, 2, 256)
    RUN(4, 0, 512) RUN(4, 1, 512) RUN(4, 2, 512)
    RUN(8, 0, 128) RUN(4, 0, 256) RUN(4, 0, 1024) RUN(2, 0, 512) RUN(2, 1, 1024)
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
