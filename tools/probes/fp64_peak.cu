// FP64 throughput probe for B200: DFMA vs DMMA (m8n8k4, m16n8k8, m16n8k16) vs both interleaved.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__global__ void dfma_kernel(double *out, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void dmma884_kernel(double *out, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma16816_kernel(double *out, double a, double b) {
    double c[NACC][4];
    double af[8], bf[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) af[i] = a + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) bf[i] = b + i;
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma16816(c[i], af, bf);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma1688_kernel(double *out, double a, double b) {
    double c[NACC][4];
    double af[4], bf[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) af[i] = a + i;
#pragma unroll
    for (int i = 0; i < 2; ++i) bf[i] = b + i;
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS / 2; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1688(c[i], af, bf);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// interleaved: per iteration NACC dmma884 + NF dfma per thread
template <int NACC, int NF>
__global__ void mixed_kernel(double *out, double a, double b) {
    double c[NACC][2], f[NF];
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 5 * 1e-3;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    for (int threads : {128, 256, 512, 1024}) {
        for (int bps : {1, 2, 4}) {
            if (threads * bps > 2048) continue;
            const int grid = sms * bps;
            const double nthr = (double)grid * threads, nwarp = nthr / 32;
            double t = timeit([&] { dfma_kernel<<<grid, threads>>>(out, 1.0000001, 1e-9); });
            printf("threads=%4d bps=%d  DFMA        %7.2f TFLOP/s\n", threads, bps, nthr * ITERS * 16 * 2 / t / 1e12);
            t = timeit([&] { dmma884_kernel<8><<<grid, threads>>>(out, 1.0000001, 1e-9); });
            printf("threads=%4d bps=%d  DMMA884 x8  %7.2f TFLOP/s\n", threads, bps, nwarp * ITERS * 8 * 512 / t / 1e12);
            t = timeit([&] { dmma1688_kernel<8><<<grid, threads>>>(out, 1.0000001, 1e-9); });
            printf("threads=%4d bps=%d  DMMA1688 x8 %7.2f TFLOP/s\n", threads, bps, nwarp * (ITERS / 2) * 8 * 2048 / t / 1e12);
            t = timeit([&] { dmma16816_kernel<8><<<grid, threads>>>(out, 1.0000001, 1e-9); });
            printf("threads=%4d bps=%d  DMMA16816x8 %7.2f TFLOP/s\n", threads, bps, nwarp * (ITERS / 4) * 8 * 4096 / t / 1e12);
            t = timeit([&] { mixed_kernel<4, 8><<<grid, threads>>>(out, 1.0000001, 1e-9); });
            printf("threads=%4d bps=%d  MIXED 4dmma+8dfma %7.2f TFLOP/s (dmma part %.2f, dfma part %.2f)\n", threads, bps,
                   (nwarp * ITERS * 4 * 512 + nthr * ITERS * 8 * 2) / t / 1e12, nwarp * ITERS * 4 * 512 / t / 1e12,
                   nthr * ITERS * 8 * 2 / t / 1e12);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
