// sink.cu -- on-device sample sink: running posterior mean / variance of the accepted models
// (SURVEY.md 8(f2)).  The reference appends every accepted model to <chain>/model.dat as text
// (inversion/hmc.py:241-249, one row of M "%.8f" numbers -- ~10 MB per sample at c5) and its
// post-processing reads the last `last` rows back to form np.mean / np.std per voxel and the forward
// data of both (example/uniformgrid/plot_uniform.py:44-135).  Here the accepted state never leaves
// the GPU: each sampler handle adds m = WmInv mw to a Welford accumulator of its chain right after
// the commit, gated ON THE DEVICE by the Metropolis flag and by the window "skip the first `skip`
// accepted samples, then take `take`" (= the reference's ndraws / last), so the streaming sampler
// needs no device->host traffic for it.
#include <string.h>

#include <vector>

#include "plan.cuh"
#include "sink.cuh"

using namespace gi;

struct gi_stats {
    int64_t M, ld;
    int32_t nslots;
    double *mean, *m2;   // [nslots + 1][ld]; the extra slot holds pooled results
    int64_t *seen, *count;  // [nslots + 1] accepted samples offered / accumulated
    int32_t *gate;          // [nslots]
    int64_t skip, take;
    int64_t launches;
};

namespace {

struct StatsDev {
    double *mean, *m2;
    int64_t *seen, *count;
    int32_t *gate;
    int64_t skip, take, M, ld;
};

StatsDev dev_of(const gi_stats *s) {
    return StatsDev{s->mean, s->m2, s->seen, s->count, s->gate, s->skip, s->take, s->M, s->ld};
}

// one thread per chain: does this finished proposal enter the statistics?
__global__ void stats_gate_kernel(StatsMap map, const DevState *__restrict__ st, int C, int slot0, StatsDev sd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int slot = slot0 + c;
    int g = 0;
    if (map.fin[c] && (!st || st[c].res.accept)) {
        const int64_t seen = ++sd.seen[slot];
        g = seen > sd.skip && (sd.take <= 0 || seen <= sd.skip + sd.take);
        if (g) sd.count[slot] += 1;
    }
    sd.gate[slot] = g;
}

// Welford update of the gated chains with m = scale * mw  (hmc.py:327: m = WmInv @ mw)
__global__ void __launch_bounds__(256)
stats_add_kernel(int slot0, const double *__restrict__ mw, int64_t mw_ld, const double *__restrict__ scale,
                 StatsDev sd) {
    const int slot = slot0 + blockIdx.y;
    if (!sd.gate[slot]) return;
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= sd.M) return;
    const double n = (double)sd.count[slot];
    const double v = mw[(int64_t)blockIdx.y * mw_ld + j];
    const double m = scale ? __dmul_rn(scale[j], v) : v;
    const int64_t o = (int64_t)slot * sd.ld + j;
    const double mean = sd.mean[o];
    const double delta = m - mean;
    const double mean_new = mean + delta / n;
    sd.mean[o] = mean_new;
    sd.m2[o] += delta * (m - mean_new);
}

// pooled statistics over all slots into slot `nslots` (Chan et al. pairwise merge, slot order)
__global__ void stats_pool_kernel(int nslots, StatsDev sd) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) {
        int64_t n = 0, seen = 0;
        for (int s = 0; s < nslots; ++s) { n += sd.count[s]; seen += sd.seen[s]; }
        sd.count[nslots] = n;
        sd.seen[nslots] = seen;
    }
    if (j >= sd.M) return;
    double n = 0.0, mean = 0.0, m2 = 0.0;
    for (int s = 0; s < nslots; ++s) {
        const double nb = (double)sd.count[s];
        if (nb == 0.0) continue;
        const double mb = sd.mean[(int64_t)s * sd.ld + j], vb = sd.m2[(int64_t)s * sd.ld + j];
        const double nt = n + nb, delta = mb - mean;
        mean += delta * (nb / nt);
        m2 += vb + delta * delta * (n * nb / nt);
        n = nt;
    }
    sd.mean[(int64_t)nslots * sd.ld + j] = mean;
    sd.m2[(int64_t)nslots * sd.ld + j] = m2;
}

// mean_out = scale * mean, std_out = scale * sqrt(m2 / n)   (np.std, ddof = 0)
__global__ void stats_result_kernel(int slot, const double *__restrict__ scale, double *__restrict__ mean_out,
                                    double *__restrict__ std_out, StatsDev sd) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= sd.ld) return;
    double mean = 0.0, sdv = 0.0;
    const double n = (double)sd.count[slot];
    if (j < sd.M && n > 0.0) {
        mean = sd.mean[(int64_t)slot * sd.ld + j];
        sdv = sqrt(sd.m2[(int64_t)slot * sd.ld + j] / n);
        if (scale) { mean *= scale[j]; sdv *= scale[j]; }
    }
    if (mean_out) mean_out[j] = mean;
    if (std_out) std_out[j] = sdv;
}

}  // namespace

int gi::stats_add_chains(gi_stats *s, const StatsMap &map, const DevState *st, const double *mw,
                         int64_t mw_ld, int C, int slot0, const double *scale, cudaStream_t stream) {
    if (!s) return GI_OK;
    GI_REQUIRE(slot0 >= 0 && slot0 + C <= s->nslots, "gi_stats: slot out of range");
    StatsDev sd = dev_of(s);
    stats_gate_kernel<<<1, 64, 0, stream>>>(map, st, C, slot0, sd);
    dim3 grid((unsigned)ceil_div(s->M, 256), (unsigned)C);
    stats_add_kernel<<<grid, 256, 0, stream>>>(slot0, mw, mw_ld, scale, sd);
    GI_LAUNCH_CHECK();
    s->launches += 2;
    return GI_OK;
}

extern "C" int gi_stats_create(int64_t M, int64_t ld, int32_t nslots, gi_stats **out) {
    GI_REQUIRE(out && M > 0 && ld >= M && nslots >= 1 && nslots <= 64, "gi_stats_create: bad argument");
    gi_stats *s = new gi_stats();
    memset(s, 0, sizeof(*s));
    s->M = M; s->ld = ld; s->nslots = nslots;
    const size_t vb = sizeof(double) * (nslots + 1) * ld;
    cudaError_t e = cudaMalloc(&s->mean, vb);
    if (e == cudaSuccess) e = cudaMalloc(&s->m2, vb);
    if (e == cudaSuccess) e = cudaMalloc(&s->seen, sizeof(int64_t) * (nslots + 1));
    if (e == cudaSuccess) e = cudaMalloc(&s->count, sizeof(int64_t) * (nslots + 1));
    if (e == cudaSuccess) e = cudaMalloc(&s->gate, sizeof(int32_t) * (nslots + 1));
    if (e != cudaSuccess) {
        gi_stats_destroy(s);
        return cuda_fail(e, "gi_stats_create", __FILE__, __LINE__);
    }
    *out = s;
    return gi_stats_reset(s);
}

extern "C" int gi_stats_destroy(gi_stats *s) {
    if (!s) return GI_OK;
    cudaFree(s->mean); cudaFree(s->m2); cudaFree(s->seen); cudaFree(s->count); cudaFree(s->gate);
    delete s;
    return GI_OK;
}

extern "C" int gi_stats_reset(gi_stats *s) {
    GI_REQUIRE(s, "gi_stats_reset: null handle");
    const size_t vb = sizeof(double) * (s->nslots + 1) * s->ld;
    GI_CUDA(cudaMemset(s->mean, 0, vb));
    GI_CUDA(cudaMemset(s->m2, 0, vb));
    GI_CUDA(cudaMemset(s->seen, 0, sizeof(int64_t) * (s->nslots + 1)));
    GI_CUDA(cudaMemset(s->count, 0, sizeof(int64_t) * (s->nslots + 1)));
    GI_CUDA(cudaMemset(s->gate, 0, sizeof(int32_t) * (s->nslots + 1)));
    return GI_OK;
}

extern "C" int gi_stats_window(gi_stats *s, int64_t skip, int64_t take) {
    GI_REQUIRE(s && skip >= 0, "gi_stats_window: bad argument");
    s->skip = skip;
    s->take = take;
    return GI_OK;
}

extern "C" int gi_stats_add(gi_stats *s, int32_t slot, const double *mw_dev, const double *scale_dev,
                            void *stream) {
    GI_REQUIRE(s && mw_dev && slot >= 0 && slot < s->nslots, "gi_stats_add: bad argument");
    StatsMap map;
    memset(&map, 0, sizeof(map));
    map.fin[0] = 1;
    return stats_add_chains(s, map, nullptr, mw_dev, s->ld, 1, slot, scale_dev, (cudaStream_t)stream);
}

extern "C" int gi_stats_result_dev(gi_stats *s, int32_t slot, const double *scale_dev, double *mean_dev,
                                   double *std_dev, int64_t *count, int64_t *seen, void *stream) {
    GI_REQUIRE(s && slot >= -1 && slot < s->nslots, "gi_stats_result: bad slot");
    cudaStream_t st = (cudaStream_t)stream;
    StatsDev sd = dev_of(s);
    if (slot < 0) {
        stats_pool_kernel<<<(unsigned)ceil_div(s->M, 256), 256, 0, st>>>(s->nslots, sd);
        slot = s->nslots;
        s->launches += 1;
    }
    stats_result_kernel<<<(unsigned)ceil_div(s->ld, 256), 256, 0, st>>>(slot, scale_dev, mean_dev, std_dev, sd);
    GI_LAUNCH_CHECK();
    s->launches += 1;
    int64_t hc[2] = {0, 0};
    GI_CUDA(cudaMemcpyAsync(&hc[0], s->count + slot, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    GI_CUDA(cudaMemcpyAsync(&hc[1], s->seen + slot, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    GI_CUDA(cudaStreamSynchronize(st));
    if (count) *count = hc[0];
    if (seen) *seen = hc[1];
    return GI_OK;
}

extern "C" int gi_stats_result(gi_stats *s, int32_t slot, double *mean_host, double *std_host,
                               int64_t *count, int64_t *seen, void *stream) {
    GI_REQUIRE(s, "gi_stats_result: null handle");
    double *tmp = nullptr;
    GI_CUDA(cudaMalloc(&tmp, sizeof(double) * 2 * s->ld));
    int rc = gi_stats_result_dev(s, slot, nullptr, tmp, tmp + s->ld, count, seen, stream);
    cudaError_t e = cudaSuccess;
    if (!rc && mean_host) e = cudaMemcpy(mean_host, tmp, sizeof(double) * s->M, cudaMemcpyDeviceToHost);
    if (!rc && e == cudaSuccess && std_host)
        e = cudaMemcpy(std_host, tmp + s->ld, sizeof(double) * s->M, cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "gi_stats_result", __FILE__, __LINE__);
    return GI_OK;
}

extern "C" int64_t gi_stats_launch_count(const gi_stats *s) { return s ? s->launches : 0; }
