// peer.cu -- the symmetric peer-memory buffer of a row-sharded run (see peer.cuh) and the small
// all-ranks scalar reduction built on it.
//
// Replaces, on the multi-GPU path, what the reference does not have at all: its chains are
// independent OS processes (example/*/run_main.sh:18 `mpiexec -n K`), each with a full copy of Aw.
// Here Aw is partitioned by observation rows (the reference's own worker chunking,
// gravmag/prism.py:986-996) and the pieces of every gradient evaluation that depend on all rows
// (potential.py:699-708: mean(d), |r|^2, Aw^T r) are exchanged over NVLink by these primitives.
#include <string.h>

#include "peer.cuh"

namespace gi {
namespace {

constexpr unsigned long long kPeerSpinLimit = 1ull << 27;  // ~ seconds; a lost peer traps instead of hanging

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// one CTA of 64 threads; thread c owns chain c
__global__ void __launch_bounds__(64)
peer_scalars_kernel(PeerScalArgs a, const double *src, double *dst, int C, int col0,
                    int ncol, unsigned long long seq, unsigned long long *epoch_slot,
                    unsigned long long epoch_val) {
    const int c = threadIdx.x;
    const int par = (int)(seq & 1ull);
    const int64_t slot_me = ((int64_t)(par * kPeerMax + a.me) * 64 + c) * 8;
    if (c < C) {
        double v[8];
        for (int k = 0; k < ncol; ++k) v[k] = src[c * 8 + col0 + k];
        for (int q = 0; q < a.nranks; ++q)
            for (int k = 0; k < ncol; ++k) a.scal[q][slot_me + col0 + k] = v[k];
    }
    __threadfence_system();
    __syncthreads();
    if (c < a.nranks) {
        st_release_sys(a.flag[c] + a.me, seq + 1ull);  // tell rank c: my values of exchange `seq` are in
        unsigned long long spins = 0;
        while (ld_acquire_sys(a.flag[a.me] + c) < seq + 1ull)
            if (++spins > kPeerSpinLimit) __trap();
    }
    __syncthreads();
    if (c < C) {
        const double *mine = a.scal[a.me];
        for (int k = 0; k < ncol; ++k) {
            double t = 0.0;
            for (int q = 0; q < a.nranks; ++q)  // rank order: the same bits on every rank
                t += __ldcg(mine + ((int64_t)(par * kPeerMax + q) * 64 + c) * 8 + col0 + k);
            dst[c * 8 + col0 + k] = t;
        }
    }
    if (epoch_slot && c == 0) *epoch_slot = epoch_val;
}

__global__ void peer_wait_x_kernel(const unsigned long long *flag, PeerMap map, unsigned long long epoch) {
    const int q = threadIdx.x;
    if (q >= map.nranks || q == map.me || map.col[q + 1] <= map.col[q]) return;
    unsigned long long spins = 0;
    while (ld_acquire_sys(flag + q) < epoch)
        if (++spins > kPeerSpinLimit) __trap();
}

}  // namespace

int peer_wait_x(gi_peer *p, const PeerMap &map, unsigned long long epoch, cudaStream_t s) {
    GI_REQUIRE(p && p->connected, "peer_wait_x: not connected");
    if (p->world == 1 || epoch == 0) return GI_OK;
    peer_wait_x_kernel<<<1, 32, 0, s>>>(reinterpret_cast<const unsigned long long *>(p->base[p->rank] + kPeerFlagX),
                                        map, epoch);
    GI_LAUNCH_CHECK();
    return GI_OK;
}

int peer_scalars(gi_peer *p, const double *src, double *dst, int C, int col0, int ncol,
                 unsigned long long *epoch_slot, unsigned long long epoch_val, cudaStream_t s) {
    GI_REQUIRE(p && p->connected && C <= 64 && col0 >= 0 && ncol >= 1 && col0 + ncol <= 8,
               "peer_scalars: bad argument");
    PeerScalArgs a;
    memset(&a, 0, sizeof(a));
    a.nranks = p->world;
    a.me = p->rank;
    for (int q = 0; q < p->world; ++q) {
        a.scal[q] = reinterpret_cast<double *>(p->base[q] + kPeerScal);
        a.flag[q] = reinterpret_cast<unsigned long long *>(p->base[q] + kPeerFlagS);
    }
    peer_scalars_kernel<<<1, 64, 0, s>>>(a, src, dst, C, col0, ncol, p->seq, epoch_slot, epoch_val);
    GI_LAUNCH_CHECK();
    p->seq += 1;
    p->nvlink_bytes += (int64_t)(p->world - 1) * (C * ncol * 8 + 8);
    return GI_OK;
}

}  // namespace gi

using namespace gi;

extern "C" int gi_peer_create(int32_t rank, int32_t world, int64_t bytes, gi_peer **out) {
    GI_REQUIRE(out && world >= 1 && world <= kPeerMax && rank >= 0 && rank < world && bytes >= kPeerCtlBytes,
               "gi_peer_create: bad argument (1..%d ranks, at least %lld bytes)", kPeerMax,
               (long long)kPeerCtlBytes);
    gi_peer *p = new gi_peer();
    memset(p, 0, sizeof(*p));
    p->rank = rank;
    p->world = world;
    p->bytes = bytes;
    void *buf = nullptr;
    cudaError_t e = cudaMalloc(&buf, (size_t)bytes);
    if (e == cudaSuccess) e = cudaMemset(buf, 0, (size_t)bytes);
    if (e != cudaSuccess) {
        if (buf) cudaFree(buf);
        delete p;
        return cuda_fail(e, "gi_peer_create", __FILE__, __LINE__);
    }
    p->base[rank] = static_cast<unsigned char *>(buf);
    p->connected = world == 1;
    *out = p;
    return GI_OK;
}

extern "C" int gi_peer_export(gi_peer *p, void *handle64) {
    GI_REQUIRE(p && handle64, "gi_peer_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    GI_CUDA(cudaIpcGetMemHandle(&h, p->base[p->rank]));
    memcpy(handle64, &h, 64);
    return GI_OK;
}

extern "C" int gi_peer_connect(gi_peer *p, const void *handles) {
    GI_REQUIRE(p && handles, "gi_peer_connect: null pointer");
    if (p->connected) return GI_OK;
    for (int q = 0; q < p->world; ++q) {
        if (q == p->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const unsigned char *>(handles) + 64 * q, 64);
        void *ptr = nullptr;
        GI_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        p->base[q] = static_cast<unsigned char *>(ptr);
    }
    p->connected = true;
    return GI_OK;
}

extern "C" int gi_peer_destroy(gi_peer *p) {
    if (!p) return GI_OK;
    cudaDeviceSynchronize();
    for (int q = 0; q < p->world; ++q) {
        if (!p->base[q]) continue;
        if (q == p->rank) cudaFree(p->base[q]);
        else cudaIpcCloseMemHandle(p->base[q]);
    }
    delete p;
    return GI_OK;
}

extern "C" int64_t gi_peer_bytes_sent(const gi_peer *p) { return p ? p->nvlink_bytes : 0; }

// sum over ranks of a small device vector (n <= 512 doubles), in place, same bits on every rank --
// the building block above exposed for setup-time reductions and for tests
extern "C" int gi_peer_allreduce_small(gi_peer *p, double *vec_dev, int32_t n, void *stream) {
    GI_REQUIRE(p && vec_dev && n >= 8 && n <= 512 && n % 8 == 0,
               "gi_peer_allreduce_small: 8..512 doubles, a multiple of 8");
    return peer_scalars(p, vec_dev, vec_dev, n / 8, 0, 8, nullptr, 0, (cudaStream_t)stream);  // [n/8][8]
}
