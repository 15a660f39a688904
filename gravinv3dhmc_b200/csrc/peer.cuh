// peer.cuh -- peer-memory exchange of the row-sharded batch over NVLink / NVSwitch (peer.cu),
// internal interface used by batched.cu.
//
// Every rank allocates ONE buffer of the same size ("symmetric": the same offset means the same
// object on every rank), exports it with cudaIpcGetMemHandle and maps its peers' buffers with
// cudaIpcOpenMemHandle, so a kernel (or a copy engine) of rank a can store straight into rank b's
// HBM.  Three things travel that way, none through a collective library:
//   * the adjoint contraction's output tiles -- stored by gemm_adj_kernel's epilogue into the OWNER
//     rank's staging block [src rank][chain][owned columns] (a reduce-scatter whose transfer is
//     spread over the whole contraction);
//   * the positions of the owner's column slice after its fused update -- pushed to every peer's
//     X buffer by the copy engines on a side stream, slice q first to the rank that reads it first,
//     while the forward contraction already runs on the columns that have arrived (an all-gather
//     hidden under the next pass); a 64-bit epoch per source rank tells the forward kernel's CTAs
//     that the slice they are about to read has landed;
//   * per-chain scalars (sum d, sum r^2, Um, K) -- written into every peer's slot table by one small
//     kernel that then waits for the other ranks' flags and sums the slots in rank order, so every
//     rank gets the SAME bits (the Metropolis decisions must agree) without a collective call.
#pragma once
#include <stdint.h>

#include "common.cuh"

constexpr int kPeerMax = 16;  // ranks of one NVLink domain

namespace gi {

// column slices: rank q owns the columns [col[q], col[q + 1]) of every chain (multiples of the
// adjoint kernel's 256-column strip, except the last bound = ld; a slice may be empty)
struct PeerMap {
    int32_t nranks, me;
    int64_t col[kPeerMax + 1];
};

// gemm_adj_kernel epilogue: the tile of a strip goes to stage[owner] + me * src_stride + chain * ldp
struct PeerOut {
    double *stage[kPeerMax];  // every rank's staging block, as mapped into this rank's address space
    int64_t src_stride;       // Cp * ldp
    int64_t ldp;              // row pitch of the staging block (widest slice)
};

// gemm_fwd_kernel: wait until the X slices a tile reads have landed (epoch of the last push)
struct PeerWait {
    const unsigned long long *flag;  // [kPeerMax] local, written by the peers' copy engines
    unsigned long long epoch;        // 0: the input is complete locally, nothing to wait for
};

// what the batched kernels' launchers need in peer mode (owned by the gi_hmcb handle)
struct PeerLaunch {
    PeerMap map;
    PeerOut out;
    const unsigned long long *xflag;  // local X-slice flags
    const double *stage_local;        // this rank's staging block [nranks][Cp][ldp]
    int64_t kc_rot;                   // forward k-chunk rotation (own slice first)
};

struct PeerScalArgs {
    double *scal[kPeerMax];              // every rank's slot table [2][nranks][64][8]
    unsigned long long *flag[kPeerMax];  // every rank's flag row [kPeerMax]
    int32_t nranks, me;
};

}  // namespace gi

struct gi_peer {
    int32_t rank, world;
    int64_t bytes;
    unsigned char *base[kPeerMax];  // base[rank] = the local buffer, the others are IPC mappings
    bool connected;
    unsigned long long seq;  // scalar exchanges so far (same on every rank: same call sequence)
    unsigned long long xepoch;  // X-slice pushes so far (likewise)
    int64_t nvlink_bytes;    // bytes this rank stored into peers' memory so far (accounting)
};

namespace gi {

// control block at the start of the symmetric buffer
constexpr int64_t kPeerFlagS = 0;                       // unsigned long long [kPeerMax]: scalar exchange
constexpr int64_t kPeerFlagX = 128;                     // unsigned long long [kPeerMax]: X slices
constexpr int64_t kPeerScal = 4096;                     // double [2][kPeerMax][64][8]
constexpr int64_t kPeerCtlBytes = 4096 + 2 * kPeerMax * 64 * 8 * 8;  // = 135168, a multiple of 256

// sum over ranks of src[c * 8 + col], col in [col0, col0 + ncol), c < C: every rank writes its values
// into all slot tables, flags, waits for the others and adds the slots in rank order -> dst (may
// alias src).  epoch_slot (optional, local) receives epoch_val: the source of the X-slice flags the
// copy engines send after the pushes.
int peer_scalars(gi_peer *p, const double *src, double *dst, int C, int col0, int ncol,
                 unsigned long long *epoch_slot, unsigned long long epoch_val, cudaStream_t s);

// make the stream wait until the X slices of every other rank carry `epoch` (for readers of a pushed
// buffer other than the forward kernel, which polls the flags itself)
int peer_wait_x(gi_peer *p, const PeerMap &map, unsigned long long epoch, cudaStream_t s);

}  // namespace gi
