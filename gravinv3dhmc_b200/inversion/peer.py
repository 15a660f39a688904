"""Peer-memory exchange of a row-sharded batch (csrc/peer.cu, include/gravinv_b200.h `gi_peer_*`).

The reference runs its chains as independent MPI processes that never communicate
(example/uniformgrid/run_main.sh:18); here ONE kernel matrix is partitioned by observation rows over
the GPUs of an NVLink / NVSwitch node and the ranks exchange, per gradient evaluation, the pieces of
potential.py:699-708 that depend on all rows -- through each other's memory, not through a collective
library: the adjoint contraction stores its tiles into the owner rank's HBM, the owners update their
column slice and the copy engines push the slices back, the scalars go through a slot table.

`PeerBuffer` owns one rank's symmetric buffer: every rank allocates the same number of bytes, the
64-byte CUDA IPC handles are all-gathered on the host (`torch.distributed.all_gather_object`:
plumbing) and every rank maps the others' buffers.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .. import _lib


class PeerBuffer:
    def __init__(self, nbytes, rank=0, world=1, group=None):
        _lib.require_cuda()
        self.L = L = _lib.lib()
        self.rank, self.world, self.group = int(rank), int(world), group
        self.h = C.c_void_p()
        _lib.check(L.gi_peer_create(self.rank, self.world, int(nbytes), C.byref(self.h)), "gi_peer_create")
        if self.world > 1:
            import torch.distributed as dist

            mine = (C.c_ubyte * 64)()
            _lib.check(L.gi_peer_export(self.h, mine), "gi_peer_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine), group=group)
            blob = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(handles))
            _lib.check(L.gi_peer_connect(self.h, blob), "gi_peer_connect")
            dist.barrier(group=group)

    def bytes_sent(self):
        return int(self.L.gi_peer_bytes_sent(self.h))

    def allreduce_small(self, t):
        """in-place sum over the ranks of a small float64 CUDA tensor (8..512 elements, multiple of 8)"""
        _lib.check(self.L.gi_peer_allreduce_small(self.h, _lib.ptr(t), int(t.numel()), _lib.stream_ptr()),
                   "gi_peer_allreduce_small")
        return t

    def close(self):
        if getattr(self, "h", None):
            if self.world > 1:
                import torch.distributed as dist

                _lib.sync()
                dist.barrier(group=self.group)  # nobody unmaps while another rank still stores here
            self.L.gi_peer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.world == 1:
                self.close()
        except Exception:
            pass


def exchange_mode(model):
    """'peer' (default on CUDA tensors) or 'nccl' (the all-reduce hook path): GI_SHARD_EXCHANGE"""
    return os.environ.get("GI_SHARD_EXCHANGE", "peer").lower()


def selftest_single_rank(model):
    """one rank is its own peer: the whole peer machinery (slot table, staged partials, column-slice
    update, epoch flags) against the plain single-GPU handle on the same draws"""
    from . import batched

    M = model.M
    b = np.zeros((M, 2))
    b[:, 1] = 0.3
    one = np.full(M, 0.001)
    out = []
    for driver in ("device", "device-peer"):
        bt = batched.HMCBatch(model, 3, 0.02, [2, 4], one, one, b, "mandatory", 1000, model.dobs, 0.5, "TV",
                              0.001, 3, 0.05, save_folder=os.path.join("/tmp", "gi_peer_selftest"), quiet=True,
                              driver=driver)
        for _ in range(3):
            bt.propose()
        bt.stream(10 ** 6, 0, max_proposals=3, write=False)
        out.append((bt.x.copy(), [list(p) for p in bt.proposals]))
        bt.close()
    assert out[0][1] == out[1][1], "peer path: different accept decisions"
    assert np.max(np.abs(out[0][0] - out[1][0])) <= 1e-12 * np.max(np.abs(out[0][0])), "peer path: positions differ"
    return True
