"""CPU, 2..3 processes: the shared-memory ring that carries the reference-order draws (L, u, p0)
between the ranks of one node in row-sharded streaming runs (inversion/batched.py: _DrawRing fed by
_DrawAhead).  Every proposal of every chain must arrive on every rank intact and in order -- equal to
what `RandomState(seed + c)` yields in the reference's call order (hmc.py:297,95,165) -- with slots
reused only after all ranks are done with them."""
import multiprocessing as mp
import os
import tempfile
import time

import numpy as np
import pytest

from gravinv3dhmc_b200.inversion.batched import _DrawAhead, _DrawRing

NCH, M, NPROP, SEED, SIGMA, LR = 5, 3000, 9, 40, 0.25, [5, 20]


def worker(path, world, rank, depth, q):
    try:
        ring = _DrawRing(NCH, M, world, rank, depth=depth, path=path)
        owned = [c % world == rank for c in range(NCH)]
        streams = [np.random.RandomState(SEED + c) for c in range(NCH)]
        ahead = _DrawAhead(streams, LR, M, SIGMA, ring, nworkers=2, owned=owned)
        ahead.limit = NPROP
        ahead.start()
        bad = 0
        ref = [np.random.RandomState(SEED + c) for c in range(NCH)]
        for k in range(NPROP):
            for c in range(NCH):
                ring.wait_ready(c, k)
                row = ring.data[c, k % depth].copy()
                ring.release(c, k)
                L = int(ref[c].randint(LR[0], LR[1] + 1))
                p0 = ref[c].randn(M) * SIGMA
                u = float(ref[c].rand())
                bad += not (row[M] == L and row[M + 1] == u and np.array_equal(row[:M], p0))
            if rank == world - 1 and k % 3 == 0:
                time.sleep(0.01)  # a slow consumer throttles the producers, no draw is lost
        ahead.stop()
        ring.close()
        q.put((rank, bad))
    except BaseException as e:  # noqa: BLE001
        q.put((rank, repr(e)))


@pytest.mark.parametrize("world,depth", [(2, 3), (3, 2)])
def test_ring_delivers_reference_order_draws(world, depth):
    fd, path = tempfile.mkstemp(prefix="gi_ring_test_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    os.close(fd)
    try:
        _DrawRing(NCH, M, world, 0, depth=depth, path=path, create=True).close()
        ctx = mp.get_context("fork")
        q = ctx.Queue()
        procs = [ctx.Process(target=worker, args=(path, world, r, depth, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = dict(q.get(timeout=120) for _ in range(world))
        for p in procs:
            p.join(timeout=30)
        assert res == {r: 0 for r in range(world)}, res
    finally:
        if os.path.exists(path):
            os.unlink(path)
