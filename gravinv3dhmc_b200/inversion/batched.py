"""Batched chains: what the reference does with `mpiexec -n C python main_*.py` (one OS process per
chain, each holding its own copy of Aw; example/uniformgrid/run_main.sh:18, main_uniform.py:20-22)
runs here as ONE device-resident loop in which the C chains are columns of a dense FP64 contraction
(`gi_hmcb_*`, DMMA tensor cores): Aw is streamed once per pass for all chains.

`HMCSampleBatch(...)` takes the arguments of `hmc.HMCSample` plus `nchains`; chain c behaves exactly
like the reference process with `myrank=c`: RNG stream `RandomState(seed + c)` consumed in the
reference's order (randint, randn, rand per proposal; hmc.py:297,95,165), output folder
`save_folder + str(c)` with the same `misfit.dat` / `model.dat`, and it stops being recorded once it
has `ndraws + nsamples` accepted proposals (hmc.py:295).
"""
from __future__ import annotations

import collections
import ctypes as C
import os
import sys
import threading

import numpy as np

from .. import _lib
from ._engine import reg_params


class _ShardedBatchState:
    """Row-sharded batch (multi-GPU): replicated [Cp][ld] chain state on every rank, the kernel rows
    split by observation; per gradient evaluation two all-reduces (Cp scalars, then Cp*ld + Cp
    doubles), see `_engine.BatchEngine`.  Every rank draws the same numbers (same seeds) and takes
    the same Metropolis decisions on the host from all-reduced sums."""

    def __init__(self, b):
        from ._engine import BatchEngine

        m = b.model
        lo, hi = m.rows
        fix = np.asarray(m.grav_fix, dtype=np.float64)[lo:hi] if m.fixed else None
        self.b = b
        self.eng = eng = BatchEngine(m.Aw_pad, m.M, b.nchains, m.dobs[lo:hi], float(np.mean(m.dobs)),
                                     m.n_total, fix, m.group)
        torch = eng.torch
        self.low, self.high, self.apr = eng.vec(b.low), eng.vec(b.high), eng.vec(b.aprior_model)
        self.logc = b.constraint == "logarithmic"
        self.x_cur, self.g_cur = eng.mat(b.x), eng.mat()
        self.xa, self.xb, self.p, self.gnew = eng.mat(), eng.mat(), eng.mat(), eng.mat()
        self.mw_cur = eng.mat() if self.logc else self.x_cur
        self.mwa = eng.mat() if self.logc else self.xa
        self.mwb = eng.mat() if self.logc else self.xb
        self.d_cur = torch.zeros_like(eng.d)
        self.L_dev = torch.zeros(eng.Cp, dtype=torch.int32, device=eng.dev)
        self.U = np.zeros(eng.Cp)
        self.Ud = np.zeros(eng.Cp)
        self.Um = np.zeros(eng.Cp)
        self._to_mw(self.x_cur, self.mw_cur)
        # gradient and potential at the start state (hmc.py:105)
        eng.data_pass(self.mw_cur)
        eng.update(b.reg, None, self.x_cur, self.mw_cur, self.apr, m.wmsq_dev, self.low, self.high,
                   self.p, self.xa, self.mwa, self.g_cur, 0.0, None, 0, 1)
        s = eng.sums.cpu().numpy()
        self.Ud[:], self.Um[:] = s[:, 1], s[:, 2]
        self.U[:] = self.Ud + b.reg.alpha * self.Um
        self.d_cur.copy_(eng.d)

    def _to_mw(self, x, mw):
        if not self.logc:
            return
        t, b = self.eng.torch, self.b
        e = t.pow(t.tensor(np.e, dtype=t.float64, device=x.device), b.log_factor * x)
        mw.copy_((self.low[None, :] + self.high[None, :] * e) / (1 + e))
        mw[:, self.eng.M:] = 0

    def leapfrog_steps(self, p0_dev, nsteps):
        """benchmark helper (like gi_hmcb_leapfrog_steps): nsteps leapfrog steps of every chain
        from the current state, no Metropolis test"""
        b, eng, m = self.b, self.eng, self.b.model
        dt = float(b.dt)
        self.p.copy_(p0_dev)
        self.L_dev.fill_(int(nsteps) + 1)
        eng.update(b.reg, self.g_cur, self.x_cur, self.mw_cur, self.apr, m.wmsq_dev, self.low,
                   self.high, self.p, self.xa, self.mwa, None, dt, self.L_dev, 0, 3)
        xin, xout, mwin, mwout = self.xa, self.xb, self.mwa, self.mwb
        for i in range(1, int(nsteps) + 1):
            eng.data_pass(mwin)
            eng.update(b.reg, None, xin, mwin, self.apr, m.wmsq_dev, self.low, self.high, self.p, xout,
                       mwout, self.gnew, dt, self.L_dev, i, 0)
            xin, xout = xout, xin
            mwin, mwout = (mwout, mwin) if self.logc else (xin, xout)

    def propose(self, p0, Ls, u, res, tx, tu):
        b, eng, m = self.b, self.eng, self.b.model
        torch = eng.torch
        nc, M, dt, alpha = b.nchains, m.M, float(b.dt), b.reg.alpha
        self.p.zero_()
        self.p[:nc, :M] = torch.as_tensor(p0, device=eng.dev)
        Lpad = np.zeros(eng.Cp, dtype=np.int32)
        Lpad[:nc] = Ls
        self.L_dev.copy_(torch.as_tensor(Lpad))
        if tx is not None:
            tx[0], tu[0] = self.x_cur[:nc, :M].cpu().numpy(), self.U[:nc]
        eng.update(b.reg, self.g_cur, self.x_cur, self.mw_cur, self.apr, m.wmsq_dev, self.low,
                   self.high, self.p, self.xa, self.mwa, None, dt, self.L_dev, 0, 3)
        xin, xout, mwin, mwout = self.xa, self.xb, self.mwa, self.mwb
        for i in range(1, int(Ls.max()) + 1):
            eng.data_pass(mwin)
            eng.update(b.reg, None, xin, mwin, self.apr, m.wmsq_dev, self.low, self.high, self.p, xout,
                       mwout, self.gnew, dt, self.L_dev, i, 0)
            if tx is not None:
                s = eng.sums.cpu().numpy()
                tx[i], tu[i] = xin[:nc, :M].cpu().numpy(), (s[:, 1] + alpha * s[:, 2])[:nc]
            xin, xout = xout, xin
            mwin, mwout = (mwout, mwin) if self.logc else (xin, xout)
        s = eng.sums.cpu().numpy()
        acc_idx = []
        for c in range(nc):
            r = res[c]
            r.L = int(Ls[c])
            if Ls[c] == 0:
                r.accept = 0
                continue
            Ud, Um, Knew, K0 = s[c, 1], s[c, 2], s[c, 3], s[c, 5]
            Unew = Ud + alpha * Um
            Hcur, Hnew = K0 + self.U[c], Knew + Unew
            acc = bool(Hnew < Hcur or u[c] < np.exp(-(Hnew - Hcur)))   # hmc.py:167
            if acc:
                self.U[c], self.Ud[c], self.Um[c] = Unew, Ud, Um
                acc_idx.append(c)
            r.accept = int(acc)
            r.U, r.U_data, r.U_model = self.U[c], self.Ud[c], self.Um[c]
            r.Hcur, r.Hnew, r.Unew, r.Unew_data, r.Unew_model = Hcur, Hnew, Unew, Ud, Um
        if acc_idx:
            idx = torch.as_tensor(acc_idx, device=eng.dev)
            self.x_cur[idx] = xin[idx]
            if self.logc:
                self.mw_cur[idx] = mwin[idx]
            self.g_cur[idx] = self.gnew[idx]
            self.d_cur[idx] = eng.d[idx]


class _DrawRing:
    """Per-chain ring of draw slots shared by the ranks of ONE node.  Slot (c, k % depth) carries
    proposal k of chain c: `[p0 (M doubles) | L | u]`.  The chain's owner rank generates straight into
    the slot and publishes `ready[c][slot] = k + 1`; every rank copies the slot to its GPU and
    publishes `done[rank][c] = k + 1`; the owner rewrites a slot only when every rank is done with
    its previous occupant.  The backing store is a memory-mapped file under /dev/shm (row-sharded
    runs; registered with CUDA so the host->device copies read it directly) or pinned memory (one
    rank).  No collective and no device work of another rank is involved, so the threads that use
    it can never entangle with the NCCL all-reduces of the device loop."""

    def __init__(self, nchains, M, world=1, rank=0, depth=3, path=None, create=False):
        import torch

        self.nc, self.M, self.world, self.rank, self.depth = nchains, M, world, rank, depth
        nctl = nchains * depth + world * nchains
        ndata = nchains * depth * (M + 2)
        self.path, self.mm, self.registered = path, None, False
        if path is None:
            self.ctl = np.zeros(nctl, dtype=np.int64)
            self.tdata = torch.zeros(ndata, dtype=torch.float64).pin_memory()
            data = self.tdata.numpy()
        else:
            import mmap

            off = (8 * nctl + 4095) // 4096 * 4096
            nbytes = off + 8 * ndata
            if create:
                with open(path, "wb") as f:
                    f.truncate(nbytes)
            self.f = open(path, "r+b")
            self.mm = mmap.mmap(self.f.fileno(), nbytes)
            self.ctl = np.frombuffer(self.mm, dtype=np.int64, count=nctl)
            data = np.frombuffer(self.mm, dtype=np.float64, count=ndata, offset=off)
            self.tdata = torch.from_numpy(data)
            # page-lock the mapping so cudaMemcpyAsync reads it directly (else torch stages the copy)
            if torch.cuda.is_available():
                rc = torch.cuda.cudart().cudaHostRegister(self.tdata.data_ptr(), 8 * ndata, 0)
                self.registered = int(rc) == 0
        self.ready = self.ctl[: nchains * depth].reshape(nchains, depth)
        self.done = self.ctl[nchains * depth:].reshape(world, nchains)
        # flag words are written with a release store and read with an acquire load (C helpers):
        # the payload is visible before the flag on every host architecture, not only x86
        self._st, self._ld = _lib.lib().gi_ring_store_release, _lib.lib().gi_ring_load_acquire
        self._ctl0 = self.ctl.ctypes.data
        self.data = data.reshape(nchains, depth, M + 2)
        self.tdata = self.tdata.view(nchains, depth, M + 2)
        self.abort = False

    def _ready_addr(self, c, slot):
        return self._ctl0 + 8 * (c * self.depth + slot)

    def _done_addr(self, r, c):
        return self._ctl0 + 8 * (self.nc * self.depth + r * self.nc + c)

    def writable(self, c, k):
        """may the owner generate proposal k of chain c now?"""
        return min(int(self._ld(self._done_addr(r, c))) for r in range(self.world)) >= k - self.depth + 1

    def publish(self, c, k):
        self._st(self._ready_addr(c, k % self.depth), k + 1)  # release: after the payload

    def wait_ready(self, c, k):
        import time

        pause = 20e-6
        addr = self._ready_addr(c, k % self.depth)
        while int(self._ld(addr)) != k + 1:  # acquire: the payload is visible once the flag is
            if self.abort:
                raise RuntimeError("draw ring closed")
            time.sleep(pause)
            pause = min(pause * 1.5, 1e-3)

    def release(self, c, k):
        self._st(self._done_addr(self.rank, c), k + 1)

    def close(self):
        self.abort = True
        if self.mm is not None:
            if self.registered:
                import torch

                torch.cuda.cudart().cudaHostUnregister(self.tdata.data_ptr())
            self.ready = self.done = self.ctl = self.data = self.tdata = None
            try:
                self.mm.close()
                self.f.close()
            except (BufferError, ValueError):
                pass
            self.mm = None


class _DrawAhead:
    """Prepares the draws of upcoming proposals -- (L, p0 = randn(M)*Sigma, u) per chain, consumed
    from `RandomState(seed + c)` in the reference's order (hmc.py:297,95,165) -- on background
    threads while the GPU runs (numpy's generators release the GIL), straight into the chain's
    `_DrawRing` slots.  Chain c is served by worker (index of c among the owned chains) % nworkers,
    so every chain's stream is consumed sequentially.  Row-sharded runs: chain c is drawn by its
    owner rank only (`owned[c]`)."""

    def __init__(self, streams, Lrange, M, Sigma, ring, nworkers=None, owned=None):
        self.streams, self.Lrange, self.M, self.Sigma, self.ring = streams, Lrange, M, Sigma, ring
        self.mine = [c for c in range(len(streams)) if owned is None or owned[c]]
        self.produced = [0] * len(streams)
        self.limit = None  # proposals per chain (None: no limit)
        self.stop_flag = False
        self.error = None
        n = nworkers or max(1, min(8, (os.cpu_count() or 2) // 2))
        n = max(1, min(n, len(self.mine)))
        self.workers = [threading.Thread(target=self._run, args=(k, n), daemon=True) for k in range(n)]

    def start(self):
        for w in self.workers:
            w.start()

    def _run(self, k, n):
        import time

        ring, M = self.ring, self.M
        chains = self.mine[k::n]
        try:
            while not self.stop_flag:
                busy = False
                for c in chains:
                    kk = self.produced[c]
                    if (self.limit is not None and kk >= self.limit) or not ring.writable(c, kk):
                        continue
                    rs = self.streams[c]
                    row = ring.data[c, kk % ring.depth]
                    L = int(rs.randint(self.Lrange[0], self.Lrange[1] + 1))
                    # = rs.randn(M) * Sigma bit for bit (numpy's legacy stream continued in C,
                    # ~1.3-2x numpy's pace, written straight into the slot)
                    _lib.legacy_randn_scaled(rs, M, self.Sigma, row[:M])
                    row[M], row[M + 1] = L, float(rs.rand())
                    ring.publish(c, kk)
                    self.produced[c] = kk + 1
                    busy = True
                    if self.stop_flag:
                        return
                if not busy:
                    time.sleep(0.0005)
        except BaseException as e:  # noqa: BLE001  (a closed ring during shutdown ends the worker)
            if not self.stop_flag:
                self.error = e

    def wait_primed(self, depth=2):
        import time

        while any(self.produced[c] < min(depth, self.limit or depth) for c in self.mine):
            if self.error is not None:
                raise RuntimeError("draw worker failed") from self.error
            time.sleep(0.001)

    def stop(self):
        self.stop_flag = True
        for w in self.workers:
            w.join(timeout=30)


class _Stager(threading.Thread):
    """Moves prepared draws to the device AHEAD of the sampler, off its stream: for every request
    (a chain index, issued by the sampler's deterministic main loop) it waits for the chain's next
    slot of the `_DrawRing`, copies it to a device slot on a SIDE stream, synchronises only that
    stream and releases the ring slot.  The sampler then feeds a proposal with a device-to-device
    copy on its own stream: the host-side exchange and the host->device transfer overlap the
    contractions already queued."""

    # per chain: the handle's queue + four staged beyond it.  A call of 16 batch steps finishes up to three
    # proposals per chain and their successors are requested all at once when its records come back; with
    # only two staged beyond the queue the sampler waited ~20 ms per call for that burst to be staged
    # (0.25 ms per draw, one stager thread) -- 1.1 ms per 16 ms batch step at 8 GPUs (measured, round 2)
    SLOTS = _lib.STREAM_QUEUE_DEPTH + 4

    def __init__(self, ring, nchains, M, dev, cols=None):
        super().__init__(daemon=True)
        # peer-mode shards update their own column slice only: that is all of p0 this rank needs
        self.cols = (0, M) if cols is None else (int(cols[0]), int(cols[1]))
        import queue

        torch = _lib.require_cuda()
        self.torch, self.ring, self.nc, self.M, self.dev = torch, ring, nchains, M, dev
        self.req = queue.Queue()
        self.ready = [collections.deque() for _ in range(nchains)]
        self.cv = threading.Condition()
        self.error = None
        self.slots = torch.zeros((nchains, self.SLOTS, M + 2), dtype=torch.float64, device=dev)
        self.bytes_h2d = 0

    def request(self, c):
        self.req.put(c)

    def run(self):
        torch, ring, M = self.torch, self.ring, self.M
        try:
            torch.cuda.set_device(self.dev)
            side = torch.cuda.Stream(self.dev)
            # a sleeping wait: the cores belong to the draw workers (GI_SLEEPING_SYNC=0: spin)
            landed = torch.cuda.Event(blocking=os.environ.get("GI_SLEEPING_SYNC", "1") != "0")
            count = [0] * self.nc
            with torch.cuda.stream(side):
                while True:
                    c = self.req.get()
                    if c is None:
                        return
                    k = count[c]
                    count[c] += 1
                    buf = self.slots[c, k % self.SLOTS]
                    ring.wait_ready(c, k)
                    row = ring.data[c, k % ring.depth]
                    L, u = int(row[M]), float(row[M + 1])
                    lo, hi = self.cols
                    if hi > lo:
                        buf[lo:hi].copy_(ring.tdata[c, k % ring.depth][lo:hi], non_blocking=True)
                    landed.record(side)
                    landed.synchronize()  # the side stream only, without spinning
                    ring.release(c, k)
                    self.bytes_h2d += 8 * (hi - lo) + 16
                    with self.cv:
                        self.ready[c].append((L, u, buf[:M]))
                        self.cv.notify_all()
        except BaseException as e:  # noqa: BLE001  (surfaced by take())
            with self.cv:
                self.error = e
                self.cv.notify_all()

    def take(self, c):
        with self.cv:
            while not self.ready[c]:
                if self.error is not None:
                    raise RuntimeError("draw stager failed") from self.error
                self.cv.wait(0.05)
            return self.ready[c].popleft()

    def try_take(self, c):
        """like take(), but None when the chain's next draw is not staged yet"""
        with self.cv:
            if self.ready[c]:
                return self.ready[c].popleft()
            if self.error is not None:
                raise RuntimeError("draw stager failed") from self.error
            return None

    def stop(self):
        self.req.put(None)


class HMCBatch:
    def __init__(self, model, nchains, delta, Lrange, initial_model, aprior_model, boundaries,
                 constraint, log_factor, dobs, RegulFactor, regularization, beta, seed, Sigma,
                 save_folder="mychain", rng="numpy", quiet=False, driver="auto"):
        if constraint not in _lib.CONSTRAINTS:
            raise ValueError("Please choose right boundary constraint(mandatory, logarithmic)!")
        extra = getattr(model, "extra_regs", {})   # JointModule: "MS1" / "MStry" (potential.py:1701-1736)
        if regularization not in _lib.REG_KINDS and regularization not in extra:
            raise ValueError("Please choose regularization from 'MS','Damping', 'Smoothness', 'TV'.")
        if regularization in ("Smoothness", "TV") and getattr(model, "nocenter", False):
            raise AttributeError("'JointModule' object has no attribute 'fd3d'")  # potential.py:1753,1765
        if not 2 <= int(nchains) <= 64:
            raise ValueError("HMCBatch: 2..64 chains per batch (use hmc.HMCSample for one chain)")
        if model.wavelet:
            raise NotImplementedError("the wavelet-compressed forward is a single-chain path")
        if regularization in ("Smoothness", "TV") and int(np.prod(model.mshape)) != model.M:
            raise ValueError("Smoothness/TV are defined on the full (nz, ny, nx) grid and cannot "
                             "be used with a topography-carved model")
        self.model, self.nchains = model, int(nchains)
        self.dt, self.Lrange, self.Sigma = delta, Lrange, Sigma
        self.constraint, self.log_factor = constraint, log_factor
        self.RegulFactor, self.regularization, self.beta = RegulFactor, regularization, beta
        self.seed, self.rng, self.quiet = seed, rng, quiet
        self.save_folder = save_folder
        boundaries = np.asarray(boundaries, dtype=np.float64)
        _, WmInv, Wm = model.kernelw()
        self.wminv = WmInv.diagonal()
        self.low = Wm @ boundaries[:, 0]            # hmc.py:391-393
        self.high = Wm @ boundaries[:, 1]
        self.initial_model = Wm @ np.asarray(initial_model, dtype=np.float64)
        self.aprior_model = Wm @ np.asarray(aprior_model, dtype=np.float64)
        self.dobs = np.asarray(dobs, dtype=np.float64)
        self.streams = [np.random.RandomState(seed + c) for c in range(self.nchains)]
        self.proposals = [[] for _ in range(self.nchains)]
        self._philox_counter = 0
        self._h = None
        self._sh = None
        self._ahead = None
        self.output = "text"   # "text" (the reference's files) | "binary" | "none"  (inversion/sink.py)
        self.sink = None       # SampleSink(model, nslots=nchains): chain c -> slot c
        self.advance_cap = None  # batch steps per gi_hmcb_stream_advance call (None: the whole runway)
        _lib.require_cuda()
        mw = self.initial_model
        if constraint == "logarithmic":     # hmc.py:271-273
            x0 = (1 / log_factor) * np.log((mw - self.low) / (self.high - mw))
        else:
            x0 = mw
        self.x = np.ascontiguousarray(np.tile(x0, (self.nchains, 1)))
        self.reg = reg_params(extra.get(regularization, regularization), constraint, model.mshape, RegulFactor,
                              beta, log_factor)
        sharded = getattr(model, "world", 1) > 1
        # row-sharded batches: "device" = the C loop, exchanging through peer memory over NVLink
        # (csrc/peer.cu; GI_SHARD_EXCHANGE=nccl or driver="device-nccl": NCCL all-reduce hooks instead);
        # "host" = the all-reduce exchange driven kernel by kernel from Python (_engine.py).
        # "device-hooks" / "device-peer" force the hook / peer machinery on an unsharded model (one
        # rank is its own peer), so that single-GPU test runs exercise them
        if driver == "auto":
            driver = "device" if getattr(model.Aw_pad, "is_cuda", False) else "host"
        self._npieces = None
        if isinstance(driver, tuple):
            driver, self._npieces = driver
        self._peer, self.exchange = None, "none"
        if sharded and driver == "host":
            self._sh = _ShardedBatchState(self)
            self.exchange = "nccl"
            return
        L = _lib.lib()
        m = model
        lo, hi = m.rows if sharded else (0, m.n_total)
        cfg = _lib.HmcConfig(hi - lo, m.M, m.ld, 1 if m.fixed else 0,
                             1 if getattr(m, "nocenter", False) else 0, self.reg)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self._host = dict(dobs=f(getattr(m, "dobs_sampler", m.dobs)[lo:hi]), low=f(self.low), high=f(self.high),
                          apr=f(self.aprior_model),
                          wmsq=f(np.ones(m.M) if regularization == "MStry" else m.WmSquare.diagonal()))
        fix = f(np.asarray(m.grav_fix, dtype=np.float64)[lo:hi]) if m.fixed else None
        h = C.c_void_p()
        _lib.check(L.gi_hmcb_create(C.byref(cfg), self.nchains, _lib.ptr(m.Aw_pad),
                                    _lib.ptr(self._host["dobs"]), _lib.ptr(fix),
                                    _lib.ptr(self._host["low"]), _lib.ptr(self._host["high"]),
                                    _lib.ptr(self._host["apr"]), _lib.ptr(self._host["wmsq"]),
                                    _lib.stream_ptr(), C.byref(h)), "gi_hmcb_create")
        self._h = h
        self._peer = None
        from . import peer as _peer

        if driver == "device-peer" or (sharded and driver == "device" and _peer.exchange_mode(m) == "peer"):
            self._attach_peer(lo, hi)
        elif sharded or driver in ("device-hooks", "device-nccl"):
            self._attach_shard_hooks(lo, hi)
        self.exchange = "peer" if self._peer is not None else ("nccl" if sharded else "none")
        _lib.check(L.gi_hmcb_set_state(self._h, _lib.ptr(self.x)), "gi_hmcb_set_state")

    def _attach_peer(self, lo, hi):
        """row-sharded model on one NVLink node: the ranks map each other's memory and the kernels move
        the data themselves (include/gravinv_b200.h: gi_hmcb_set_peer)"""
        from .peer import PeerBuffer

        m, L = self.model, _lib.lib()
        world, rank = getattr(m, "world", 1), getattr(m, "rank", 0)
        nbytes = int(L.gi_hmcb_peer_bytes(self._h, world))
        self._peer = PeerBuffer(nbytes, rank, world, m.group if world > 1 else None)
        dobs_c = np.ascontiguousarray(m.dobs[lo:hi] - float(np.mean(m.dobs)))
        _lib.check(L.gi_hmcb_set_peer(self._h, self._peer.h, m.n_total, _lib.ptr(dobs_c)), "gi_hmcb_set_peer")
        if world > 1:
            import torch.distributed as dist

            _lib.sync()
            dist.barrier(group=m.group)  # every rank's buffers are zeroed and in place before the first store

    def _attach_shard_hooks(self, lo, hi):
        """row-sharded model: the device loop calls back here at its exchange points and
        torch.distributed (NCCL) sums over the ranks, ordered on the current stream"""
        import torch
        import torch.distributed as dist

        m, L = self.model, _lib.lib()
        Cp = int(L.gi_hmcb_padded_chains(self._h))
        ld = m.ld
        # 4 pieces: the all-reduce of a piece hides under the contraction of the next one and each
        # piece still fills the 148 SMs for ~7 waves (8 pieces: 3.5 waves, 14 % tail loss)
        npieces = self._npieces or next(
            (k for k in (4, 2) if ld % (256 * k) == 0 and ld // k >= 65536), 1)
        if os.environ.get("GI_NPIECES"):       # diagnostics
            npieces = int(os.environ["GI_NPIECES"])
        dev = m.Aw_pad.device
        self._gext = torch.zeros(Cp * ld, dtype=torch.float64, device=dev)
        self._red = torch.zeros(2 * Cp, dtype=torch.float64, device=dev)
        n = Cp * (ld // npieces)
        pieces = [self._gext[k * n:(k + 1) * n] for k in range(npieces)]
        red0, red1 = self._red[:Cp], self._red[Cp:]
        pending, group = [], m.group
        single = getattr(m, "world", 1) == 1  # one rank: every sum is already complete
        if os.environ.get("GI_NULL_HOOK"):     # diagnostics: measure the loop without communication
            single = True


        def hook(user, what, piece, async_):
            try:
                if single:
                    return 0
                if what == 3:
                    for w in pending:
                        w.wait()
                    pending.clear()
                    return 0
                t = red0 if what == 0 else red1 if what == 2 else pieces[piece]
                if async_:
                    pending.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True))
                else:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
                return 0
            except Exception:  # an exception must not unwind through the C frames
                import traceback

                traceback.print_exc()
                return -1

        self._hook = _lib.SHARD_HOOK(hook)  # keep the callback object alive with the handle
        dobs_c = np.ascontiguousarray(m.dobs[lo:hi] - float(np.mean(m.dobs)))
        _lib.check(L.gi_hmcb_set_shard(self._h, m.n_total, _lib.ptr(dobs_c), _lib.ptr(self._gext),
                                       npieces, _lib.ptr(self._red), self._hook, None),
                   "gi_hmcb_set_shard")

    def close(self):
        if self._h is not None:
            _lib.lib().gi_hmcb_destroy(self._h)
            self._h = None
        if getattr(self, "_peer", None) is not None:
            self._peer.close()
            self._peer = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def propose(self, active=None, trace=None):
        """One proposal on every active chain (hmc.py:297-300 for each of them).  Returns the list
        of `gi_hmc_result`-like tuples (accept, L, U, U_data, U_model) per chain (None if inactive)
        and refreshes `self.x` for accepted chains."""
        nc, M = self.nchains, self.model.M
        lib = _lib.lib()
        act = [True] * nc if active is None else list(active)
        Ls = np.zeros(nc, dtype=np.int32)
        res = (_lib.HmcResult * nc)()
        if self.rng == "philox":
            if self._sh is not None:
                raise NotImplementedError("rng='philox' is a single-GPU option")
            rs = np.random.RandomState(self.seed + 7919 * (self._philox_counter + 1))
            for c in range(nc):
                if act[c]:
                    Ls[c] = rs.randint(self.Lrange[0], self.Lrange[1] + 1)
            if trace is not None:
                raise ValueError("tracing needs injected draws (rng='numpy')")
            _lib.check(lib.gi_hmcb_propose_philox(self._h, int(self.seed), self._philox_counter,
                                                  float(self.Sigma), _lib.ptr(Ls), float(self.dt),
                                                  res), "gi_hmcb_propose_philox")
            self._philox_counter += 1
        else:
            p0 = np.zeros((nc, M))
            u = np.zeros(nc)
            for c in range(nc):
                if not act[c]:
                    continue
                rs = self.streams[c]
                Ls[c] = rs.randint(self.Lrange[0], self.Lrange[1] + 1)   # hmc.py:297
                p0[c] = rs.randn(M) * self.Sigma                          # hmc.py:95
                u[c] = rs.rand()                                          # hmc.py:165
            tx = tu = None
            if trace is not None:
                Lmax = int(Ls.max())
                tx = np.zeros((Lmax + 1, nc, M))
                tu = np.zeros((Lmax + 1, nc))
            if self._sh is not None:
                self._sh.propose(p0, Ls, u, res, tx, tu)
            else:
                _lib.check(lib.gi_hmcb_propose(self._h, _lib.ptr(p0), _lib.ptr(Ls), float(self.dt),
                                               _lib.ptr(u), res, _lib.ptr(tx), _lib.ptr(tu)),
                           "gi_hmcb_propose")
            if trace is not None:
                trace.update(x=tx, U=tu, L=Ls.copy(),
                             Hcur=np.array([r.Hcur for r in res]), Hnew=np.array([r.Hnew for r in res]))
        out = []
        any_accept = False
        for c in range(nc):
            if not act[c]:
                out.append(None)
                continue
            r = res[c]
            self.proposals[c].append((int(Ls[c]), bool(r.accept)))
            any_accept |= bool(r.accept)
            out.append((bool(r.accept), int(Ls[c]), r.U, r.U_data, r.U_model))
        if any_accept:
            if self._sh is not None:
                self.x[:] = self._sh.x_cur[:nc, :M].cpu().numpy()
            else:
                _lib.check(lib.gi_hmcb_get_state(self._h, _lib.ptr(self.x), None, None),
                           "gi_hmcb_get_state")
        return out

    def start_draws(self, wait=False, limit=None):
        """start preparing the draws of the next proposals on background threads (optional; `stream`
        does it itself).  `wait=True` returns once two proposals per chain are ready; `limit` caps
        the proposals drawn per chain."""
        world, rank = getattr(self.model, "world", 1), getattr(self.model, "rank", 0)
        nc, M = self.nchains, self.model.M
        sharded = world > 1 and self._sh is None
        self._owner = [c % world for c in range(nc)]
        owned = [o == rank for o in self._owner] if sharded else None
        if sharded:
            import tempfile

            import torch.distributed as dist

            if int(os.environ.get("LOCAL_WORLD_SIZE", world)) != world:
                raise NotImplementedError("the streaming sampler shares draws through host memory: "
                                          "all ranks must run on one node")
            # rank 0 creates the ring file, the others map it (collective: every rank gets here)
            name = [None]
            if rank == 0:
                # /dev/shm when it has room for the ring (plus slack), else the temp directory (a
                # page-cache backed file: same semantics, the kernel may write it back lazily)
                need = 8 * nc * 3 * (M + 2) + (64 << 20)
                base = None
                try:
                    st = os.statvfs("/dev/shm")
                    if st.f_bavail * st.f_frsize > need:
                        base = "/dev/shm"
                except OSError:
                    pass
                fd, name[0] = tempfile.mkstemp(prefix="gi_draws_", dir=base)
                os.close(fd)
                ring = _DrawRing(nc, M, world, rank, path=name[0], create=True)
            dist.broadcast_object_list(name, src=0, group=self.model.group)
            if rank != 0:
                ring = _DrawRing(nc, M, world, rank, path=name[0])
            dist.barrier(group=self.model.group)
            if rank == 0:
                os.unlink(name[0])  # every rank has it mapped; the memory lives until they unmap
        else:
            ring = _DrawRing(nc, M)
        # host threads for the draws: this rank's share of the cores minus one (the sampler and the stager
        # threads sleep in their waits, see sync_sleeping in csrc/batched.cu)
        nw = max(1, min(8, (os.cpu_count() or 2) // max(world, 1) - 1))
        if os.environ.get("GI_DRAW_WORKERS"):
            nw = max(1, int(os.environ["GI_DRAW_WORKERS"]))
        self._ahead = _DrawAhead(self.streams, self.Lrange, M, self.Sigma, ring, nworkers=nw, owned=owned)
        self._ahead.limit = limit
        self._ahead_workers = self._ahead.workers
        self._ahead.start()
        self._stream_buffers()
        if wait:
            self._ahead.wait_primed()
        return self._ahead

    def _stream_buffers(self):
        """record array + pinned staging for the positions that come back with every record"""
        if getattr(self, "_recs", None) is None:
            torch = _lib.require_cuda()
            cap = max(2 * self.nchains, 64)
            self._recs = (_lib.StreamRecord * cap)()
            self._xh = torch.empty((cap, self.model.M), dtype=torch.float64).pin_memory()
        return self._recs, self._xh

    def stream(self, nsamples, ndraws, max_proposals=None, write=True, on_record=None):
        """hmc.py:252-343 for every chain, streaming: each chain runs its proposals back to back and
        a chain that ends a trajectory opens the next one in the same batch step (`gi_hmcb_stream_*`),
        so no chain idles while others finish longer trajectories.  Same draws, same decisions and
        same files as `sample()`; single-GPU only."""
        if self._sh is not None:
            raise NotImplementedError("streaming needs the device driver (HMCBatch(driver='device'))")
        if self.rng != "numpy":
            raise NotImplementedError("streaming uses the host (reference-order) RNG")
        torch = _lib.require_cuda()
        lib = _lib.lib()
        nc, M = self.nchains, self.model.M
        world, rank = getattr(self.model, "world", 1), getattr(self.model, "rank", 0)
        folders = [self.save_folder + str(c) for c in range(nc)]
        write = write and rank == 0  # every rank holds the same chains
        writers = self._writers(write)
        write = writers[0].mode != "none"
        self._attach_sink(ndraws, nsamples)
        data_size, model_size = self.dobs.shape[0], self.initial_model.shape[0]
        alpha, target = self.RegulFactor, ndraws + nsamples
        recs, xh = self._stream_buffers()
        # output "none": the accepted positions stay on the device (the sink, if any, has them)
        keep_x = write or on_record is not None or self.sink is None
        cap = len(recs)
        mirror = keep_x and write  # keep self.x current per record (else it is read back once at the end)
        nrec, ndone = C.c_int32(), C.c_int32()
        count, fed, inflight, cancelled, closed = [0] * nc, [0] * nc, [0] * nc, [False] * nc, [False] * nc
        live = [True] * nc           # still needs accepted samples
        ahead = self._ahead or self.start_draws(limit=max_proposals)
        self._ahead = None
        if max_proposals is not None:
            ahead.limit = max_proposals if ahead.limit is None else min(ahead.limit, max_proposals)
        cols = None
        if self._peer is not None:
            lo_c, hi_c = C.c_int64(), C.c_int64()
            _lib.check(lib.gi_hmcb_owned_columns(self._h, C.byref(lo_c), C.byref(hi_c)), "gi_hmcb_owned_columns")
            cols = (lo_c.value, hi_c.value)
        stager = _Stager(ahead.ring, nc, M, self.model.Aw_pad.device, cols)
        stager.start()
        requested, finished = [0] * nc, [0] * nc
        depth = _lib.STREAM_QUEUE_DEPTH

        def top_up(c):
            # keep the handle's queue full and two more proposals staged beyond it.  A device slot is
            # reused SLOTS proposals later: that one is requested only after the slot's previous
            # occupant has FINISHED (its record was read back, so the device-to-device copy that fed
            # it has long completed on the sampler's stream)
            while live[c] and requested[c] < finished[c] + stager.SLOTS and \
                    (max_proposals is None or requested[c] < max_proposals):
                stager.request(c)
                requested[c] += 1

        for k in range(stager.SLOTS):  # proposal-major: every chain's first draws are staged first
            for c in range(nc):
                if requested[c] == k and (max_proposals is None or k < max_proposals):
                    stager.request(c)
                    requested[c] += 1
        _lib.check(lib.gi_hmcb_stream_begin(self._h, float(self.dt)), "gi_hmcb_stream_begin")
        space = C.c_int32()

        def feed(block):
            """hand staged draws to every live chain whose device queue has room; `block`: wait for a
            draw that is not staged yet (needed for progress before a call, not in a call's shadow)"""
            for c in range(nc):
                if not closed[c] and (not live[c] or (max_proposals is not None and fed[c] >= max_proposals)):
                    # nothing more will be fed to this chain: it may run dry without ending a call early
                    # (fed / live are the same on every rank, so the calls stay aligned across ranks)
                    _lib.check(lib.gi_hmcb_stream_close_chain(self._h, c), "gi_hmcb_stream_close_chain")
                    closed[c] = True
                while live[c] and (max_proposals is None or fed[c] < max_proposals):
                    _lib.check(lib.gi_hmcb_stream_queue_space(self._h, c, C.byref(space)),
                               "gi_hmcb_stream_queue_space")
                    if space.value <= 0:
                        break
                    if block:
                        _tw = _time.perf_counter()
                        item = stager.take(c)
                        _tw = _time.perf_counter() - _tw
                        prof["fed_before_call"] += 1
                        if _tw > 1e-4:  # the draw was not staged yet: the device idles meanwhile
                            prof["waits"] += 1
                            prof["wait_seconds"] += _tw
                    else:
                        item = stager.try_take(c)
                        prof["fed_in_shadow"] += item is not None
                    if item is None:
                        break
                    L, u, p0d = item  # staged on the device by the side stream
                    _lib.check(lib.gi_hmcb_stream_feed_dev(self._h, c, L, u, _lib.ptr(p0d)),
                               "gi_hmcb_stream_feed_dev")
                    inflight[c] += 1
                    fed[c] += 1

        self.stream_steps = 0
        import time as _time
        prof = self.stream_profile = dict(feed=0.0, advance=0.0, records=0.0, calls=0, fed_before_call=0,
                                          fed_in_shadow=0, waits=0, wait_seconds=0.0)
        self.stream_calls = []  # per call: (batch steps, records, seconds feeding before it, seconds in it)
        self.stream_mark = (0.0, 0, 0)  # (host time, batch steps, records) at the end of the call being handled
        nrec_total = 0
        try:
            while True:
                _t0 = _time.perf_counter()
                feed(True)
                run = C.c_int32()
                _lib.check(lib.gi_hmcb_stream_runway(self._h, C.byref(run)), "gi_hmcb_stream_runway")
                if run.value == 0:
                    break
                _t1 = _time.perf_counter()
                nrun = run.value if self.advance_cap is None else min(run.value, int(self.advance_cap))
                # queue the device work, then -- in its shadow -- feed the queue slots the scheduled steps
                # free (the schedule is deterministic: it does not wait for the Metropolis outcomes)
                _lib.check(lib.gi_hmcb_stream_advance_begin(self._h, nrun, cap, _lib.ptr(xh) if keep_x else None),
                           "gi_hmcb_stream_advance_begin")
                feed(False)
                _lib.check(lib.gi_hmcb_stream_advance_end(self._h, recs, cap, C.byref(nrec), C.byref(ndone)),
                           "gi_hmcb_stream_advance_end")
                self.stream_steps += ndone.value
                _t2 = _time.perf_counter()
                nrec_total += int(nrec.value)
                self.stream_mark = (_t2, self.stream_steps, nrec_total)
                self.stream_calls.append((int(ndone.value), int(nrec.value), _t1 - _t0, _t2 - _t1))
                prof["feed"] += _t1 - _t0
                prof["advance"] += _t2 - _t1
                prof["calls"] += 1
                for i in range(nrec.value):
                    r = recs[i]
                    c = r.chain
                    inflight[c] -= 1
                    finished[c] += 1
                    top_up(c)
                    if not live[c]:
                        continue  # a queued proposal that ran after the chain reached its target
                    acc = bool(r.accept)
                    self.proposals[c].append((int(r.L), acc))
                    Udn, Umn = r.U_data / data_size, r.U_model / model_size
                    Un = Udn + alpha * Umn
                    if acc:
                        if mirror:  # (an 8 MB host copy per accepted sample at c5: only when it is written)
                            self.x[c] = xh[i].numpy()
                        if count[c] >= ndraws and write:
                            x = self.x[c]
                            if self.constraint == "logarithmic":
                                mw = (self.low + self.high * np.e ** (self.log_factor * x)) / \
                                     (1 + np.e ** (self.log_factor * x))
                            else:
                                mw = x
                            writers[c].append([r.U, r.U_data, r.U_model, Un, Udn, Umn, alpha],
                                              self.wminv * mw)
                        count[c] += 1
                        if count[c] >= target:
                            live[c] = False
                    if on_record is not None:
                        on_record(c, r, acc)
                    if not self.quiet:
                        print("chain {}: {:.2%}, misfit(total, data, alpha, model)=({:.7f},{:.7f},{:.2f},"
                              "{:.7f}) -- accept ratio {:.2%}\n".format(
                                  c, count[c] / target, Un, Udn, alpha, Umn,
                                  count[c] / len(self.proposals[c])))
                        sys.stdout.flush()
                    if max_proposals is not None and len(self.proposals[c]) >= max_proposals:
                        live[c] = False
                for c in range(nc):
                    if not live[c] and inflight[c] > 0 and not cancelled[c]:
                        # the chain has its samples: what is still queued for it would only be thrown away
                        _lib.check(lib.gi_hmcb_stream_cancel(self._h, c, C.byref(space)), "gi_hmcb_stream_cancel")
                        inflight[c] -= space.value
                        cancelled[c] = True
                prof["records"] += _time.perf_counter() - _t2
        finally:
            # the stager finishes the requests already queued (the same sequence on every rank, so
            # its broadcasts pair up) before the RNG workers are stopped
            stager.stop()
            stager.join(timeout=120)
            ahead.stop()
            if world > 1:
                import torch.distributed as dist

                dist.barrier(group=self.model.group)  # nobody unmaps while another rank still reads
            ahead.ring.close()
        if not mirror:  # the device's current positions, once (a chain may have run a queued
            # proposal past its target; the recorded statistics are gated and unaffected)
            _lib.check(lib.gi_hmcb_get_state(self._h, _lib.ptr(self.x), None, None), "gi_hmcb_get_state")
        return self.x

    def _writers(self, write=True):
        from .sink import SampleWriter

        mode = self.output if write else "none"
        return [SampleWriter(self.save_folder + str(c), mode, self.model.M) for c in range(self.nchains)]

    def _attach_sink(self, ndraws, nsamples):
        if self.sink is None:
            return
        if self._h is None:
            raise NotImplementedError("the device sink needs the device driver")
        if self.sink.nslots < self.nchains:
            raise ValueError("SampleSink needs one slot per chain")
        _lib.check(_lib.lib().gi_hmcb_attach_stats(self._h, self.sink.h, _lib.ptr(self.model.wminv_dev)),
                   "gi_hmcb_attach_stats")
        if not self.sink.user_window:
            self.sink.window(ndraws, nsamples)

    def sample(self, nsamples, ndraws, max_proposals=None):
        """hmc.py:252-343 for every chain of the batch."""
        nc = self.nchains
        folders = [self.save_folder + str(c) for c in range(nc)]
        writes = getattr(self.model, "rank", 0) == 0
        writers = self._writers(writes)
        writes = writers[0].mode != "none"
        self._attach_sink(ndraws, nsamples)
        data_size, model_size = self.dobs.shape[0], self.initial_model.shape[0]
        alpha = self.RegulFactor
        count = [0] * nc      # accepted proposals (the reference's i)
        ntried = 0
        while min(count) < ndraws + nsamples:
            if max_proposals is not None and ntried >= max_proposals:
                break
            active = [count[c] < ndraws + nsamples for c in range(nc)]
            out = self.propose(active)
            ntried += 1
            for c in range(nc):
                if out[c] is None:
                    continue
                acc, L, U, Ud, Um = out[c]
                Udn, Umn = Ud / data_size, Um / model_size
                Un = Udn + alpha * Umn
                if acc:
                    if count[c] >= ndraws and writes:
                        x = self.x[c]
                        if self.constraint == "logarithmic":
                            mw = (self.low + self.high * np.e ** (self.log_factor * x)) / \
                                 (1 + np.e ** (self.log_factor * x))
                        else:
                            mw = x
                        writers[c].append([U, Ud, Um, Un, Udn, Umn, alpha], self.wminv * mw)
                    count[c] += 1
                if not self.quiet:
                    print("chain {}: {:.2%}, misfit(total, data, alpha, model)=({:.7f},{:.7f},{:.2f},"
                          "{:.7f}) -- accept ratio {:.2%}\n".format(
                              c, count[c] / (ndraws + nsamples), Un, Udn, alpha, Umn,
                              count[c] / len(self.proposals[c])))
                    sys.stdout.flush()
        return self.x


def HMCSampleBatch(model, nchains, nsamples, ndraws, delta, Lrange, initial_model, aprior_model,
                   boundaries, constraint, log_factor, dobs, adaptiveRegul, RegulRate, RegulFactor,
                   regularization, beta, seed, Sigma, nbest=100, save_folder="mychain", rng="numpy",
                   quiet=False, max_proposals=None, mode="auto"):
    """`hmc.HMCSample` for ranks 0..nchains-1 at once (argument order of hmc.py:358-361, with
    `nchains` inserted after `model` and `myrank` implied by the chain index)."""
    batch = HMCBatch(model, nchains, delta, Lrange, initial_model, aprior_model, boundaries,
                     constraint, log_factor, dobs, RegulFactor, regularization, beta, seed, Sigma,
                     save_folder=save_folder, rng=rng, quiet=quiet)
    # "stream": no chain idles (single GPU, host RNG); "lockstep": one proposal per chain per round
    if mode == "auto":
        mode = "stream" if (batch._sh is None and rng == "numpy") else "lockstep"
    if mode == "stream":
        batch.stream(nsamples, ndraws, max_proposals=max_proposals)
    else:
        batch.sample(nsamples, ndraws, max_proposals=max_proposals)
    return batch
