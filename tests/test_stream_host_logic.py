"""Host logic of the streaming batch sampler (`HMCBatch.stream`, inversion/batched.py) on CPU: the
Python loop that feeds the device queues, cuts the run into calls, reads records back, cancels the
queue of a chain that has its samples and closes chains that will not be fed again -- driven against
a TEST DOUBLE of the `gi_hmcb_stream_*` entry points that implements the scheduler's documented
contract (include/gravinv_b200.h) in plain Python.  The real draw pipeline runs underneath
(`_DrawAhead` -> `_DrawRing` with the host RNG helper), so the test also pins that every chain's
trajectory lengths arrive in the reference's RNG order (`RandomState(seed + c).randint`, hmc.py:297).

What it guards: termination (a feed that waits for a draw nobody requested deadlocks the loop -- found
on GPUs in round 2), no feed into a full queue, exactly-once records, the sample / proposal limits,
and "no idling": the batch runs about as many steps as its busiest chain needs."""
import ctypes as C
import os
import tempfile
import types

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from gravinv3dhmc_b200 import _lib  # noqa: E402
from gravinv3dhmc_b200.inversion import batched  # noqa: E402

DEPTH = _lib.STREAM_QUEUE_DEPTH
M, SEED, LRANGE = 257, 11, [3, 9]


def accept_rule(chain, seq):
    return (chain * 7919 + seq * 104729) % 10 < 7


class FakeSchedulerLib:
    """the streaming entry points of libgravinv_b200 as a host-only state machine"""

    def __init__(self, nchains):
        self.nc = nchains
        self.q = [dict(L_cur=0, pos=0, queue=[], seq=0, closed=False) for _ in range(nchains)]
        self.pending = None
        self.fed_log = [[] for _ in range(nchains)]  # L of every proposal fed, in order
        self.max_queued = 0
        self.calls = []

    # ---- plumbing ----
    def gi_last_error(self):
        return b"fake"

    def gi_hmcb_stream_begin(self, h, dt):
        return 0

    def gi_hmcb_get_state(self, h, x, d, mw):
        return 0

    def gi_hmcb_stream_queue_space(self, h, c, space):
        space._obj.value = DEPTH - len(self.q[c]["queue"])
        return 0

    def gi_hmcb_stream_feed_dev(self, h, c, L, u, p0):
        q = self.q[c]
        if len(q["queue"]) >= DEPTH:
            return _lib.GI_ERR_BUSY
        assert not q["closed"], "fed a chain that was closed"
        q["queue"].append(int(L))
        self.fed_log[c].append(int(L))
        self.max_queued = max(self.max_queued, len(q["queue"]))
        return 0

    def gi_hmcb_stream_close_chain(self, h, c):
        self.q[c]["closed"] = True
        return 0

    def gi_hmcb_stream_cancel(self, h, c, dropped):
        assert self.pending is None
        dropped._obj.value = len(self.q[c]["queue"])
        self.q[c]["queue"].clear()
        return 0

    def gi_hmcb_stream_runway(self, h, steps):
        best, longest = -1, 0
        for q in self.q:
            if q["L_cur"] == 0 and not q["queue"]:
                continue
            rem = (q["L_cur"] - q["pos"] if q["L_cur"] > 0 else 0) + sum(q["queue"])
            longest = max(longest, rem)
            if q["closed"]:
                continue
            best = rem if best < 0 or rem < best else best
        steps._obj.value = longest if best < 0 else best
        return 0

    def gi_hmcb_stream_advance_begin(self, h, nsteps, max_records, x_host):
        assert self.pending is None
        recs, done = [], 0
        while True:
            active = [c for c in range(self.nc) if self.q[c]["L_cur"] > 0]
            fins = [c for c in active if self.q[c]["pos"] + 1 >= self.q[c]["L_cur"]]
            if active and (done >= nsteps or len(recs) + len(fins) > max_records):
                break
            starts = [c for c in range(self.nc)
                      if (c in fins or self.q[c]["L_cur"] == 0) and self.q[c]["queue"]]
            if not active and not starts:
                break
            for c in fins:
                q = self.q[c]
                recs.append((c, int(accept_rule(c, q["seq"])), q["L_cur"], q["seq"]))
            for c in active:
                if c not in fins:
                    self.q[c]["pos"] += 1
            for c in fins:
                q = self.q[c]
                q["L_cur"], q["pos"], q["seq"] = 0, 0, q["seq"] + 1
            for c in starts:
                q = self.q[c]
                q["L_cur"], q["pos"] = q["queue"].pop(0), 0
            if active:
                done += 1
        self.pending = (recs, done)
        self.calls.append(done)
        return 0

    def gi_hmcb_stream_advance_end(self, h, records, max_records, nrecords, steps_done):
        recs, done = self.pending
        self.pending = None
        for i, (c, acc, L, seq) in enumerate(recs):
            r = records[i]
            r.chain, r.accept, r.L, r.seq = c, acc, L, seq
            r.U, r.U_data, r.U_model, r.Hcur, r.Hnew = 10.0 + c, 9.0, 1.0 + seq, 0.0, 0.0
        nrecords._obj.value = len(recs)
        steps_done._obj.value = done
        return 0


class FakeStager:
    """`_Stager` without a GPU: hands out the ring's slots in request order (synchronously)"""

    SLOTS = batched._Stager.SLOTS

    def __init__(self, ring, nchains, M_, dev, cols=None):
        self.ring, self.M = ring, M_
        self.requested = [0] * nchains
        self.taken = [0] * nchains

    def start(self):
        pass

    def join(self, timeout=None):
        pass

    def stop(self):
        pass

    def request(self, c):
        self.requested[c] += 1

    def _get(self, c, block):
        k = self.taken[c]
        if k >= self.requested[c]:
            assert not block, "the sampler waits for a draw of chain %d that nobody requested (deadlock)" % c
            return None
        if not block and int(self.ring._ld(self.ring._ready_addr(c, k % self.ring.depth))) != k + 1:
            return None
        self.ring.wait_ready(c, k)
        row = self.ring.data[c, k % self.ring.depth]
        item = (int(row[self.M]), float(row[self.M + 1]), row[: self.M].copy())
        self.ring.release(c, k)
        self.taken[c] += 1
        return item

    def take(self, c):
        return self._get(c, True)

    def try_take(self, c):
        return self._get(c, False)


def make_batch(nchains, monkeypatch, tmp_path):
    fake = FakeSchedulerLib(nchains)
    real_lib = _lib.lib()
    for name in ("gi_legacy_randn_scaled", "gi_ring_store_release", "gi_ring_load_acquire"):
        setattr(fake, name, getattr(real_lib, name))  # the host helpers are real
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setattr(_lib, "require_cuda", lambda: torch)
    monkeypatch.setattr(batched, "_Stager", FakeStager)
    real_ring = batched._DrawRing

    def cpu_ring(nc, M_, world=1, rank=0, depth=3, path=None, create=False):
        fd, p = tempfile.mkstemp(prefix="gi_ring_", dir=str(tmp_path))
        os.close(fd)
        return real_ring(nc, M_, world, rank, depth=depth, path=p, create=True)  # mmap, no pinned memory

    monkeypatch.setattr(batched, "_DrawRing", cpu_ring)
    bt = object.__new__(batched.HMCBatch)
    bt._sh, bt._h, bt._peer, bt._ahead, bt.sink = None, C.c_void_p(1), None, None, None
    bt.rng, bt.quiet, bt.output, bt.advance_cap = "numpy", True, "none", None
    bt.nchains, bt.seed, bt.Lrange, bt.Sigma, bt.dt = nchains, SEED, LRANGE, 0.5, 0.01
    bt.model = types.SimpleNamespace(M=M, world=1, rank=0, group=None, Aw_pad=types.SimpleNamespace(device="cpu"))
    bt.save_folder = str(tmp_path / "chain")
    bt.dobs, bt.initial_model = np.zeros(7), np.zeros(M)
    bt.RegulFactor, bt.constraint, bt.log_factor = 1.0, "mandatory", 1000
    bt.low, bt.high, bt.wminv = np.zeros(M), np.ones(M), np.ones(M)
    bt.x = np.zeros((nchains, M))
    bt.proposals = [[] for _ in range(nchains)]
    bt.streams = [np.random.RandomState(SEED + c) for c in range(nchains)]
    cap = max(2 * nchains, 64)
    bt._recs = (_lib.StreamRecord * cap)()
    bt._xh = torch.zeros((cap, M), dtype=torch.float64)
    return bt, fake


def reference_lengths(c, n):
    """the trajectory lengths of chain c's first n proposals: RandomState(seed + c) consumed in the
    reference's order (randint, randn(M), rand per proposal; hmc.py:297,95,165)"""
    rs = np.random.RandomState(SEED + c)
    out = []
    for _ in range(n):
        out.append(int(rs.randint(LRANGE[0], LRANGE[1] + 1)))
        rs.randn(M)
        rs.rand()
    return out


@pytest.mark.timeout(120)
@pytest.mark.parametrize("cap", [None, 4])
def test_stream_until_every_chain_has_its_samples(monkeypatch, tmp_path, cap):
    nch, nsamples = 6, 9
    bt, fake = make_batch(nch, monkeypatch, tmp_path)
    bt.advance_cap = cap
    seen = []
    bt.stream(nsamples, 0, write=False, on_record=lambda c, r, acc: seen.append((c, int(r.seq), acc)))
    assert fake.max_queued <= DEPTH
    busiest = 0
    for c in range(nch):
        props = bt.proposals[c]
        # recording stops with the sample that completes the chain's quota
        assert sum(a for _, a in props) == nsamples and props[-1][1]
        assert [a for _, a in props] == [accept_rule(c, k) for k in range(len(props))]
        # the draws reached the device queue in the reference's RNG order
        assert [L for L, _ in props] == reference_lengths(c, len(props))
        assert fake.fed_log[c][: len(props)] == [L for L, _ in props]
        assert [s for cc, s, _ in seen if cc == c] == list(range(len(props)))  # exactly once, in order
        busiest = max(busiest, sum(L for L, _ in props))
        # what was queued beyond the quota is cancelled as soon as the host sees the completing record:
        # at most the proposals that were already on the device for the call in flight ran
        assert fake.q[c]["seq"] <= len(props) + DEPTH
    assert busiest <= bt.stream_steps <= busiest + LRANGE[1]  # no idling, short drain
    if cap:
        assert max(fake.calls) <= cap


@pytest.mark.timeout(120)
def test_stream_with_a_proposal_limit(monkeypatch, tmp_path):
    nch, nprop = 5, 7
    bt, fake = make_batch(nch, monkeypatch, tmp_path)
    bt.advance_cap = 16
    bt.stream(10 ** 9, 0, max_proposals=nprop, write=False)
    for c in range(nch):
        assert len(bt.proposals[c]) == nprop == len(fake.fed_log[c]) == fake.q[c]["seq"]
        assert [L for L, _ in bt.proposals[c]] == reference_lengths(c, nprop)
        assert fake.q[c]["closed"]  # nothing more to feed: the chain may run dry without ending calls early
    # chains that ran out of proposals no longer cut the calls short: far fewer calls than batch steps
    assert len(fake.calls) <= 2 + bt.stream_steps // 8
