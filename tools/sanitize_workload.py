#!/usr/bin/env python
"""One tiny invocation of every kernel family of libgravinv_b200.so, for compute-sanitizer (SURVEY.md
section 5: the reference has no race detector; here `memcheck`, `racecheck` and `synccheck` run over
the kernels themselves):

    compute-sanitizer --tool memcheck  --error-exitcode 9 python tools/sanitize_workload.py
    compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_workload.py
    compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_workload.py

Shapes are tiny (the tools slow kernels down 10-100x); results are still checked, loosely, so that a
kernel that silently did nothing fails.  `--only a,b` restricts to named sections."""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys
import tempfile
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    import torch

    torch.cuda.set_device(0)
    from gravinv3dhmc_b200 import _lib, mesher
    from gravinv3dhmc_b200.gravmag import prism, tesseroid
    from gravinv3dhmc_b200.inversion import batched, hmc, potential, reginv, sink

    L = _lib.lib()
    s = _lib.stream_ptr()
    f64 = dict(dtype=torch.float64, device="cuda")
    rng = np.random.RandomState(0)
    done = []

    def section(name):
        ok = not only or name in only
        if ok:
            done.append(name)
        return ok

    xs, ys = np.meshgrid(np.linspace(25, 775, 6), np.linspace(25, 575, 4))
    xp, yp, zp = xs.ravel(), ys.ravel(), np.full(xs.size, -1.0)
    mrange, mspacing = (0, 800, 0, 600, 0, 400), (100, 100, 100)
    tmp = tempfile.mkdtemp()

    def model_for(**kw):
        m = potential.GravMagModule(np.zeros(xp.size), mrange, mspacing, (xp, yp, zp), verbose=False, **kw)
        rho = np.zeros(m.mshape)
        rho[1:3, 2:4, 3:6] = 1.0
        d = (m.Aw @ torch.as_tensor(m.Wm.diagonal() * rho.ravel(), device="cuda")).cpu().numpy()
        m.set_dobs(d + 0.01 * rng.randn(d.size))
        return m

    model = None
    if section("assembly") or not only:
        # structured-grid prism kernel, weighting (colsumsq / weights / scale_columns)
        model = model_for()
        assert abs(float((model.Aw ** 2).sum(0).min()) - 1.0) < 1e-12
        # per-cell prism kernel (carved table) + another prism field
        tab = model.mesh.bounds_table()[::2]
        Ad, M = prism.assemble(xp, yp, zp, tab)
        assert torch.isfinite(Ad).all() and float(Ad.abs().sum()) > 0
        Ad, M = prism.assemble_field("gzz", xp, yp, zp, tab)
        assert torch.isfinite(Ad).all()
        # tesseroid gz with subdivision + a gradient field + leaf counts
        tm = mesher.TesseroidMesh((-10, 10, -10, 10, 0, -300000), (-100000, 5, 5))
        tt, _ = tesseroid._check_table(tm.bounds_table())
        lo, la = np.meshgrid(np.linspace(-9, 9, 3), np.linspace(-9, 9, 3))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            Kd, M = tesseroid.assemble(lo.ravel(), la.ravel(), np.full(9, 2000.0), tt)
            assert torch.isfinite(Kd).all()
            lv = tesseroid.leaf_counts(lo.ravel(), la.ravel(), np.full(9, 2000.0), tt)
            assert lv.max() > 8
            Kd, M = tesseroid.assemble(lo.ravel(), la.ravel(), np.full(9, 20000.0), tt, ratio=8, field="gzz")
            assert torch.isfinite(Kd).all()
    if model is None:
        model = model_for()
    M = model.M
    b = np.zeros((M, 2))
    b[:, 1] = 0.3
    one = np.full(M, 0.001)
    dobs = model.dobs
    if section("hmc"):
        # single chain: gemv_fwd / data_misfit / gemv_adj / update (all four regularisers), Metropolis,
        # commit, logarithmic transform
        for reg in ("Damping", "MS", "Smoothness", "TV"):
            ch = hmc.HMCSample(model, 2, 0, 0.02, [2, 4], one, one, b, "mandatory", 1000, dobs, "Fixed", 0.8,
                               0.5, reg, 0.001, 3, 0.05, save_folder=os.path.join(tmp, "h" + reg), quiet=True,
                               max_proposals=4)
            assert np.isfinite(ch.x_final).all()
            ch.close()
        bl = np.zeros((M, 2))
        bl[:, 0], bl[:, 1] = -0.5, 1.5
        ch = hmc.HMCSample(model, 1, 0, 1e-5, [2, 3], np.full(M, 0.3), np.full(M, 0.3), bl, "logarithmic", 1000,
                           dobs, "Fixed", 0.8, 1.0, "Damping", 0.001, 5, 1e-4,
                           save_folder=os.path.join(tmp, "hl"), quiet=True, max_proposals=2)
        ch.close()
        ch = hmc.HMCSample(model, 1, 0, 0.02, [2, 3], one, one, b, "mandatory", 1000, dobs, "Fixed", 0.8, 0.5,
                           "Damping", 0.001, 3, 0.05, save_folder=os.path.join(tmp, "hp"), quiet=True,
                           rng="philox", max_proposals=2)
        ch.close()
    if section("fused"):
        # single-pass evaluation (TMA bulk copies, mbarriers, cross-SM hand-off), ragged strip
        n, m = 40, 148 * 12 + 8
        ld = _lib.padded_ld(m)
        A = torch.zeros((n, ld), **f64)
        A[:, :m] = torch.as_tensor(rng.standard_normal((n, m)))
        x = torch.zeros(ld, **f64)
        x[:m] = torch.as_tensor(rng.standard_normal(m))
        dc = torch.as_tensor(rng.standard_normal(n)).cuda()
        d, g = torch.zeros(n, **f64), torch.zeros(ld, **f64)
        fh = C.c_void_p()
        _lib.check(L.gi_fused_create(n, m, ld, _lib.ptr(A), s, C.byref(fh)), "gi_fused_create")
        for _ in range(3):
            _lib.check(L.gi_fused_pass(fh, _lib.ptr(x), _lib.ptr(dc), None, 1, _lib.ptr(d), _lib.ptr(g), s))
        torch.cuda.synchronize()
        dr = A[:, :m] @ x[:m]
        assert float((d - dr).abs().max()) < 1e-10
        L.gi_fused_destroy(fh)
    if section("batched"):
        # DMMA contractions (cp.async ring + mbarriers), batched update, lockstep + streaming samplers,
        # the shard-hook machinery on one rank
        for driver, reg in (("auto", "TV"), ("device-hooks", "MS")):
            bt = batched.HMCBatch(model, 3, 0.02, [2, 4], one, one, b, "mandatory", 1000, dobs, 0.5, reg,
                                  0.001, 3, 0.05, save_folder=os.path.join(tmp, "b" + reg), quiet=True,
                                  driver=driver)
            bt.propose()
            bt.stream(10 ** 6, 0, max_proposals=2, write=False)
            assert np.isfinite(bt.x).all()
            bt.close()
    if section("wavelet"):
        for kind in ("1D", "3D"):
            wm = model_for(wavelet=kind)
            ch = hmc.HMCSample(wm, 1, 0, 0.02, [2, 3], one, one, b, "mandatory", 1000, wm.dobs, "Fixed", 0.8,
                               0.5, "MS", 0.001, 3, 0.05, save_folder=os.path.join(tmp, "w" + kind),
                               quiet=True, max_proposals=2)
            assert np.isfinite(ch.x_final).all()
            ch.close()
    if section("reginv"):
        cg = reginv.ConjugateGradient(dobs, mrange, mspacing, (xp, yp, zp), verbose=False)
        for reg in ("MS", "TV"):
            out = cg.CG(one, np.zeros(M), (0.0, 1.0), regularization=reg, beta=0.001, maxk=4)
            assert np.isfinite(np.asarray(out[0])).all()
        bs = reginv.BootStrap(mrange, mspacing, (xp, yp, zp), dobs, (0.0, 1.0), samples=3, beta=0.01, maxk=3,
                              verbose=False)
        try:
            bs.BSCG(one)
        except ValueError:
            pass  # a replicate that stops early raises like the reference
    if section("sink"):
        sk = sink.SampleSink(model, nslots=2)
        for row in rng.rand(3, M):
            sk.add(row, slot=1)
        mean, std, cnt = sk.result(None)
        assert cnt == 3 and np.isfinite(mean).all()
        sk.forward(1)
    if section("peer") and hasattr(L, "gi_peer_create"):
        from gravinv3dhmc_b200.inversion import peer

        peer.selftest_single_rank(model)
    torch.cuda.synchronize()
    print("sanitize workload ok:", ",".join(done))


if __name__ == "__main__":
    main()
