"""TEST INFRASTRUCTURE ONLY -- golden vectors for SURVEY.md section 8(f1): the regularised
conjugate-gradient inversion and its bootstrap (reference ``inversion/reginv.py``), produced by
running the UNMODIFIED reference in the build container (needs /root/reference):

    python oracle/make_golden_reginv.py            # -> tests/golden/reginv.npz

* ``ConjugateGradient.CG`` (reginv.py:357-491) for the four regularisers on a small Cartesian grid,
  with a non-zero prior model (exercises the MS-gradient denominator quirk, reginv.py:283-293) and
  bounds that clip (reginv.py:434-437, 463-466);
* ``BootStrap.BSCG`` (reginv.py:715-748): row resampling with ``np.random.seed(sample)`` +
  ``np.random.choice``, per-replicate CG (reginv.py:631-713);
* a spherical (tesseroid) Damping run of ``ConjugateGradient``.

The npz stores the inputs too, because /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_harness  # noqa: E402
from oracle.make_golden import OUT, in_tmpdir, quiet  # noqa: E402

MRANGE = (0, 800, 0, 600, 0, 400)
MSPACING = (100, 100, 100)
MAXK = 14
BS_SAMPLES, BS_MAXK = 5, 9


def cart_setup():
    xs = np.linspace(50, 750, 10)
    ys = np.linspace(50, 550, 8)
    X, Y = np.meshgrid(xs, ys)
    xp, yp = X.ravel(), Y.ravel()
    zp = np.full(xp.shape, -1.0)
    return xp, yp, zp


def main():
    ns = ref_harness.load()
    reginv = importlib.import_module("inversion.reginv")
    out = {}
    rng = np.random.default_rng(2024)
    xp, yp, zp = cart_setup()
    with quiet(), in_tmpdir():
        mesh = ns.mesher.PrismMesh(MRANGE, MSPACING)
        nz, ny, nx = mesh.shape
        rho = np.zeros(mesh.shape)
        rho[1:3, 2:4, 3:6] = 1.0
        mesh.addprop("density", rho.ravel())
        d0, _ = ns.prism.gz(xp, yp, zp, mesh)
        dobs = d0 + rng.normal(0.0, 0.02 * np.abs(d0).max(), d0.shape)
        out["obs"] = np.c_[xp, yp, zp]
        out["dobs"] = dobs
        out["rho_true"] = rho.ravel()
        M = mesh.size
        initial = np.full(M, 0.01)
        aprior = 0.05 + 0.02 * np.sin(np.arange(M) * 0.37)
        out["initial"], out["aprior"] = initial, aprior
        boundary = (0.0, 0.6)  # clips the 1 g/cm3 block
        out["boundary"] = np.array(boundary)
        cg = reginv.ConjugateGradient(dobs, MRANGE, MSPACING, (xp, yp, zp))
        out["cg_wm"] = cg.Wm.diagonal()
        out["cg_Aw_rows"] = np.asarray(cg.Aw[[0, 17, 79]])
        for reg, beta in (("Damping", 0.01), ("MS", 0.003), ("Smoothness", 0.01), ("TV", 0.002)):
            m, dinv, dm, mm, rf = cg.CG(initial, aprior, boundary, regularization=reg, beta=beta,
                                        q=0.9, maxk=MAXK)
            out["cg_%s_beta" % reg] = np.array(beta)
            out["cg_%s_model" % reg] = np.asarray(m)
            out["cg_%s_data" % reg] = np.asarray(dinv)
            out["cg_%s_data_misfit" % reg] = np.asarray(dm)
            out["cg_%s_model_misfit" % reg] = np.asarray(mm)
            out["cg_%s_regul" % reg] = np.asarray(rf)
        # early stop (reginv.py:486-488): weak anomaly, the normed data error drops below 0.001
        d_small = 0.04 * dobs
        cg2 = reginv.ConjugateGradient(d_small, MRANGE, MSPACING, (xp, yp, zp))
        m, dinv, dm, mm, rf = cg2.CG(initial, np.zeros(M), (-5.0, 5.0), regularization="Damping",
                                     beta=0.01, q=0.5, maxk=50)
        out["cgstop_dobs"] = d_small
        out["cgstop_model"], out["cgstop_data"] = np.asarray(m), np.asarray(dinv)
        out["cgstop_data_misfit"], out["cgstop_model_misfit"] = np.asarray(dm), np.asarray(mm)
        out["cgstop_regul"] = np.asarray(rf)
        # bootstrap
        bs = reginv.BootStrap(MRANGE, MSPACING, (xp, yp, zp), dobs, boundary, samples=BS_SAMPLES,
                              beta=0.05, maxk=BS_MAXK)
        mi, dmi, mmi, rfi = bs.BSCG(initial)
        out["bs_beta"] = np.array(0.05)
        out["bs_models"], out["bs_data_misfit"] = mi, dmi
        out["bs_model_misfit"], out["bs_regul"] = mmi, rfi
        idx = []
        for s in range(BS_SAMPLES):
            np.random.seed(s)
            idx.append(np.random.choice(np.arange(dobs.size), size=dobs.size, replace=True, p=None))
        out["bs_index"] = np.array(idx)
        # a replicate that stops early (reginv.py:693-696) leaves short lists and BSCG's row
        # assignment (reginv.py:744-746) raises ValueError
        bs2 = reginv.BootStrap(MRANGE, MSPACING, (xp, yp, zp), 0.02 * dobs, (-5.0, 5.0), samples=2,
                               beta=0.05, maxk=6)
        try:
            bs2.BSCG(np.zeros(M))
            out["bs_stop_raises"] = np.array(0)
        except ValueError:
            out["bs_stop_raises"] = np.array(1)
        # spherical
        trange = (-10, 10, -10, 10, 0, -300000)
        tspacing = (-100000, 5, 5)
        l = np.linspace(-9, 9, 5)
        LON, LAT = np.meshgrid(l, l)
        lon, lat = LON.ravel(), LAT.ravel()
        h = np.full(lon.shape, 10000.0)
        tm = ns.mesher.TesseroidMesh(trange, tspacing)
        trho = np.zeros(tm.shape)
        trho[0:2, 1:3, 1:3] = 0.5
        tm.addprop("density", trho.ravel())
        td, _ = ns.tesseroid.gz(lon, lat, h, tm)
        tdobs = td + rng.normal(0.0, 0.01 * np.abs(td).max(), td.shape)
        out["t_obs"], out["t_dobs"] = np.c_[lon, lat, h], tdobs
        tcg = reginv.ConjugateGradient(tdobs, trange, tspacing, (lon, lat, h), coordinate="spherical")
        m, dinv, dm, mm, rf = tcg.CG(np.full(tm.size, 0.001), np.zeros(tm.size), (0.0, 0.4),
                                     regularization="Damping", beta=0.01, q=0.9, maxk=10)
        out["t_model"], out["t_data"] = np.asarray(m), np.asarray(dinv)
        out["t_data_misfit"], out["t_model_misfit"], out["t_regul"] = (np.asarray(dm), np.asarray(mm),
                                                                       np.asarray(rf))
    np.savez_compressed(os.path.join(OUT, "reginv.npz"), **out)
    print("reginv.npz", len(out), "arrays;",
          "cg iters:", {r: len(out["cg_%s_regul" % r]) for r in ("Damping", "MS", "Smoothness", "TV")},
          "stop iters:", len(out["cgstop_regul"]))


if __name__ == "__main__":
    main()
