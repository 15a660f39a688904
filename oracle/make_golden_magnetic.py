"""TEST INFRASTRUCTURE ONLY -- golden vectors for the Cartesian magnetic branch of the reference's
``GravMagModule`` (inversion/potential.py:125-149: total-field anomaly kernel ``prism.tf`` along the
regional field (inc, dec)) and a short HMC chain on it, produced by the UNMODIFIED reference:

    python oracle/make_golden_magnetic.py            # -> tests/golden/magnetic.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_harness  # noqa: E402
from oracle.make_golden import OUT, in_tmpdir, quiet, run_chain, small_prism_setup  # noqa: E402


def main():
    ns = ref_harness.load()
    out = {}
    xp, yp, zp = small_prism_setup()
    zp = zp - 20.0
    mrange, mspacing, mangle = (0, 400, 0, 600, 0, 500), (100, 100, 100), (55.0, -8.0)
    rng = np.random.RandomState(17)
    with quiet(), in_tmpdir():
        model = ns.potential.GravMagModule(np.zeros(xp.size), mrange, mspacing, (xp, yp, zp),
                                           coordinate="cartesian", njobs=1, field="magnetic",
                                           mangle=mangle, wavelet=False)
    wm = model.Wm.diagonal()
    sus = np.zeros(model.mshape)
    sus[1:3, 2:4, 1:3] = 2.0
    d = model.Aw @ (wm * sus.ravel())
    dobs = d + 0.02 * np.abs(d).max() * rng.randn(d.size)
    model.dobs = dobs
    out["obs"], out["dobs"], out["mangle"] = np.c_[xp, yp, zp], dobs, np.array(mangle)
    out["Aw"], out["wm"] = np.asarray(model.Aw), wm
    M = wm.size
    x, x0 = wm * (0.5 * np.linspace(0, 1, M)), wm * (0.001 * np.ones(M))
    out["mg_x"], out["mg_x0"] = x, x0
    U, g, dpre, Ud, Um = model.misfit_and_grad(x, x0, None, None, "mandatory", 1000, 0.7,
                                               regulization="MS", beta=0.001)
    out["mg_scalars"], out["mg_grad"], out["mg_dpre"] = np.array([U, Ud, Um]), np.asarray(g), np.asarray(dpre)
    r = run_chain(ns, model, dobs, "Damping", nsamples=8, Lrange=[3, 8], delta=0.02, Sigma=0.05, alpha=1.0,
                  beta=0.001, seed=21, bounds=(0.0, 3.0))
    for k, v in r.items():
        out["chain_" + k] = v
    np.savez_compressed(os.path.join(OUT, "magnetic.npz"), **out)
    print("magnetic.npz", len(out), "arrays; chain proposals:", len(out["chain_prop_log"]))


if __name__ == "__main__":
    main()
