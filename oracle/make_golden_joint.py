"""TEST INFRASTRUCTURE ONLY -- golden vectors for the reference's `JointModule`
(inversion/potential.py:847-1812: joint gz + total-field inversion on one prism mesh, block kernel,
`weightKDM` data / model weighting) and a short HMC chain on it, produced by the UNMODIFIED reference:

    python oracle/make_golden_joint.py            # -> tests/golden/joint.npz
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
    rng = np.random.RandomState(29)
    n = xp.size
    with quiet(), in_tmpdir():
        model = ns.potential.JointModule(np.zeros(n), np.zeros(n), mrange, mspacing, (xp, yp, zp),
                                         coordinate="cartesian", njobs=1, mangle=mangle, wavelet=False)
    M2 = model.Aw.shape[1]
    Mc = M2 // 2
    rho = np.zeros(model.mshape)
    rho[1:3, 2:4, 1:3] = 0.8
    sus = np.zeros(model.mshape)
    sus[1:3, 2:4, 1:3] = 2.0
    m_true = np.append(rho.ravel(), sus.ravel())
    d = model.forward(m_true)
    d_gz = d[:n] + 0.02 * np.abs(d[:n]).max() * rng.randn(n)
    d_tf = d[n:] + 0.02 * np.abs(d[n:]).max() * rng.randn(n)
    with quiet(), in_tmpdir():
        model = ns.potential.JointModule(d_gz, d_tf, mrange, mspacing, (xp, yp, zp), coordinate="cartesian",
                                         njobs=1, mangle=mangle, wavelet=False)
    out["obs"], out["dobs_gz"], out["dobs_tf"], out["mangle"] = np.c_[xp, yp, zp], d_gz, d_tf, np.array(mangle)
    out["Aw"] = np.asarray(model.Aw.todense() if hasattr(model.Aw, "todense") else model.Aw)
    out["wm"], out["wb"], out["dobsw"] = model.Wm.diagonal(), model.Wb.diagonal(), np.asarray(model.dobsw)
    out["forward_true"], out["m_true"] = d, m_true
    wm = out["wm"]
    x, x0 = wm * (0.5 * np.linspace(0, 1, M2)), wm * (0.001 * np.ones(M2))
    out["mg_x"], out["mg_x0"] = x, x0
    for reg in ("MS", "MS1", "MStry", "Damping"):
        U, g, dpre, Ud, Um = model.misfit_and_grad(x, x0, None, None, "mandatory", 1000, 0.7,
                                                   regulization=reg, beta=0.001)
        out["mg_%s_scalars" % reg] = np.array([U, Ud, Um])
        out["mg_%s_grad" % reg], out["mg_%s_dpre" % reg] = np.asarray(g), np.asarray(dpre)
    for reg in ("Smoothness", "TV"):
        try:
            model.misfit_and_grad(x, x0, None, None, "mandatory", 1000, 0.7, regulization=reg, beta=0.001)
            out["err_" + reg] = np.array("none")
        except Exception as e:  # noqa: BLE001
            out["err_" + reg] = np.array(type(e).__name__)
    dobs = np.append(d_gz, d_tf)
    for reg, alpha in (("Damping", 1.0), ("MS", 0.5), ("MStry", 0.5)):
        r = run_chain(ns, model, dobs, reg, nsamples=8, Lrange=[3, 8], delta=0.02, Sigma=0.05, alpha=alpha,
                      beta=0.001, seed=21, bounds=(0.0, 3.0))
        for k, v in r.items():
            out["chain_%s_%s" % (reg, k)] = v
    try:
        with quiet(), in_tmpdir():
            ns.potential.JointModule(d_gz, d_tf, (0, 10, 0, 10, 0, -1000), (-500, 5, 5), (xp, yp, zp),
                                     coordinate="spherical", njobs=1)
        out["err_spherical"] = np.array("none")
    except Exception as e:  # noqa: BLE001
        out["err_spherical"] = np.array(type(e).__name__)
    np.savez_compressed(os.path.join(OUT, "joint.npz"), **out)
    print("joint.npz", len(out), "arrays; errors:", out["err_Smoothness"], out["err_TV"], out["err_spherical"])


if __name__ == "__main__":
    main()
