"""TEST INFRASTRUCTURE ONLY -- golden vectors for SURVEY.md section 8(f3): the other prism fields
(reference ``gravmag/prism.py`` potential, geoid, gx, gy, gz, gxx .. gzz, tf, bx, by, bz over the
compiled ``gravmag/_prism.pyx``), produced by the UNMODIFIED reference in the build container:

    python oracle/make_golden_fields.py            # -> tests/golden/fields.npz

Observation points include mesh nodes, edges and faces, inside and on top of the mesh, so every
``safe_log`` / ``safe_atan2`` branch and the displaced radius of gxy / gxz / gyz
(_prism.pyx:345-350, 380-385, 442-447) are exercised.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_harness  # noqa: E402
from oracle.make_golden import OUT, in_tmpdir, quiet  # noqa: E402

MRANGE, MSPACING = (0, 400, 0, 600, 0, 500), (100, 100, 100)
GRAV = ("potential", "geoid", "gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz")


def field_obs():
    xs = np.linspace(0, 400, 5)
    ys = np.linspace(0, 600, 4)
    X, Y = np.meshgrid(xs, ys)
    top = np.c_[X.ravel(), Y.ravel(), np.zeros(X.size)]                 # nodes of the top face
    inside = np.array([[100.0, 200.0, 250.0], [200.0, 300.0, 300.0], [300.0, 100.0, 400.0],
                       [150.0, 200.0, 200.0], [100.0, 250.0, 100.0], [123.4, 456.7, 500.0],
                       [0.0, 0.0, 500.0], [400.0, 600.0, 250.0]])       # nodes / edges / faces inside
    above = np.array([[50.0, 50.0, -1.0], [333.3, 123.4, -10.0], [410.0, 610.0, -5.0],
                      [-20.0, -30.0, -0.5], [200.0, 300.0, -150.0], [100.0, 200.0, -50.0]])
    o = np.vstack([top, inside, above])
    return o[:, 0].copy(), o[:, 1].copy(), o[:, 2].copy()


def main():
    ns = ref_harness.load()
    out = {}
    xp, yp, zp = field_obs()
    out["obs"] = np.c_[xp, yp, zp]
    with quiet(), in_tmpdir():
        mesh = ns.mesher.PrismMesh(MRANGE, MSPACING)
        dens = 0.1 + 0.01 * np.arange(mesh.size)
        mesh.addprop("density", dens)
        out["dens"] = dens
        with np.errstate(all="ignore"):
            for f in GRAV:
                res, K = getattr(ns.prism, f)(xp, yp, zp, mesh)
                out[f + "_result"], out[f + "_kernel"] = res, K
            inc, dec = 52.0, -13.0
            mag = ns.utils.ang2vec(1.5 + 0.01 * np.arange(mesh.size), 40.0, 25.0)  # remanent: not along f
            mesh.addprop("magnetization", mag)
            out["inc_dec"] = np.array([inc, dec])
            out["mag"] = np.asarray(mag)
            res, K = ns.prism.tf(xp, yp, zp, mesh, inc, dec)
            out["tf_result"], out["tf_kernel"] = res, K
            res2, K2 = ns.prism.tf(xp, yp, zp, mesh, inc, dec, pmag=2.5)  # induced, scalar intensity
            out["tf_scalar_result"], out["tf_scalar_kernel"] = res2, K2
            out["bx_result"] = ns.prism._bx(xp, yp, zp, mesh)
            out["by_result"] = ns.prism._by(xp, yp, zp, mesh)
            out["bz_result"] = ns.prism._bz(xp, yp, zp, mesh)
            out["bx_pmag_result"] = ns.prism._bx(xp, yp, zp, mesh, pmag=[0.3, -1.2, 2.0])
        # carved mesh: only the active prisms get a column
        mesh2 = ns.mesher.PrismMesh(MRANGE, MSPACING)
        from oracle.make_golden import synthetic_topo
        tx, ty, th = synthetic_topo(0, 400, 0, 600, 120.0, -150.0)
        mesh2.carvetopo(tx, ty, th)
        mesh2.addprop("density", np.zeros(mesh2.size))
        out["carved_topo"] = np.c_[tx, ty, th]
        with np.errstate(all="ignore"):
            _, K = ns.prism.gzz(xp, yp, zp - 200.0, mesh2)
        out["carved_gzz_kernel"] = K
    np.savez_compressed(os.path.join(OUT, "fields.npz"), **out)
    bad = {k: int((~np.isfinite(v)).sum()) for k, v in out.items() if not np.isfinite(v).all()}
    print("fields.npz", len(out), "arrays; non-finite entries:", bad)


if __name__ == "__main__":
    main()
