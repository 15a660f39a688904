"""TEST INFRASTRUCTURE ONLY -- golden vectors for SURVEY.md section 8(f3/f4): the tesseroid fields
other than gz (reference ``gravmag/tesseroid.py`` potential, geoid, gx, gy, gxx .. gzz) and the
forward-only module (``gravmag/tesseroidforward.py``), produced by the UNMODIFIED reference:

    python oracle/make_golden_tessfields.py            # -> tests/golden/tessfields.npz
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

FIELDS = ("potential", "geoid", "gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz")
TRANGE, TSPACING = (-10, 10, -10, 10, 0, -300000), (-100000, 5, 5)


def tess_obs():
    l = np.linspace(-9, 9, 4)
    LON, LAT = np.meshgrid(l, l)
    lon = np.concatenate([LON.ravel(), [0.0, 2.5, -7.3, 11.0]])
    lat = np.concatenate([LAT.ravel(), [0.0, -2.5, 4.1, -12.0]])
    h = np.concatenate([np.full(16, 10000.0), [2000.0, 500.0, 250000.0, 40000.0]])  # two close to the top
    return lon, lat, h


def main():
    ns = ref_harness.load()
    fwd = importlib.import_module("gravmag.tesseroidforward")
    out = {}
    lon, lat, h = tess_obs()
    out["obs"] = np.c_[lon, lat, h]
    with quiet(), in_tmpdir():
        mesh = ns.mesher.TesseroidMesh(TRANGE, TSPACING)
        dens = 0.2 + 0.01 * np.arange(mesh.size)
        mesh.addprop("density", dens)
        out["dens"] = dens
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for f in FIELDS:
                res, K = getattr(ns.tesseroid, f)(lon, lat, h, mesh)
                out[f + "_result"], out[f + "_kernel"] = res, K
            for f in FIELDS:
                if f == "geoid":
                    continue
                out["fwd_" + f] = getattr(fwd, f)(lon, lat, h, mesh)
            out["fwd_gz_dens"] = fwd.gz(lon, lat, h, mesh, dens=2.67)
    np.savez_compressed(os.path.join(OUT, "tessfields.npz"), **out)
    print("tessfields.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
