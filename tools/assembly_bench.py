#!/usr/bin/env python
"""Time the sensitivity-matrix assembly kernels alone.
    python tools/assembly_bench.py prism [--scale mid|c5q]   # Cartesian closed-form prism gz
    python tools/assembly_bench.py tess  [--scale c4|small]  # tesseroid GLQ + adaptive subdivision
Prints Mpairs/s; `c4` is BASELINE.json configs[3] (121x61 obs at 5 km, 10x60x120 tesseroids)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gravinv3dhmc_b200 import mesher  # noqa: E402
from gravinv3dhmc_b200.gravmag import prism, tesseroid  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("kind", choices=["prism", "tess"])
ap.add_argument("--scale", default=None)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / a.reps, out


if a.kind == "prism":
    nz, ny, nx, side = {"mid": (32, 64, 64, 64), "c5q": (64, 128, 128, 64)}[a.scale or "mid"]
    mesh = mesher.PrismMesh((0, nx * 100.0, 0, ny * 100.0, 0, nz * 100.0), (100.0, 100.0, 100.0))
    xs = np.linspace(50, nx * 100 - 50, side)
    X, Y = np.meshgrid(xs, np.linspace(50, ny * 100 - 50, side))
    tab = mesh.bounds_table()
    t, (G, M) = timeit(lambda: prism.assemble(X.ravel(), Y.ravel(), np.full(X.size, -1.0), tab))
    print("prism per-cell kernel   %d obs x %d cells: %.3f s  %.1f Mpairs/s" % (X.size, M, t, X.size * M / t / 1e6))
    t, (G2, M) = timeit(lambda: prism.assemble_grid(X.ravel(), Y.ravel(), np.full(X.size, -1.0), mesh))
    print("prism shared-node kernel %d obs x %d cells: %.3f s  %.1f Mpairs/s  bit-identical: %s"
          % (X.size, M, t, X.size * M / t / 1e6, torch.equal(G, G2)))
else:
    if (a.scale or "c4") == "c4":
        mesh = mesher.TesseroidMesh((-180, 180, -90, 90, 0, -3000000), (-300000, 3, 3))
        lons, lats = np.meshgrid(np.linspace(-180, 180, 121), np.linspace(-90, 90, 61))
        h = np.full(lons.size, 5000.0)
    else:
        mesh = mesher.TesseroidMesh((100, 112, 20, 30, 1000, -80000), (-9000, 1.0, 1.5))
        lons, lats = np.meshgrid(np.linspace(99, 113, 40), np.linspace(19, 31, 40))
        h = np.full(lons.size, 2000.0)
    tab, _ = tesseroid._check_table(mesh.bounds_table())
    t, (G, M) = timeit(lambda: tesseroid.assemble(lons.ravel(), lats.ravel(), h, tab))
    print("tess   %d obs x %d cells (shape %s): %.3f s  %.1f Mpairs/s"
          % (lons.size, M, mesh.shape, t, lons.size * M / t / 1e6))
