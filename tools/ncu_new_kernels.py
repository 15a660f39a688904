#!/usr/bin/env python
"""Drive the kernels added for SURVEY.md 8(f) once each at a size worth profiling, for
`ncu --set full -k regex:...` captures (profiles/r01_*_ncu_full.txt):
    cg / bootstrap loop (reginv.cu), sample sink (sink.cu), prism fields (fields.cu),
    tesseroid fields (tess_fields.cu)."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def main():
    import torch

    from gravinv3dhmc_b200 import mesher
    from gravinv3dhmc_b200.gravmag import prism, tesseroid
    from gravinv3dhmc_b200.inversion import reginv, sink

    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "cg"):
        mrange, msp, obs, rho = bench.workload_geometry("mid")
        (nz, ny, nx), _, side = bench.WORKLOADS["mid"]
        N, M = side * side, nz * ny * nx
        cg = reginv.ConjugateGradient(np.zeros(N), mrange, msp, obs, verbose=False)
        cg.dobs = cg._mod.forward_local(cg.Wm @ rho).cpu().numpy()
        init = np.full(M, 0.001)
        cg.CG(init, init, (0.0, 1.0), regularization="TV", beta=0.001, q=0.9, maxk=3)
        bs = reginv.BootStrap.__new__(reginv.BootStrap)
        bs.__dict__.update(cg.__dict__)
        bs.boundary, bs.samples, bs.maxk, bs.beta, bs.batch = (0.0, 1.0), 64, 3, 0.05, 64
        bs.BSCG(init)
        sk = sink.SampleSink(cg._mod, nslots=2)
        for k in range(3):
            sk.add(init * (1 + k), 1)
        sk.result()
    if what in ("all", "fields"):
        mesh = mesher.PrismMesh((0, 6400, 0, 6400, 0, 3200), (100, 100, 100))  # 64 x 64 x 32
        xs = np.linspace(50, 6350, 32)
        X, Y = np.meshgrid(xs, xs)
        xp, yp, zp = X.ravel(), Y.ravel(), np.full(X.size, -1.0)
        tab = mesh.bounds_table()
        for f in ("gxx", "gxy", "potential", "tf"):
            prism.assemble_field(f, xp, yp, zp, tab, vec=[0.6, 0.0, 0.8] if f == "tf" else None)
    if what in ("all", "tess"):
        mesh = mesher.TesseroidMesh((0, 60, 0, 30, 0, -100000), (-10000, 0.5, 0.5))  # 10 x 60 x 120
        l1, l2 = np.linspace(0.25, 59.75, 40), np.linspace(0.25, 29.75, 25)
        LON, LAT = np.meshgrid(l1, l2)
        tab = mesh.bounds_table()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for f in ("gzz", "gx"):
                r0 = tesseroid.field_scales(f)[0]
                tesseroid.assemble(LON.ravel(), LAT.ravel(), np.full(LON.size, 20000.0), tab, ratio=r0, field=f)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
