"""TEST INFRASTRUCTURE ONLY -- golden vectors for BASELINE.json configs 2-4 from the UNMODIFIED
reference (build container only; needs /root/reference):

    python oracle/make_golden_examples.py            # -> tests/golden/examples.npz

* c2 example/segmentgrid: segmented-z prisms (100/200/300 m) on the shipped observation file,
  Smoothness + MS chains for ranks 0 and 1 (the shipped driver's wavelet='3D' needs PyWavelets,
  absent here: the reference is run with wavelet=False; the wavelet path is checked against the
  oracle restatement in tests/test_gpu_wavelet.py).
* c3 example/realdata SC: spherical, segmented, topography-carved grid, grav_fix water-layer
  subtraction, prior model from data/SC_ApriorModel.txt, MS and Damping chains, ranks 0 and 1.
  The carve mask is the in-container scipy's (the shipped maskindex_SC.txt is
  scipy-version dependent); the fixture stores the inputs so the product recomputes everything.
* c4 example/global: 10x60x120 tesseroids, kernel rows of a few of the 7381 observations.

The inputs the product needs (observation files, topography, prior) are stored in the same npz,
because /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_harness  # noqa: E402
from oracle.make_golden import OUT, in_tmpdir, quiet  # noqa: E402

EX = os.path.join(ref_harness.REF_ROOT, "example")


def run_ranks(ns, model, dobs, nsamples, delta, Lrange, init, apr, bounds, alpha, reg, beta, seed, Sigma,
              ranks=(0, 1), max_props=25):
    """ns.hmc.HMCSample unmodified for each rank (the reference's `mpiexec -n 2`); observes the
    per-proposal (L, accept) log through a wrapper that does not alter the run."""
    out = {}
    M = model.Aw.shape[1]
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = bounds
    for rank in ranks:
        log = []
        orig = ns.hmc.HamitonianMC._leapfrog

        class Stop(Exception):
            pass

        def lf(self, xcur, dt, L, alpha_, fignum):
            if len(log) >= max_props:
                raise Stop
            r = orig(self, xcur, dt, L, alpha_, fignum)
            log.append((L, int(bool(r[3]))))
            return r

        ns.hmc.HamitonianMC._leapfrog = lf
        try:
            with quiet(), in_tmpdir() as d:
                try:
                    ns.hmc.HMCSample(model, nsamples, 0, delta, Lrange, init, apr, b, "mandatory", 1000,
                                     dobs, "Fixed", 0.8, alpha, reg, beta, seed, Sigma, myrank=rank,
                                     save_folder=os.path.join(d, "chain"))
                except Stop:
                    pass
                f = os.path.join(d, "chain%d" % rank, "misfit.dat")
                mis = np.loadtxt(f, ndmin=2) if os.path.exists(f) else np.zeros((0, 7))
                f = os.path.join(d, "chain%d" % rank, "model.dat")
                mod = np.loadtxt(f, ndmin=2) if os.path.exists(f) else np.zeros((1, M))
        finally:
            ns.hmc.HamitonianMC._leapfrog = orig
        out[rank] = dict(misfit=mis, last_model=mod[-1], log=np.array(log, dtype=np.int64))
        print(reg, "rank", rank, "proposals", len(log), "accepted", mis.shape[0], file=sys.stderr)
    return out


def main():
    if not ref_harness.available():
        raise SystemExit("reference tree not mounted")
    ns = ref_harness.load()
    out = {}
    # ------------------------------------------------------------------ c2 segmentgrid
    xo, yo, ho, go = np.loadtxt(os.path.join(EX, "segmentgrid", "modeldata", "model_seg_gz_noise.txt"),
                                usecols=[0, 1, 2, 3], unpack=True)
    out["c2_obs"], out["c2_dobs"] = np.c_[xo, yo, ho], go
    with quiet(), in_tmpdir():
        model = ns.potential.GravMagModule(go, (0, 2000, 0, 3000, 0, 2100), ([100, 200, 300], 100, 100),
                                           (xo, yo, ho), mseg=True, mdivisionsection=[0, 300, 900, 2100],
                                           coordinate="cartesian", njobs=1, field="gravity",
                                           wavelet=False)
    M = model.Aw.shape[1]
    out["c2_mshape"] = np.array(model.mshape)
    out["c2_wm"] = model.Wm.diagonal()
    rows = np.array([0, 17, 299, 450, 599])
    out["c2_rows"], out["c2_Aw_rows"] = rows, np.asarray(model.Aw)[rows]
    for reg, alpha in (("Smoothness", 1.0), ("MS", 1.0)):
        r = run_ranks(ns, model, go, 4, 0.01, [5, 20], np.ones(M) * 0.001, np.ones(M) * 0.001, (0.0, 1.0),
                      alpha, reg, 0.001, 100, 0.001)
        for rank, v in r.items():
            for k, a in v.items():
                out[f"c2_{reg}_r{rank}_{k}"] = a
    # ------------------------------------------------------------------ c3 realdata
    D = os.path.join(EX, "realdata", "data")
    lons, lats, heights, dobs = np.loadtxt(os.path.join(D, "gravinv_12d05d.dat"), usecols=[0, 1, 2, 3],
                                           unpack=True)
    grav_sea = np.loadtxt(os.path.join(D, "grasea_12d05d.dat"), usecols=[2], unpack=True)
    tl, tb, th = np.loadtxt(os.path.join(D, "topo_12d05d.dat"), usecols=[0, 1, 2], unpack=True)
    apr_mesh = np.loadtxt(os.path.join(D, "SC_ApriorModel.txt"), usecols=[3], unpack=True)
    out["c3_obs"], out["c3_dobs"], out["c3_grav_sea"] = np.c_[lons, lats, heights], dobs, grav_sea
    out["c3_topo"], out["c3_apr_mesh"] = np.c_[tl, tb, th], apr_mesh
    mrange = (106.5, 118.5, 16, 28, 2000, -60000)
    mspacing = ([-1000, -2000, -5000], 0.5, 0.5)
    with quiet(), in_tmpdir():
        model = ns.potential.GravMagModule(dobs, mrange, mspacing, (lons, lats, heights), fixed=True,
                                           grav_fix=grav_sea, mseg=True,
                                           mdivisionsection=[2000, -5000, -15000, -60000],
                                           coordinate="spherical", njobs=1, field="gravity",
                                           wavelet=False, mtopo=(tl, tb, th))
    M = model.Aw.shape[1]
    out["c3_mshape"] = np.array(model.mshape)
    out["c3_mask"] = np.array(model.mask, dtype=np.int64)
    out["c3_wm"] = model.Wm.diagonal()
    rows = np.array([0, 100, 312, 500, 624])
    out["c3_rows"], out["c3_Aw_rows"] = rows, np.asarray(model.Aw)[rows]
    nz, ny, nx = model.mshape
    init = ns.utils.rho2carve(np.ones(nz * ny * nx) * 0.01, model.mask)
    apr = ns.utils.rho2carve(apr_mesh, model.mask)
    x = model.Wm @ init
    U, g, dpre, Ud, Um = model.misfit_and_grad(x, model.Wm @ apr, None, None, "mandatory", 1000, 1,
                                               regulization="MS", beta=0.01)
    out["c3_mg_MS"] = np.array([U, Ud, Um, g[0], g[M // 2], np.linalg.norm(g)])
    out["c3_mg_dpre"] = np.asarray(dpre)
    # SetPMTS.txt T0 (Damping, delta 0.01, Sigma 0.01) and an MS chain with a step size it accepts at
    for reg, delta in (("Damping", 0.01), ("MS", 0.001)):
        r = run_ranks(ns, model, dobs, 3, delta, [5, 20], init, apr, (-0.5, 0.5), 1, reg, 0.01, 100, 0.01)
        for rank, v in r.items():
            for k, a in v.items():
                out[f"c3_{reg}_r{rank}_{k}"] = a
        out[f"c3_{reg}_delta"] = np.array(delta)
    # ------------------------------------------------------------------ c4 global (kernel rows)
    lons, lats, heights, dobs = np.loadtxt(
        os.path.join(EX, "global", "modeldata", "model_global_gz_noise.txt"), usecols=[0, 1, 2, 3],
        unpack=True)
    rows = np.array([0, 1234, 3690, 5000, 7380])
    out["c4_obs_rows"] = np.c_[lons, lats, heights][rows]
    mesh = ns.mesher.TesseroidMesh((-180, 180, -90, 90, 0, -3000000), (-300000, 3, 3))
    mesh.addprop("density", np.zeros(mesh.size))
    with quiet():
        _, K = ns.tesseroid.gz(lons[rows], lats[rows], heights[rows], mesh, njobs=1)
    out["c4_kernel_rows"] = K
    out["c4_shape"] = np.array(mesh.shape)
    np.savez_compressed(os.path.join(OUT, "examples.npz"), **out)
    print("examples.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
