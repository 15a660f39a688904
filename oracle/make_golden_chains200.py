"""TEST INFRASTRUCTURE ONLY -- 200-accepted-sample chains of the UNMODIFIED reference on BASELINE.json
configs 1-4 (build container only; needs /root/reference):

    python oracle/make_golden_chains200.py [case ...]     # -> tests/golden/chains200_<case>.npz

north_star asks for per-leapfrog positions, potentials and accept decisions over the first 200
samples.  Each case runs `inversion.hmc.HMCSample` (hmc.py:358-403 -> sample :252-343 -> _leapfrog
:85-177) with the example's shipped `SetPMTS.txt` parameters (or the regulariser BASELINE.json names
for that config) until 200 proposals are ACCEPTED, and records, through wrappers that observe and do
not alter the run:

  * `log`      [nprop, 4]  (L, accept, first, last+1 index into `U`/`x32`) per proposal
  * `U`        [ncalls]    the potential returned by every `misfit_and_grad` call (hmc.py:105,147)
  * `x32`      [ncalls,32] the position passed to that call at 32 fixed indices `idx32`
  * `misfit`   [200, 7]    the rows the reference appends to misfit.dat, UNROUNDED
  * `model_last`, `model_100` the rows it appends to model.dat (m = WmInv mw), samples 200 and 100
  * the inputs the product needs to rebuild the problem (observations, topography, prior, ...)

The two file appenders (`_save_misfit_add`, `_save_models_add`, hmc.py:241-249) are replaced by
in-memory captures: they are sinks (200 x 72 000 `%.8f` floats for c4) and take no part in the
arithmetic.  The shipped drivers of c1/c2 use wavelet='3D' (PyWavelets, absent here): those chains
run with wavelet=False -- the dense forward the wavelet path approximates.

Cases: c1_MS (as shipped), c1_Damping (BASELINE.json), c2_MS (as shipped), c2_Smoothness
(BASELINE.json), c3_Damping (as shipped, T0), c3_MS (BASELINE.json), c4_Damping (as shipped),
c4_TV (BASELINE.json) -- c4 on 256 of the 7381 observation rows x all 72 000 tesseroids.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_harness  # noqa: E402
from oracle.make_golden import OUT, in_tmpdir, quiet  # noqa: E402

EX = os.path.join(ref_harness.REF_ROOT, "example")
NSAMPLES = 200


def run_chain200(ns, model, dobs, init, apr, bounds, delta, Lrange, Sigma, alpha, reg, beta, seed=100,
                 rank=0, nsamples=NSAMPLES):
    M = model.Aw.shape[1]
    idx = np.unique(np.linspace(0, M - 1, 32).astype(np.int64))
    U, X, log, mis = [], [], [], []
    orig_mg = model.misfit_and_grad

    def mg(x, *a, **k):
        r = orig_mg(x, *a, **k)
        U.append(float(r[0]))
        X.append(np.asarray(x, dtype=np.float64)[idx].copy())
        return r

    H = ns.hmc.HamitonianMC
    orig_lf, orig_sm, orig_sf = H._leapfrog, H._save_models_add, H._save_misfit_add

    def lf(self, xcur, dt, L, alpha_, fignum):
        n0 = len(U)
        r = orig_lf(self, xcur, dt, L, alpha_, fignum)
        log.append((L, int(bool(r[3])), n0, len(U)))
        return r

    def save_misfit(self, m):
        mis.append(np.array(m[0], dtype=np.float64))

    keep = {}

    def save_models_lean(self, x):
        n = keep.get("n", 0) + 1
        keep["n"] = n
        if n == 100:
            keep["m100"] = np.array(x[0], dtype=np.float64)
        keep["last"] = np.array(x[0], dtype=np.float64)

    model.misfit_and_grad = mg
    H._leapfrog, H._save_models_add, H._save_misfit_add = lf, save_models_lean, save_misfit
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = bounds
    t0 = time.time()
    try:
        with quiet(), in_tmpdir() as d:
            ns.hmc.HMCSample(model, nsamples, 0, delta, Lrange, init, apr, b, "mandatory", 1000, dobs,
                             "Fixed", 0.8, alpha, reg, beta, seed, Sigma, myrank=rank,
                             save_folder=os.path.join(d, "chain"))
    finally:
        H._leapfrog, H._save_models_add, H._save_misfit_add = orig_lf, orig_sm, orig_sf
        model.misfit_and_grad = orig_mg
    log = np.array(log, dtype=np.int64)
    print("%s rank %d: %d proposals, %d accepted, %d evaluations, %.0f s" %
          (reg, rank, len(log), int(log[:, 1].sum()), len(U), time.time() - t0), file=sys.stderr)
    assert len(mis) == nsamples and keep["n"] == nsamples
    return dict(idx32=idx, U=np.array(U), x32=np.array(X), log=log, misfit=np.array(mis),
                model_last=keep["last"], model_100=keep["m100"],
                params=np.array([delta, Lrange[0], Lrange[1], Sigma, alpha, beta, seed, rank,
                                 bounds[0], bounds[1]], dtype=np.float64))


def case_c1(ns, reg):
    xo, yo, ho, go = np.loadtxt(os.path.join(EX, "uniformgrid", "modeldata", "model01_singlecube_gz_noise.txt"),
                                usecols=[0, 1, 2, 3], unpack=True)
    with quiet(), in_tmpdir():
        model = ns.potential.GravMagModule(go, (0, 2000, 0, 3000, 0, 1000), (100, 100, 100), (xo, yo, ho),
                                           coordinate="cartesian", njobs=1, field="gravity", wavelet=False)
    M = model.Aw.shape[1]
    # example/uniformgrid/SetPMTS.txt T1: rho 0..1, L [5,20], delta 0.01, Sigma 0.001, alpha 1, beta 0.001
    out = run_chain200(ns, model, go, np.ones(M) * 0.001, np.ones(M) * 0.001, (0.0, 1.0), 0.01, [5, 20],
                       0.001, 1, reg, 0.001)
    out.update(obs=np.c_[xo, yo, ho], dobs=go)
    return out


def case_c2(ns, reg):
    xo, yo, ho, go = np.loadtxt(os.path.join(EX, "segmentgrid", "modeldata", "model_seg_gz_noise.txt"),
                                usecols=[0, 1, 2, 3], unpack=True)
    with quiet(), in_tmpdir():
        model = ns.potential.GravMagModule(go, (0, 2000, 0, 3000, 0, 2100), ([100, 200, 300], 100, 100),
                                           (xo, yo, ho), mseg=True, mdivisionsection=[0, 300, 900, 2100],
                                           coordinate="cartesian", njobs=1, field="gravity", wavelet=False)
    M = model.Aw.shape[1]
    # example/segmentgrid/SetPMTS.txt T0 (the reference's `mpiexec -n 2` second rank for Smoothness)
    out = run_chain200(ns, model, go, np.ones(M) * 0.001, np.ones(M) * 0.001, (0.0, 1.0), 0.01, [5, 20],
                       0.001, 1, reg, 0.001, rank=0 if reg == "MS" else 1)
    out.update(obs=np.c_[xo, yo, ho], dobs=go)
    return out


def case_c3(ns, reg):
    D = os.path.join(EX, "realdata", "data")
    lons, lats, heights, dobs = np.loadtxt(os.path.join(D, "gravinv_12d05d.dat"), usecols=[0, 1, 2, 3],
                                           unpack=True)
    grav_sea = np.loadtxt(os.path.join(D, "grasea_12d05d.dat"), usecols=[2], unpack=True)
    tl, tb, th = np.loadtxt(os.path.join(D, "topo_12d05d.dat"), usecols=[0, 1, 2], unpack=True)
    apr_mesh = np.loadtxt(os.path.join(D, "SC_ApriorModel.txt"), usecols=[3], unpack=True)
    with quiet(), in_tmpdir():
        model = ns.potential.GravMagModule(dobs, (106.5, 118.5, 16, 28, 2000, -60000),
                                           ([-1000, -2000, -5000], 0.5, 0.5), (lons, lats, heights),
                                           fixed=True, grav_fix=grav_sea, mseg=True,
                                           mdivisionsection=[2000, -5000, -15000, -60000],
                                           coordinate="spherical", njobs=1, field="gravity",
                                           wavelet=False, mtopo=(tl, tb, th))
    nz, ny, nx = model.mshape
    init = ns.utils.rho2carve(np.ones(nz * ny * nx) * 0.01, model.mask)
    apr = ns.utils.rho2carve(apr_mesh, model.mask)
    # example/realdata/SetPMTS.txt T0: Damping, delta 0.01, Sigma 0.01; MS with the step it accepts at
    delta = 0.01 if reg == "Damping" else 0.001
    out = run_chain200(ns, model, dobs, init, apr, (-0.5, 0.5), delta, [5, 20], 0.01, 1, reg, 0.01,
                       rank=0 if reg == "Damping" else 1)
    out.update(mask=np.array(model.mask, dtype=np.int64))  # inputs: tests/golden/examples.npz (c3_*)
    return out


def case_c4(ns, reg):
    lons, lats, heights, dobs = np.loadtxt(
        os.path.join(EX, "global", "modeldata", "model_global_gz_noise.txt"), usecols=[0, 1, 2, 3],
        unpack=True)
    rows = np.linspace(0, lons.size - 1, 256).astype(np.int64)
    lo, la, he, do = lons[rows], lats[rows], heights[rows], dobs[rows]
    # example/global/main_global.py:22-27 reorders SetPMTS's [3, 3, -300000] to (dz, dy, dx)
    with quiet(), in_tmpdir():
        model = ns.potential.GravMagModule(do, (-180, 180, -90, 90, 0, -3000000), (-300000, 3, 3),
                                           (lo, la, he), coordinate="spherical", njobs=1, field="gravity",
                                           wavelet=False)
    M = model.Aw.shape[1]
    # example/global/SetPMTS.txt T1: rho 0..0.8, L [5,20], delta 0.005, Sigma 0.001, alpha 0.05, beta 0.01
    out = run_chain200(ns, model, do, np.ones(M) * 0.001, np.ones(M) * 0.001, (0.0, 0.8), 0.005, [5, 20],
                       0.001, 0.05, reg, 0.01)
    out.update(rows=rows, obs=np.c_[lo, la, he], dobs=do, wm32=model.Wm.diagonal()[out["idx32"]])
    return out


CASES = {"c1_MS": (case_c1, "MS"), "c1_Damping": (case_c1, "Damping"), "c2_MS": (case_c2, "MS"),
         "c2_Smoothness": (case_c2, "Smoothness"), "c3_Damping": (case_c3, "Damping"),
         "c3_MS": (case_c3, "MS"), "c4_Damping": (case_c4, "Damping"), "c4_TV": (case_c4, "TV")}


def main():
    if not ref_harness.available():
        raise SystemExit("reference tree not mounted")
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(1)  # inversion/hmc.py:18-19 pins BLAS to one thread
    except Exception:
        pass
    ns = ref_harness.load()
    for name in sys.argv[1:] or list(CASES):
        fn, reg = CASES[name]
        out = fn(ns, reg)
        np.savez_compressed(os.path.join(OUT, "chains200_%s.npz" % name), **out)
        print("chains200_%s.npz" % name, {k: v.shape for k, v in out.items()}, file=sys.stderr)


if __name__ == "__main__":
    main()
