"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the UNMODIFIED reference
(``/root/reference``, via ``oracle/ref_harness.py``) in the build container.

The reference is Python (numpy/Cython/numba) and does not travel to the GPU box, so its
outputs on small seeded inputs are committed as fixtures.  Re-run with::

    python oracle/make_golden.py            # writes tests/golden/*.npz

Every array below is an OUTPUT of the reference's own code (``gravmag.prism.gz``,
``gravmag.tesseroid.gz``, ``mesher.*Mesh*``, ``inversion.potential.GravMagModule``,
``inversion.hmc.HMCSample``) or an input it was fed.  Nothing here is imported by the product.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


@contextlib.contextmanager
def in_tmpdir():
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            yield d
        finally:
            os.chdir(cwd)


def mesh_table(mesh):
    """[size,6] bounds of every cell (NaN rows for masked cells), by iterating the reference mesh."""
    tab = np.full((mesh.size, 6), np.nan)
    for i in range(mesh.size):
        c = mesh[i]
        if c is not None:
            tab[i] = c.get_bounds()
    return tab


def synthetic_topo(x1, x2, y1, y2, amp, base, n=9):
    xs = np.linspace(x1, x2, n)
    ys = np.linspace(y1, y2, n)
    X, Y = np.meshgrid(xs, ys)
    H = base + amp * np.sin(2.1 * (X - x1) / (x2 - x1) + 0.3) * np.cos(1.7 * (Y - y1) / (y2 - y1))
    return X.ravel(), Y.ravel(), H.ravel()


# ----------------------------------------------------------------------------------------------
def gen_meshes(ns):
    out = {}
    m = ns.mesher
    with quiet(), in_tmpdir():
        cases = {
            "prism_uniform": (m.PrismMesh, ((0, 400, 0, 600, 0, 500), (100, 100, 100)), {}),
            "prism_uniform_nondiv": (m.PrismMesh, ((0, 410, -35, 600, 10, 505), (100, 90, 70)), {}),
            "prism_ratio": (m.PrismMesh, ((0, 300, 0, 300, 0, 2000), (100, 100, 100), 1.3), {}),
            "prism_segment": (m.PrismMeshSegment,
                              ((0, 400, 0, 300, 0, 2100), ([100, 200, 300], 100, 100),
                               [0, 300, 900, 2100]), {}),
            "tess_uniform": (m.TesseroidMesh, ((-10, 10, -10, 10, 0, -300000), (-100000, 5, 5)), {}),
            "tess_segment": (m.TesseroidMeshSegment,
                             ((106.5, 109.5, 16, 18, 2000, -60000), ([-1000, -2000, -5000], 0.5, 0.5),
                              [2000, -5000, -15000, -60000]), {}),
        }
        for name, (cls, args, kw) in cases.items():
            mesh = cls(*args, **kw)
            out[name + "_shape"] = np.array(mesh.shape)
            out[name + "_bounds"] = np.array(mesh.bounds, dtype=np.float64)
            out[name + "_table"] = mesh_table(mesh)
            out[name + "_xs"] = np.asarray(mesh.get_xs(), dtype=np.float64)
            out[name + "_ys"] = np.asarray(mesh.get_ys(), dtype=np.float64)
            out[name + "_zs"] = np.asarray(mesh.get_zs(), dtype=np.float64)
        # carve masks (z-down prisms: heights positive up; tesseroids z-up)
        tx, ty, th = synthetic_topo(0, 400, 0, 600, 120.0, -150.0)
        mesh = m.PrismMesh((0, 400, 0, 600, 0, 500), (100, 100, 100))
        out["carve_prism_topo"] = np.c_[tx, ty, th]
        out["carve_prism_mask"] = np.array(mesh.carvetopo(tx, ty, th), dtype=np.int64)
        tx, ty, th = synthetic_topo(0, 400, 0, 300, 250.0, -300.0)
        mesh = m.PrismMeshSegment((0, 400, 0, 300, 0, 2100), ([100, 200, 300], 100, 100),
                                  [0, 300, 900, 2100])
        out["carve_prismseg_topo"] = np.c_[tx, ty, th]
        out["carve_prismseg_mask"] = np.array(mesh.carvetopo(tx, ty, th), dtype=np.int64)
        tx, ty, th = synthetic_topo(106.4, 109.6, 15.9, 18.1, 2500.0, -1500.0, n=11)
        mesh = m.TesseroidMeshSegment((106.5, 109.5, 16, 18, 2000, -60000),
                                      ([-1000, -2000, -5000], 0.5, 0.5),
                                      [2000, -5000, -15000, -60000])
        out["carve_tessseg_topo"] = np.c_[tx, ty, th]
        out["carve_tessseg_mask"] = np.array(mesh.carvetopo(tx, ty, th), dtype=np.int64)
        out["carve_tessseg_table"] = mesh_table(mesh)
        # rho2carve / carve2rho (utils.py:714-749)
        rho = np.arange(mesh.size, dtype=np.float64) * 0.5
        rc = ns.utils.rho2carve(rho, mesh.mask)
        out["rho2carve_out"] = np.asarray(rc, dtype=np.float64)
        back = ns.utils.carve2rho(np.asarray(rc) + 1.0, np.full(mesh.size, -7.0), mesh.mask)
        out["carve2rho_out"] = np.asarray(back, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "meshes.npz"), **out)
    print("meshes.npz", len(out), "arrays")


# ----------------------------------------------------------------------------------------------
def small_prism_setup():
    """24 obs on the mesh top (cell edges/corners -> safe_log/safe_atan2 branches), 6 above."""
    xs = np.linspace(0, 400, 4)
    ys = np.linspace(0, 600, 6)
    X, Y = np.meshgrid(xs, ys)
    xp = np.concatenate([X.ravel(), [50.0, 150.0, 333.3, 410.0, -20.0, 200.0]])
    yp = np.concatenate([Y.ravel(), [50.0, 250.0, 123.4, 610.0, -30.0, 300.0]])
    zp = np.concatenate([np.zeros(24), [-1.0, -50.0, -10.0, -5.0, -0.5, -150.0]])
    return xp, yp, zp


def gen_prism(ns):
    out = {}
    with quiet(), in_tmpdir():
        # KA1: the raw Cython call on one prism
        xp = np.array([500.0, 1000.0, 0.0])
        yp = np.array([500.0, 1500.0, 0.0])
        zp = np.array([-1.0, -1.0, 0.0])
        res = np.zeros(3)
        k1 = np.zeros(3)
        ns._prism.gz(xp, yp, zp, 0.0, 100.0, 0.0, 100.0, 0.0, 100.0, 1.0, res, k1)
        out["ka1_obs"] = np.c_[xp, yp, zp]
        out["ka1_kernel1d"] = k1
        # small mesh through prism.gz
        xp, yp, zp = small_prism_setup()
        mesh = ns.mesher.PrismMesh((0, 400, 0, 600, 0, 500), (100, 100, 100))
        dens = 0.1 + 0.01 * np.arange(mesh.size)
        mesh.addprop("density", dens)
        res, K = ns.prism.gz(xp, yp, zp, mesh)
        out["small_obs"] = np.c_[xp, yp, zp]
        out["small_dens"] = dens
        out["small_result"] = res
        out["small_kernel"] = K
        # carved small mesh
        tx, ty, th = synthetic_topo(0, 400, 0, 600, 120.0, -150.0)
        mesh = ns.mesher.PrismMesh((0, 400, 0, 600, 0, 500), (100, 100, 100))
        mesh.carvetopo(tx, ty, th)
        mesh.addprop("density", np.zeros(mesh.size))
        _, K = ns.prism.gz(xp, yp, zp - 200.0, mesh)
        out["carved_obs"] = np.c_[xp, yp, zp - 200.0]
        out["carved_topo"] = np.c_[tx, ty, th]
        out["carved_mask"] = np.array(mesh.mask, dtype=np.int64)
        out["carved_kernel"] = K
        # segmented
        mesh = ns.mesher.PrismMeshSegment((0, 400, 0, 300, 0, 2100), ([100, 200, 300], 100, 100),
                                          [0, 300, 900, 2100])
        mesh.addprop("density", np.zeros(mesh.size))
        xs, ys, zs = small_prism_setup()
        _, K = ns.prism.gz(xs, ys * 0.5, zs, mesh)
        out["seg_obs"] = np.c_[xs, ys * 0.5, zs]
        out["seg_kernel"] = K
        # config 1 (example/uniformgrid): inputs + slices of the 600 x 6000 kernel (KA2, KA3)
        f = os.path.join(ref_harness.REF_ROOT, "example", "uniformgrid", "modeldata",
                         "model01_singlecube_gz_noise.txt")
        xo, yo, ho, go = np.loadtxt(f, usecols=[0, 1, 2, 3], unpack=True)
        mesh = ns.mesher.PrismMesh((0, 2000, 0, 3000, 0, 1000), (100, 100, 100))
        mesh.addprop("density", np.zeros(mesh.size))
        _, A = ns.prism.gz(xo, yo, ho, mesh)
        out["c1_obs"] = np.c_[xo, yo, ho]
        out["c1_dobs"] = go
        rows = np.array([0, 1, 37, 299, 300, 598, 599])
        out["c1_rows"] = rows
        out["c1_kernel_rows"] = A[rows]
        out["c1_kernel_colsum"] = A.sum(axis=0)
        out["c1_kernel_rowsum"] = A.sum(axis=1)
        out["c1_kernel_stats"] = np.array([A.sum(), A.max(), A.min()])
        np.save(os.path.join(tempfile.gettempdir(), "gravinv_c1_A.npy"), A)
    np.savez_compressed(os.path.join(OUT, "prism.npz"), **out)
    print("prism.npz", len(out), "arrays")


# ----------------------------------------------------------------------------------------------
def gen_tess(ns):
    out = {}
    with quiet(), in_tmpdir():
        # KA6
        mesh = ns.mesher.TesseroidMesh((-10, 10, -10, 10, 0, -300000), (-100000, 5, 5))
        mesh.addprop("density", np.zeros(mesh.size))
        l = np.linspace(-9, 9, 4)
        LON, LAT = np.meshgrid(l, l)
        lon, lat = LON.ravel(), LAT.ravel()
        h = np.full(lon.shape, 10000.0)
        _, K = ns.tesseroid.gz(lon, lat, h, mesh)
        out["ka6_obs"] = np.c_[lon, lat, h]
        out["ka6_kernel"] = K
        # near-field: observations 2 km above the mesh top -> deep subdivision
        lon = np.array([-7.5, -2.4, 0.0, 3.3, 9.9, 12.0])
        lat = np.array([-7.5, 1.1, 0.0, -4.2, 9.9, -12.0])
        h = np.array([2000.0, 2500.0, 2000.0, 5000.0, 2000.0, 3000.0])
        _, K = ns.tesseroid.gz(lon, lat, h, mesh)
        out["near_obs"] = np.c_[lon, lat, h]
        out["near_kernel"] = K
        # segmented + carved (config-3-like, small)
        tx, ty, th = synthetic_topo(106.4, 109.6, 15.9, 18.1, 2500.0, -1500.0, n=11)
        mesh = ns.mesher.TesseroidMeshSegment((106.5, 109.5, 16, 18, 2000, -60000),
                                              ([-1000, -2000, -5000], 0.5, 0.5),
                                              [2000, -5000, -15000, -60000])
        mesh.carvetopo(tx, ty, th)
        mesh.addprop("density", np.zeros(mesh.size))
        lo = np.linspace(106.75, 109.25, 6)
        la = np.linspace(16.25, 17.75, 4)
        LON, LAT = np.meshgrid(lo, la)
        lon, lat = LON.ravel(), LAT.ravel()
        h = np.full(lon.shape, 3000.0)
        _, K = ns.tesseroid.gz(lon, lat, h, mesh)
        out["segcarve_obs"] = np.c_[lon, lat, h]
        out["segcarve_topo"] = np.c_[tx, ty, th]
        out["segcarve_mask"] = np.array(mesh.mask, dtype=np.int64)
        out["segcarve_kernel"] = K
    np.savez_compressed(os.path.join(OUT, "tesseroid.npz"), **out)
    print("tesseroid.npz", len(out), "arrays")


# ----------------------------------------------------------------------------------------------
def build_small_model(ns, fixed=False):
    xp, yp, zp = small_prism_setup()
    zp = zp - 20.0
    rng = np.random.RandomState(7)
    mrange = (0, 400, 0, 600, 0, 500)
    mspacing = (100, 100, 100)
    # provisional dobs; replaced after G is known
    grav_fix = 0.05 * rng.randn(xp.size) if fixed else []
    model = ns.potential.GravMagModule(np.zeros(xp.size), mrange, mspacing, (xp, yp, zp),
                                       fixed=fixed, grav_fix=grav_fix, coordinate="cartesian",
                                       njobs=1, field="gravity", wavelet=False)
    rho = np.zeros(model.mshape)
    rho[1:3, 2:4, 1:3] = 1.0
    wm = model.Wm.diagonal()
    d = model.Aw @ (wm * rho.ravel())
    if fixed:
        d = d + grav_fix
    dobs = d + 0.02 * np.abs(d).max() * rng.randn(d.size)
    model.dobs = dobs
    return model, dobs, (xp, yp, zp), np.asarray(grav_fix, dtype=np.float64)


def run_chain(ns, model, dobs, regularization, nsamples, Lrange, delta, Sigma, alpha, beta, seed,
              bounds=(0.0, 1.0), init=0.001, apr=0.001, constraint="mandatory", max_props=400):
    """Run inversion.hmc.HMCSample unmodified; wrap only to observe (not to alter) the run."""
    M = model.Aw.shape[1]
    steps_x, steps_U, prop_log = [], [], []
    orig_mg = model.misfit_and_grad

    def mg(x, *a, **k):
        r = orig_mg(x, *a, **k)
        steps_x.append(np.array(x, dtype=np.float64).copy())
        steps_U.append(float(r[0]))
        return r

    model.misfit_and_grad = mg
    orig_lf = ns.hmc.HamitonianMC._leapfrog

    class StopChain(Exception):
        pass

    def lf(self, xcur, dt, L, alpha_, fignum):
        if len(prop_log) >= max_props:  # the reference loops until nsamples ACCEPTED proposals
            raise StopChain
        n0 = len(steps_U)
        r = orig_lf(self, xcur, dt, L, alpha_, fignum)
        prop_log.append((L, int(bool(r[3])), n0, len(steps_U)))
        return r

    ns.hmc.HamitonianMC._leapfrog = lf
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = bounds
    try:
        with quiet(), in_tmpdir() as d:
            try:
                ns.hmc.HMCSample(model, nsamples, 0, delta, Lrange, np.ones(M) * init,
                                 np.ones(M) * apr, b, constraint, 1000, dobs, "Fixed", 0.8, alpha,
                                 regularization, beta, seed, Sigma, myrank=0,
                                 save_folder=os.path.join(d, "chain"))
            except StopChain:
                print("chain stopped at max_props", regularization, constraint, file=sys.stderr)
            mf = os.path.join(d, "chain0", "misfit.dat")
            misfit = np.loadtxt(mf, ndmin=2) if os.path.exists(mf) else np.zeros((0, 7))
            mf = os.path.join(d, "chain0", "model.dat")
            models = np.loadtxt(mf, ndmin=2) if os.path.exists(mf) else np.zeros((0, M))
    finally:
        ns.hmc.HamitonianMC._leapfrog = orig_lf
        model.misfit_and_grad = orig_mg
    print("chain", regularization, constraint, "proposals", len(prop_log), "accepted",
          misfit.shape[0], file=sys.stderr)
    return dict(steps_x=np.array(steps_x), steps_U=np.array(steps_U),
                prop_log=np.array(prop_log, dtype=np.int64), misfit=misfit, models=models)


def gen_potential_hmc(ns):
    out = {}
    with quiet(), in_tmpdir():
        model, dobs, obs, _ = build_small_model(ns)
    wm = model.Wm.diagonal()
    out["small_obs"] = np.c_[obs]
    out["small_dobs"] = dobs
    out["small_Aw"] = np.asarray(model.Aw)
    out["small_wm"] = wm
    out["small_wminv"] = model.WmInv.diagonal()
    out["small_wmsq"] = model.WmSquare.diagonal()
    out["small_mshape"] = np.array(model.mshape)
    M = wm.size
    x = wm * (0.5 * np.linspace(0, 1, M))
    x0 = wm * (0.001 * np.ones(M))
    out["mg_x"], out["mg_x0"] = x, x0
    for reg in ("Damping", "MS", "Smoothness", "TV"):
        U, g, dpre, Ud, Um = model.misfit_and_grad(x, x0, None, None, "mandatory", 1000, 0.7,
                                                   regulization=reg, beta=0.001)
        out[f"mg_{reg}_scalars"] = np.array([U, Ud, Um])
        out[f"mg_{reg}_grad"] = np.asarray(g, dtype=np.float64)
        out[f"mg_{reg}_dpre"] = np.asarray(dpre, dtype=np.float64)
    out["fd3d_dense_2x3x4"] = model.fd3d((2, 3, 4)).toarray()
    # chains with per-leapfrog traces; bounds [0, 0.3] make the clamp-and-flip branch fire
    for reg, alpha, beta in (("Damping", 1.0, 0.001), ("MS", 0.5, 0.001),
                             ("Smoothness", 2.0, 0.001), ("TV", 0.05, 0.001)):
        r = run_chain(ns, model, dobs, reg, nsamples=25, Lrange=[3, 8], delta=0.02, Sigma=0.05,
                      alpha=alpha, beta=beta, seed=100, bounds=(0.0, 0.3))
        for k, v in r.items():
            out[f"chain_{reg}_{k}"] = v
        out[f"chain_{reg}_params"] = np.array([alpha, beta, 0.02, 0.05, 3, 8, 100, 25, 0.0, 0.3])
    # a chain with rejections (large step) to pin the Metropolis branch
    r = run_chain(ns, model, dobs, "Damping", nsamples=12, Lrange=[4, 9], delta=0.1, Sigma=1.0,
                  alpha=1.0, beta=0.001, seed=3, bounds=(-5.0, 5.0))
    for k, v in r.items():
        out[f"chain_reject_{k}"] = v
    out["chain_reject_params"] = np.array([1.0, 0.001, 0.1, 1.0, 4, 9, 3, 12, -5.0, 5.0])
    # fixed cells (grav_fix) variant
    with quiet(), in_tmpdir():
        modelf, dobsf, _, gfix = build_small_model(ns, fixed=True)
    out["fixed_dobs"], out["fixed_gravfix"] = dobsf, gfix
    r = run_chain(ns, modelf, dobsf, "Damping", nsamples=10, Lrange=[3, 8], delta=0.02, Sigma=0.05,
                  alpha=1.0, beta=0.001, seed=11, bounds=(0.0, 1.0))
    for k, v in r.items():
        out[f"chain_fixed_{k}"] = v
    out["chain_fixed_params"] = np.array([1.0, 0.001, 0.02, 0.05, 3, 8, 11, 10, 0.0, 1.0])
    # logarithmic constraint (hmc.py:271-273, potential.py:819-820)
    r = run_chain(ns, model, dobs, "Damping", nsamples=6, Lrange=[3, 6], delta=1e-5, Sigma=1e-4,
                  alpha=1.0, beta=0.001, seed=5, bounds=(-0.5, 1.5), init=0.3, apr=0.3,
                  constraint="logarithmic")
    for k, v in r.items():
        out[f"chain_log_{k}"] = v
    out["chain_log_params"] = np.array([1.0, 0.001, 1e-5, 1e-4, 3, 6, 5, 6, -0.5, 1.5])
    np.savez_compressed(os.path.join(OUT, "potential_hmc.npz"), **out)
    print("potential_hmc.npz", len(out), "arrays")

    # config 1: KA3/KA4/KA5 anchors from the real example data
    out = {}
    pz = np.load(os.path.join(OUT, "prism.npz"))
    xo, yo, ho = pz["c1_obs"].T
    go = pz["c1_dobs"]
    with quiet(), in_tmpdir():
        model = ns.potential.GravMagModule(go, (0, 2000, 0, 3000, 0, 1000), (100, 100, 100),
                                           (xo, yo, ho), coordinate="cartesian", njobs=1,
                                           field="gravity", wavelet=False)
    wm = model.Wm.diagonal()
    out["c1_wm"] = wm
    out["c1_Aw_rows"] = np.asarray(model.Aw)[pz["c1_rows"]]
    M = wm.size
    x = wm * (0.5 * np.linspace(0, 1, M))
    x0 = wm * (0.001 * np.ones(M))
    for reg in ("Damping", "MS", "Smoothness", "TV"):
        U, g, dpre, Ud, Um = model.misfit_and_grad(x, x0, None, None, "mandatory", 1000, 1,
                                                   regulization=reg, beta=0.001)
        out[f"c1_mg_{reg}_scalars"] = np.array([U, Ud, Um, g[0], g[3000], np.linalg.norm(g)])
        if reg == "Damping":
            out["c1_mg_dpre"] = np.asarray(dpre)
            out["c1_mg_Damping_grad"] = np.asarray(g)
    r = run_chain(ns, model, go, "Damping", nsamples=8, Lrange=[5, 20], delta=0.01, Sigma=0.001,
                  alpha=1, beta=0.001, seed=100)
    out["c1_chain_misfit"] = r["misfit"]
    out["c1_chain_prop_log"] = r["prop_log"]
    out["c1_chain_steps_U"] = r["steps_U"]
    out["c1_chain_last_model"] = r["models"][-1]
    np.savez_compressed(os.path.join(OUT, "config1.npz"), **out)
    print("config1.npz", len(out), "arrays")


def main():
    if not ref_harness.available():
        raise SystemExit("reference tree not mounted; golden vectors can only be regenerated in the "
                         "build container")
    os.makedirs(OUT, exist_ok=True)
    ns = ref_harness.load()
    which = sys.argv[1:] or ["meshes", "prism", "tess", "hmc"]
    if "meshes" in which:
        gen_meshes(ns)
    if "prism" in which:
        gen_prism(ns)
    if "tess" in which:
        gen_tess(ns)
    if "hmc" in which:
        gen_potential_hmc(ns)


if __name__ == "__main__":
    main()
