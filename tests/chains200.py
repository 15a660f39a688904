"""Shared by the CPU (oracle pinning) and GPU (product parity) tests of the 200-sample reference chains
in tests/golden/chains200_<case>.npz (made by oracle/make_golden_chains200.py from the UNMODIFIED
reference: `HMCSample` until 200 accepted samples on BASELINE.json configs 1-4, with the per-leapfrog
potential and the position at 32 fixed indices recorded for every `misfit_and_grad` call)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["c1_MS", "c1_Damping", "c2_MS", "c2_Smoothness", "c3_Damping", "c3_MS", "c4_Damping", "c4_TV"]

C3_RANGE = (106.5, 118.5, 16, 28, 2000, -60000)
C3_SPACING = ([-1000, -2000, -5000], 0.5, 0.5)
C3_DIV = [2000, -5000, -15000, -60000]


def load(case):
    return np.load(os.path.join(GOLDEN, "chains200_%s.npz" % case))


def params(g):
    delta, L0, L1, Sigma, alpha, beta, seed, rank, lo, hi = g["params"]
    return dict(delta=float(delta), Lrange=[int(L0), int(L1)], Sigma=float(Sigma), alpha=float(alpha),
                beta=float(beta), seed=int(seed), rank=int(rank), bounds=(float(lo), float(hi)))


def geometry(case, g):
    """constructor arguments of `GravMagModule` for the case (inputs stored with the golden chains;
    c3's shipped data files live in tests/golden/examples.npz)"""
    cfg = case.split("_")[0]
    if cfg == "c1":   # example/uniformgrid/main_uniform.py
        o = g["obs"]
        return dict(dobs=g["dobs"], mrange=(0, 2000, 0, 3000, 0, 1000), mspacing=(100, 100, 100),
                    obs=(o[:, 0].copy(), o[:, 1].copy(), o[:, 2].copy()), kw=dict(coordinate="cartesian"))
    if cfg == "c2":   # example/segmentgrid/main_seg.py
        o = g["obs"]
        return dict(dobs=g["dobs"], mrange=(0, 2000, 0, 3000, 0, 2100), mspacing=([100, 200, 300], 100, 100),
                    obs=(o[:, 0].copy(), o[:, 1].copy(), o[:, 2].copy()),
                    kw=dict(mseg=True, mdivisionsection=[0, 300, 900, 2100], coordinate="cartesian"))
    if cfg == "c3":   # example/realdata/main_SC.py
        e = np.load(os.path.join(GOLDEN, "examples.npz"))
        o, t = e["c3_obs"], e["c3_topo"]
        return dict(dobs=e["c3_dobs"], mrange=C3_RANGE, mspacing=C3_SPACING,
                    obs=(o[:, 0].copy(), o[:, 1].copy(), o[:, 2].copy()),
                    kw=dict(fixed=True, grav_fix=e["c3_grav_sea"], mseg=True, mdivisionsection=C3_DIV,
                            coordinate="spherical", mtopo=(t[:, 0].copy(), t[:, 1].copy(), t[:, 2].copy())),
                    apr_mesh=e["c3_apr_mesh"])
    if cfg == "c4":   # example/global/main_global.py on 256 of the 7381 observation rows
        o = g["obs"]
        return dict(dobs=g["dobs"], mrange=(-180, 180, -90, 90, 0, -3000000), mspacing=(-300000, 3, 3),
                    obs=(o[:, 0].copy(), o[:, 1].copy(), o[:, 2].copy()), kw=dict(coordinate="spherical"))
    raise KeyError(case)


def start_models(case, geo, M, mask=None, mesh_size=None, rho2carve=None):
    """(initial_model, aprior_model) in density units, as the example drivers set them"""
    if case.startswith("c3"):
        init = rho2carve(np.ones(mesh_size) * 0.01, mask)
        apr = rho2carve(geo["apr_mesh"], mask)
        return init, apr
    return np.ones(M) * 0.001, np.ones(M) * 0.001


def oracle_problem(case, g, threads=8):
    """(OracleModel, init, apr) for the case, assembled by the CPU oracle"""
    from oracle import oracle_np as onp

    geo = geometry(case, g)
    cfg = case.split("_")[0]
    lo, la, he = geo["obs"]
    mask = None
    if cfg in ("c1", "c2"):
        mesh = onp.OracleMesh(geo["mrange"], geo["mspacing"],
                              divisionsection=geo["kw"].get("mdivisionsection"))
        tab, _ = mesh.active_bounds()
        _, A = onp.prism_gz(lo, la, he, tab, threads=threads)
    elif cfg == "c3":
        mesh = onp.OracleMesh(geo["mrange"], geo["mspacing"], divisionsection=C3_DIV, zdown=False)
        t = geo["kw"]["mtopo"]
        mask = mesh.carvetopo(t[0], t[1], t[2])
        assert np.array_equal(np.array(mask), g["mask"])
        tab, _ = mesh.active_bounds()
        A, _ = onp.tess_gz(lo, la, he, tab, threads=threads)
    else:
        mesh = onp.OracleMesh(geo["mrange"], geo["mspacing"], zdown=False)
        tab = uniform_tess_table(mesh)
        A, _ = onp.tess_gz(lo, la, he, tab, threads=threads)
    Aw, wm, _, _ = onp.sensitivity_weighting(A)
    kw = geo["kw"]
    om = onp.OracleModel(Aw, wm, geo["dobs"], mesh.shape, fixed=kw.get("fixed", False),
                         grav_fix=kw.get("grav_fix"))
    init, apr = start_models(case, geo, wm.size, mask, mesh.size, onp.rho2carve)
    return om, init, apr


def uniform_tess_table(mesh):
    """vectorised bounds table of a uniform OracleMesh (the per-cell loop over 72 000 cells is slow);
    same expressions as mesher/mesh.py:229-270"""
    nz, ny, nx = mesh.shape
    x1 = np.array([mesh.bounds[0] + mesh.dims[0] * i for i in range(nx)])
    y1 = np.array([mesh.bounds[2] + mesh.dims[1] * j for j in range(ny)])
    z1 = np.array([mesh.bounds[4] + mesh.dims[2] * k for k in range(nz)])
    z2 = np.array([z1[k] + mesh.dims[2] if k < nz - 1 else mesh.bounds[5] for k in range(nz)])
    tab = np.empty((nz, ny, nx, 6))
    tab[..., 0], tab[..., 1] = x1[None, None, :], (x1 + mesh.dims[0])[None, None, :]
    tab[..., 2], tab[..., 3] = y1[None, :, None], (y1 + mesh.dims[1])[None, :, None]
    tab[..., 4], tab[..., 5] = z1[:, None, None], z2[:, None, None]
    tab = tab.reshape(-1, 6)
    for idx in (0, nx * ny - 1, tab.shape[0] // 2, tab.shape[0] - 1):
        assert tuple(tab[idx]) == mesh.cell(idx)
    return tab


def compare(g, log, U, x32, nprops=None, tol=1e-9):
    """identical (L, accept) decisions; per-leapfrog potential and positions to `tol` relative
    (positions normwise per evaluation over the 32 recorded indices) -- north_star's chain bar"""
    ref_log = g["log"][: nprops] if nprops else g["log"]
    assert [(int(L), int(a)) for L, a in log] == [(int(L), int(a)) for L, a in ref_log[:, :2]]
    ncalls = int(ref_log[-1, 3])
    ref_U, ref_x = g["U"][:ncalls], g["x32"][:ncalls]
    U, x32 = np.asarray(U), np.asarray(x32)
    assert U.shape == ref_U.shape and x32.shape == ref_x.shape, (U.shape, ref_U.shape, x32.shape)
    eU = np.max(np.abs(U - ref_U) / np.abs(ref_U))
    scale = np.max(np.abs(ref_x), axis=1, keepdims=True)
    ex = np.max(np.abs(x32 - ref_x) / scale)
    assert eU < tol and ex < tol, (eU, ex)
    return eU, ex
