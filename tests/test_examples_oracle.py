"""Oracle pinning on BASELINE.json configs 2-4 (CPU): the oracle restatement reproduces what the
UNMODIFIED reference produced on the shipped example data (tests/golden/examples.npz, made by
oracle/make_golden_examples.py) -- meshes, carve mask, kernel rows, weights, misfit_and_grad and the
first samples of the 2-rank chains."""
import numpy as np
import pytest

from oracle import oracle_np as onp

C3_RANGE = (106.5, 118.5, 16, 28, 2000, -60000)
C3_SPACING = ([-1000, -2000, -5000], 0.5, 0.5)
C3_DIV = [2000, -5000, -15000, -60000]


def normwise(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


def check_chain(g, key, om, M, nsamples, delta, init, apr, bounds, alpha, reg, beta, Sigma):
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = bounds
    for rank in (0, 1):
        log = g[f"{key}_{reg}_r{rank}_log"]
        ref = onp.hmc_sample(om, nsamples, 0, delta, [5, 20], init, apr, b, "mandatory", 1000, alpha,
                             reg, beta, 100, Sigma, myrank=rank, max_proposals=len(log))
        assert [(L, int(a)) for L, a in ref["log"]] == [(int(L), int(a)) for L, a in log]
        mis = g[f"{key}_{reg}_r{rank}_misfit"]
        assert np.allclose(ref["misfit"], mis, rtol=1e-8, atol=2e-8)
        assert np.allclose(ref["models"][-1], g[f"{key}_{reg}_r{rank}_last_model"], rtol=0, atol=2e-8)


def test_c2_segmentgrid(golden):
    g = golden["examples"]
    o = g["c2_obs"]
    mesh = onp.OracleMesh((0, 2000, 0, 3000, 0, 2100), ([100, 200, 300], 100, 100),
                          divisionsection=[0, 300, 900, 2100])
    assert mesh.shape == tuple(g["c2_mshape"]) == (10, 30, 20)
    tab, _ = mesh.active_bounds()
    _, A = onp.prism_gz(o[:, 0], o[:, 1], o[:, 2], tab, threads=4)
    Aw, wm, _, _ = onp.sensitivity_weighting(A)
    assert np.allclose(wm, g["c2_wm"], rtol=1e-12)
    assert normwise(Aw[g["c2_rows"]], g["c2_Aw_rows"]) < 1e-10
    om = onp.OracleModel(Aw, wm, g["c2_dobs"], mesh.shape)
    M = wm.size
    for reg in ("Smoothness", "MS"):
        check_chain(g, "c2", om, M, 4, 0.01, np.ones(M) * 0.001, np.ones(M) * 0.001, (0.0, 1.0), 1.0,
                    reg, 0.001, 0.001)


def test_c3_realdata(golden):
    g = golden["examples"]
    o, t = g["c3_obs"], g["c3_topo"]
    mesh = onp.OracleMesh(C3_RANGE, C3_SPACING, divisionsection=C3_DIV, zdown=False)
    assert mesh.shape == tuple(g["c3_mshape"]) == (21, 24, 24)
    mask = mesh.carvetopo(t[:, 0], t[:, 1], t[:, 2])
    assert np.array_equal(np.array(mask), g["c3_mask"])  # bit-exact bookkeeping
    tab, _ = mesh.active_bounds()
    assert tab.shape[0] == mesh.size - len(mask) == g["c3_wm"].size
    A, err = onp.tess_gz(o[:, 0], o[:, 1], o[:, 2], tab, threads=4)
    Aw, wm, _, _ = onp.sensitivity_weighting(A)
    assert np.allclose(wm, g["c3_wm"], rtol=1e-11)
    assert normwise(Aw[g["c3_rows"]], g["c3_Aw_rows"]) < 1e-10
    om = onp.OracleModel(Aw, wm, g["c3_dobs"], mesh.shape, fixed=True, grav_fix=g["c3_grav_sea"])
    M = wm.size
    init = onp.rho2carve(np.ones(mesh.size) * 0.01, mask)
    apr = onp.rho2carve(g["c3_apr_mesh"], mask)
    U, gr, dpre, Ud, Um = om.misfit_and_grad(wm * init, wm * apr, None, None, "mandatory", 1000, 1,
                                             regulization="MS", beta=0.01)
    assert np.allclose([U, Ud, Um, gr[0], gr[M // 2], np.linalg.norm(gr)], g["c3_mg_MS"], rtol=1e-9)
    assert normwise(dpre, g["c3_mg_dpre"]) < 1e-10
    for reg in ("Damping", "MS"):
        check_chain(g, "c3", om, M, 3, float(g[f"c3_{reg}_delta"]), init, apr, (-0.5, 0.5), 1, reg,
                    0.01, 0.01)


def test_c4_global_kernel_rows(golden):
    g = golden["examples"]
    o = g["c4_obs_rows"]
    mesh = onp.OracleMesh((-180, 180, -90, 90, 0, -3000000), (-300000, 3, 3), zdown=False)
    assert mesh.shape == tuple(g["c4_shape"]) == (10, 60, 120)
    nz, ny, nx = mesh.shape
    # vectorised table (the literal per-cell loop over 72 000 cells is slow and covered elsewhere)
    x1 = np.array([mesh.bounds[0] + mesh.dims[0] * i for i in range(nx)])
    y1 = np.array([mesh.bounds[2] + mesh.dims[1] * j for j in range(ny)])
    z1 = np.array([mesh.bounds[4] + mesh.dims[2] * k for k in range(nz)])
    z2 = np.array([z1[k] + mesh.dims[2] if k < nz - 1 else mesh.bounds[5] for k in range(nz)])
    tab = np.empty((nz, ny, nx, 6))
    tab[..., 0], tab[..., 1] = x1[None, None, :], (x1 + mesh.dims[0])[None, None, :]
    tab[..., 2], tab[..., 3] = y1[None, :, None], (y1 + mesh.dims[1])[None, :, None]
    tab[..., 4], tab[..., 5] = z1[:, None, None], z2[:, None, None]
    tab = tab.reshape(-1, 6)
    for idx in (0, 7199, 35000, 71999):
        assert tuple(tab[idx]) == mesh.cell(idx)
    K, err = onp.tess_gz(o[:, 0], o[:, 1], o[:, 2], tab, threads=4)
    ref = g["c4_kernel_rows"]
    scale = np.max(np.abs(ref))
    bad = np.argwhere(np.abs(K - ref) > 1e-10 * scale)
    assert len(bad) == 0, bad[:5]
