"""Pin the CPU oracle (oracle/oracle_np.py + oracle/csrc/*.c) against golden vectors produced by
the UNMODIFIED reference (oracle/make_golden.py) and, when /root/reference is mounted, against the
reference run live.  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle_np as onp
from oracle import ref_harness

from tests.helpers import small_prism_setup, synthetic_topo, chain_from_golden


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


# ------------------------------------------------------------------ meshes (bit-exact)
MESH_CASES = {
    "prism_uniform": dict(bounds=(0, 400, 0, 600, 0, 500), spacing=(100, 100, 100)),
    "prism_uniform_nondiv": dict(bounds=(0, 410, -35, 600, 10, 505), spacing=(100, 90, 70)),
    "prism_ratio": dict(bounds=(0, 300, 0, 300, 0, 2000), spacing=(100, 100, 100), ratio=1.3),
    "prism_segment": dict(bounds=(0, 400, 0, 300, 0, 2100), spacing=([100, 200, 300], 100, 100),
                          divisionsection=[0, 300, 900, 2100]),
    "tess_uniform": dict(bounds=(-10, 10, -10, 10, 0, -300000), spacing=(-100000, 5, 5), zdown=False),
    "tess_segment": dict(bounds=(106.5, 109.5, 16, 18, 2000, -60000),
                         spacing=([-1000, -2000, -5000], 0.5, 0.5),
                         divisionsection=[2000, -5000, -15000, -60000], zdown=False),
}


@pytest.mark.parametrize("name", list(MESH_CASES))
def test_oracle_mesh_bit_exact(golden, name):
    g = golden["meshes"]
    m = onp.OracleMesh(**MESH_CASES[name])
    assert tuple(g[name + "_shape"]) == m.shape
    assert np.array_equal(np.array(m.bounds, dtype=np.float64), g[name + "_bounds"])
    tab, idx = m.active_bounds()
    assert np.array_equal(tab, g[name + "_table"])
    assert np.array_equal(m.get_xs(), g[name + "_xs"])
    assert np.array_equal(m.get_ys(), g[name + "_ys"])
    assert np.array_equal(m.get_zs(), g[name + "_zs"])


def test_oracle_carve_masks_bit_exact(golden):
    g = golden["meshes"]
    m = onp.OracleMesh(**MESH_CASES["prism_uniform"])
    t = g["carve_prism_topo"]
    assert np.array_equal(np.array(m.carvetopo(t[:, 0], t[:, 1], t[:, 2])), g["carve_prism_mask"])
    m = onp.OracleMesh(**MESH_CASES["prism_segment"])
    t = g["carve_prismseg_topo"]
    assert np.array_equal(np.array(m.carvetopo(t[:, 0], t[:, 1], t[:, 2])), g["carve_prismseg_mask"])
    m = onp.OracleMesh(**MESH_CASES["tess_segment"])
    t = g["carve_tessseg_topo"]
    mask = np.array(m.carvetopo(t[:, 0], t[:, 1], t[:, 2]))
    assert np.array_equal(mask, g["carve_tessseg_mask"])
    assert mask.size > 0
    tab, idx = m.active_bounds()
    gt = g["carve_tessseg_table"]
    assert np.array_equal(tab, gt[~np.isnan(gt[:, 0])])
    assert np.array_equal(idx, np.where(~np.isnan(gt[:, 0]))[0])
    rho = np.arange(m.size, dtype=np.float64) * 0.5
    rc = onp.rho2carve(rho, mask)
    assert np.array_equal(rc, g["rho2carve_out"])
    assert np.array_equal(onp.carve2rho(rc + 1.0, np.full(m.size, -7.0), mask), g["carve2rho_out"])


# ------------------------------------------------------------------ prism gz
def test_oracle_prism_ka1(golden):
    g = golden["prism"]
    o = g["ka1_obs"]
    _, K = onp.prism_gz(o[:, 0], o[:, 1], o[:, 2], np.array([[0, 100, 0, 100, 0, 100.0]]))
    raw = K[:, 0] / (onp.G * onp.SI2MGAL)
    assert np.allclose(raw, g["ka1_kernel1d"], rtol=1e-13, atol=0)
    # SURVEY section 9 KA1 (values produced by the reference in the survey container)
    assert np.allclose(g["ka1_kernel1d"],
                       [0.19597617292492941, 0.009777761917575845, 96.9388052712568], rtol=1e-12)


def test_oracle_prism_small_meshes(golden):
    g = golden["prism"]
    o = g["small_obs"]
    m = onp.OracleMesh(**MESH_CASES["prism_uniform"])
    tab, _ = m.active_bounds()
    res, K = onp.prism_gz(o[:, 0], o[:, 1], o[:, 2], tab, dens=g["small_dens"])
    assert rel(K, g["small_kernel"]) < 1e-13
    assert np.max(np.abs(K - g["small_kernel"]) / np.abs(g["small_kernel"])) < 1e-9
    assert rel(res, g["small_result"]) < 1e-12
    # carved
    t = g["carved_topo"]
    m.carvetopo(t[:, 0], t[:, 1], t[:, 2])
    assert np.array_equal(np.array(m.mask), g["carved_mask"])
    tab, _ = m.active_bounds()
    o = g["carved_obs"]
    _, K = onp.prism_gz(o[:, 0], o[:, 1], o[:, 2], tab)
    assert K.shape == g["carved_kernel"].shape
    assert rel(K, g["carved_kernel"]) < 1e-13
    # segmented
    m = onp.OracleMesh(**MESH_CASES["prism_segment"])
    tab, _ = m.active_bounds()
    o = g["seg_obs"]
    _, K = onp.prism_gz(o[:, 0], o[:, 1], o[:, 2], tab)
    assert rel(K, g["seg_kernel"]) < 1e-13


def test_oracle_prism_config1_rows(golden):
    g = golden["prism"]
    o = g["c1_obs"]
    rows = g["c1_rows"]
    m = onp.OracleMesh((0, 2000, 0, 3000, 0, 1000), (100, 100, 100))
    tab, _ = m.active_bounds()
    _, K = onp.prism_gz(o[rows, 0], o[rows, 1], o[rows, 2], tab)
    assert rel(K, g["c1_kernel_rows"]) < 1e-13
    # KA2
    assert abs(g["c1_kernel_rows"][0, 0] - 0.6468726475750967) < 1e-13
    assert abs(g["c1_kernel_stats"][0] - 13064.130593145906) < 1e-6


# ------------------------------------------------------------------ tesseroid gz
def _tess_tab(name, topo=None):
    m = onp.OracleMesh(**MESH_CASES[name])
    if topo is not None:
        m.carvetopo(topo[:, 0], topo[:, 1], topo[:, 2])
    tab, _ = m.active_bounds()
    return tab[onp.check_tesseroids(tab)]


def test_oracle_tess_golden(golden):
    g = golden["tesseroid"]
    tab = _tess_tab("tess_uniform")
    o = g["ka6_obs"]
    K, err = onp.tess_gz(o[:, 0], o[:, 1], o[:, 2], tab)
    assert rel(K, g["ka6_kernel"]) < 1e-13
    assert abs(g["ka6_kernel"][0, 0] - 3001.833983253058) < 1e-8  # KA6
    o = g["near_obs"]
    stats = []
    K, err = onp.tess_gz(o[:, 0], o[:, 1], o[:, 2], tab, stats=stats)
    assert rel(K, g["near_kernel"]) < 1e-13
    assert stats[0][0] > 2 * K.size  # subdivision did happen
    tab = _tess_tab("tess_segment", g["segcarve_topo"])
    o = g["segcarve_obs"]
    K, err = onp.tess_gz(o[:, 0], o[:, 1], o[:, 2], tab)
    assert K.shape == g["segcarve_kernel"].shape
    assert rel(K, g["segcarve_kernel"]) < 1e-13


# ------------------------------------------------------------------ potential + sampler
def test_oracle_weighting_and_misfit(golden):
    g = golden["potential_hmc"]
    o = g["small_obs"]
    m = onp.OracleMesh(**MESH_CASES["prism_uniform"])
    tab, _ = m.active_bounds()
    _, A = onp.prism_gz(o[:, 0], o[:, 1], o[:, 2], tab)
    Aw, wm, wminv, wmsq = onp.sensitivity_weighting(A)
    assert rel(Aw, g["small_Aw"]) < 1e-13
    assert np.allclose(wm, g["small_wm"], rtol=1e-14)
    assert np.allclose(wminv, g["small_wminv"], rtol=1e-14)
    assert np.allclose(wmsq, g["small_wmsq"], rtol=1e-14)
    model = onp.OracleModel(g["small_Aw"], g["small_wm"], g["small_dobs"], tuple(g["small_mshape"]))
    for reg in ("Damping", "MS", "Smoothness", "TV"):
        U, grad, dpre, Ud, Um = model.misfit_and_grad(g["mg_x"], g["mg_x0"], None, None,
                                                      "mandatory", 1000, 0.7, regulization=reg,
                                                      beta=0.001)
        assert np.allclose([U, Ud, Um], g[f"mg_{reg}_scalars"], rtol=1e-12)
        assert rel(grad, g[f"mg_{reg}_grad"]) < 1e-12
        assert rel(dpre, g[f"mg_{reg}_dpre"]) < 1e-12
    assert np.array_equal(onp.fd3d((2, 3, 4)).toarray(), g["fd3d_dense_2x3x4"])


@pytest.mark.parametrize("name", ["Damping", "MS", "Smoothness", "TV", "reject", "fixed", "log"])
def test_oracle_chain_matches_reference_trace(golden, name):
    g = golden["potential_hmc"]
    out = chain_from_golden(g, name, runner="oracle")
    ref_x, ref_U = g[f"chain_{name}_steps_x"], g[f"chain_{name}_steps_U"]
    assert out["steps_x"].shape == ref_x.shape
    assert np.array_equal(out["prop_log"][:, :2], g[f"chain_{name}_prop_log"][:, :2])  # L, accept
    assert rel(out["steps_x"], ref_x) < 1e-10
    assert np.max(np.abs(out["steps_U"] - ref_U) / np.abs(ref_U)) < 1e-9
    if g[f"chain_{name}_misfit"].shape[0]:
        assert np.allclose(out["misfit"], g[f"chain_{name}_misfit"], rtol=0, atol=2e-8)
        assert np.allclose(out["models"], g[f"chain_{name}_models"], rtol=0, atol=2e-8)


def test_oracle_config1_anchors(golden):
    g, p = golden["config1"], golden["prism"]
    # KA3 / KA4 / KA5 are what the reference produced in the survey container
    assert abs(g["c1_wm"][0] - 1.1168840553654293) < 1e-12
    assert abs(g["c1_wm"][5999] - 0.04869533943913128) < 1e-13
    assert np.allclose(g["c1_mg_Damping_scalars"][1:],
                       [380.5128465021882, 8.245065461138022, -9.290285077128054,
                        -13.755536472498427, 594.0347299093844], rtol=1e-10)
    assert np.allclose(g["c1_chain_misfit"][:3, :3],
                       [[152.09503395, 148.93294381, 3.16209013],
                        [46.72860629, 44.79280194, 1.93580436],
                        [21.88240457, 19.42843983, 2.45396474]], atol=2e-8)


# ------------------------------------------------------------------ live cross-check
@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted")
def test_oracle_vs_live_reference_prism():
    mod = ref_harness.load_prism_ext()
    rng = np.random.RandomState(0)
    xp, yp, zp = rng.uniform(-500, 1500, 50), rng.uniform(-500, 1500, 50), rng.uniform(-300, 0, 50)
    b = np.array([[0, 100, 200, 350, 10, 90.0], [400, 1000, -100, 0, 0, 500.0]])
    _, K = onp.prism_gz(xp, yp, zp, b)
    for c in range(2):
        res, k1 = np.zeros(50), np.zeros(50)
        mod.gz(xp, yp, zp, *b[c], 1.0, res, k1)
        assert rel(K[:, c], k1 * onp.G * onp.SI2MGAL) < 1e-14
