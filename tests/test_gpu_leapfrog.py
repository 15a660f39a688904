"""GPU parity of the leapfrog hot path: G.m, G^T r, regulariser gradients, the fused update with
clamp-and-flip, and the Metropolis test -- CUDA through the C ABI vs the golden traces recorded
from the unmodified reference (and the CPU oracle on seeded inputs).

Tolerances (BASELINE.json north_star): per-leapfrog positions and potentials 1e-9 relative, accept
decisions identical, over the recorded chains; forward data 1e-10."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from gravinv3dhmc_b200 import _lib  # noqa: E402
from gravinv3dhmc_b200.inversion import hmc, potential  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402
from tests.helpers import chain_params  # noqa: E402


def normwise(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def small_model(g, fixed=False):
    o = g["small_obs"]
    kw = dict(fixed=True, grav_fix=g["fixed_gravfix"]) if fixed else {}
    dobs = g["fixed_dobs"] if fixed else g["small_dobs"]
    return potential.GravMagModule(dobs, (0, 400, 0, 600, 0, 500), (100, 100, 100),
                                   (o[:, 0], o[:, 1], o[:, 2]), verbose=False, **kw), dobs


# ------------------------------------------------------------------ raw GEMV kernels
@pytest.mark.parametrize("n,m", [(1, 1), (3, 33), (7, 1024), (600, 6000), (1000, 5003), (2049, 4100)])
def test_gemv_fwd_adj_vs_numpy(n, m):
    """ragged shapes: every tile/remainder path of the two streaming kernels"""
    L = _lib.lib()
    rng = np.random.RandomState(n * 7 + m)
    ld = _lib.padded_ld(m)
    A = rng.standard_normal((n, m))
    x, r = rng.standard_normal(m), rng.standard_normal(n)
    Ad = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
    Ad[:, :m] = torch.as_tensor(A)
    xd = torch.zeros(ld, dtype=torch.float64, device="cuda")
    xd[:m] = torch.as_tensor(x)
    rd = torch.as_tensor(r).cuda()
    d = torch.empty(n, dtype=torch.float64, device="cuda")
    g = torch.empty(ld, dtype=torch.float64, device="cuda")
    plan = C.c_void_p()
    _lib.check(L.gi_plan_create(n, m, ld, 1, C.byref(plan)))
    s = _lib.stream_ptr()
    _lib.check(L.gi_gemv_fwd(plan, _lib.ptr(Ad), _lib.ptr(xd), _lib.ptr(d), s))
    _lib.check(L.gi_gemv_adj(plan, _lib.ptr(Ad), _lib.ptr(rd), _lib.ptr(g), s))
    d1, g1 = d.cpu().numpy(), g.cpu().numpy()
    assert normwise(d1, A @ x) < 1e-13
    assert normwise(g1[:m], A.T @ r) < 1e-13
    assert np.all(g1[m:] == 0)
    # deterministic: bitwise identical on a second run
    _lib.check(L.gi_gemv_fwd(plan, _lib.ptr(Ad), _lib.ptr(xd), _lib.ptr(d), s))
    _lib.check(L.gi_gemv_adj(plan, _lib.ptr(Ad), _lib.ptr(rd), _lib.ptr(g), s))
    assert np.array_equal(d.cpu().numpy(), d1) and np.array_equal(g.cpu().numpy(), g1)
    # linearity (size independent property): A(2x) == 2 A x bitwise, A^T(r+r) == 2 A^T r
    xd.mul_(2.0)
    _lib.check(L.gi_gemv_fwd(plan, _lib.ptr(Ad), _lib.ptr(xd), _lib.ptr(d), s))
    assert np.array_equal(d.cpu().numpy(), 2.0 * d1)
    L.gi_plan_destroy(plan)


# ------------------------------------------------------------------ misfit_and_grad
@pytest.mark.parametrize("reg", ["Damping", "MS", "Smoothness", "TV"])
def test_misfit_and_grad_matches_reference(golden, reg):
    g = golden["potential_hmc"]
    model, _ = small_model(g)
    U, grad, dpre, Ud, Um = model.misfit_and_grad(g["mg_x"], g["mg_x0"], None, None, "mandatory",
                                                  1000, 0.7, regulization=reg, beta=0.001)
    assert np.allclose([U, Ud, Um], g[f"mg_{reg}_scalars"], rtol=1e-10)
    assert normwise(grad, g[f"mg_{reg}_grad"]) < 1e-10
    assert normwise(dpre, g[f"mg_{reg}_dpre"]) < 1e-10


def test_model_terms_and_errors(golden):
    g = golden["potential_hmc"]
    model, dobs = small_model(g)
    om = onp.OracleModel(g["small_Aw"], g["small_wm"], dobs, tuple(g["small_mshape"]))
    x, x0 = g["mg_x"], g["mg_x0"]
    for got, ref in ((model.model_Damping_all(x, x0), om.model_Damping_all(x, x0)),
                     (model.model_MS_all(x, x0, 0.01), om.model_MS_all(x, x0, 0.01)),
                     (model.model_Smoothness_all(x, x0), om.model_Smoothness_all(x, x0)),
                     (model.model_TV_all(x, x0, 0.01), om.model_TV_all(x, x0, 0.01))):
        assert abs(got[0] - ref[0]) <= 1e-11 * abs(ref[0])
        assert normwise(got[1], ref[1]) < 1e-11
    dpre, dv, dg = model.data_all(x)
    rd, rv, rg = om.data_all(x)
    assert normwise(dpre, rd) < 1e-10 and abs(dv - rv) < 1e-10 * rv and normwise(dg, rg) < 1e-10
    with pytest.raises(ValueError, match="regularization"):
        model.misfit_and_grad(x, x0, None, None, "mandatory", 1000, 1.0, regulization="L1")
    with pytest.raises(ValueError, match="boundary constraint"):
        model.misfit_and_grad(x, x0, None, None, "periodic", 1000, 1.0)


# ------------------------------------------------------------------ chains vs reference traces
def run_product_chain(g, name, tmp_path, rng="numpy"):
    p = chain_params(g, name)
    model, dobs = small_model(g, fixed=p["fixed"])
    M = model.M
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = p["lo"], p["hi"]
    nprops = g[f"chain_{name}_prop_log"].shape[0]
    chain = hmc.HamitonianMC(model)
    traces = []
    orig = chain._leapfrog

    def lf(xcur, dt, L, alpha, fignum=0):
        tr = {}
        out = orig(xcur, dt, L, alpha, fignum, trace=tr)
        traces.append(tr)
        return out

    # HMCSample builds its own chain object; re-use its setup by patching the class method
    hmc.HamitonianMC._leapfrog_traced = None
    real = hmc.HamitonianMC._leapfrog

    def patched(self, xcur, dt, L, alpha, fignum=0, trace=None):
        tr = {}
        out = real(self, xcur, dt, L, alpha, fignum, trace=tr)
        traces.append(tr)
        return out

    hmc.HamitonianMC._leapfrog = patched
    try:
        ch = hmc.HMCSample(model, p["nsamples"], 0, p["delta"], p["Lrange"], np.ones(M) * p["init"],
                           np.ones(M) * p["init"], b, p["constraint"], 1000, dobs, "Fixed", 0.8,
                           p["alpha"], p["reg"], p["beta"], p["seed"], p["Sigma"], myrank=0,
                           save_folder=str(tmp_path / "chain"), quiet=True, max_proposals=nprops)
    finally:
        hmc.HamitonianMC._leapfrog = real
    xs = np.concatenate([t["x"] for t in traces])
    Us = np.concatenate([t["U"] for t in traces])
    mis = np.loadtxt(tmp_path / "chain0" / "misfit.dat", ndmin=2) \
        if (tmp_path / "chain0" / "misfit.dat").exists() else np.zeros((0, 7))
    mod = np.loadtxt(tmp_path / "chain0" / "model.dat", ndmin=2) \
        if (tmp_path / "chain0" / "model.dat").exists() else np.zeros((0, M))
    return ch, xs, Us, mis, mod


@pytest.mark.parametrize("name", ["Damping", "MS", "Smoothness", "TV", "reject", "fixed", "log"])
def test_chain_matches_reference_trace(golden, name, tmp_path):
    g = golden["potential_hmc"]
    ch, xs, Us, mis, mod = run_product_chain(g, name, tmp_path)
    ref_x, ref_U = g[f"chain_{name}_steps_x"], g[f"chain_{name}_steps_U"]
    log = g[f"chain_{name}_prop_log"]
    assert [(L, int(a)) for L, a in ch.proposals] == [(int(L), int(a)) for L, a in log[:, :2]]
    assert xs.shape == ref_x.shape
    # per-leapfrog positions (normwise per step) and potentials: 1e-9 relative
    scale = np.max(np.abs(ref_x), axis=1, keepdims=True)
    assert np.max(np.abs(xs - ref_x) / scale) < 1e-9
    assert np.max(np.abs(Us - ref_U) / np.abs(ref_U)) < 1e-9
    if g[f"chain_{name}_misfit"].shape[0]:
        assert np.allclose(mis, g[f"chain_{name}_misfit"], rtol=0, atol=2e-8)
        assert np.allclose(mod, g[f"chain_{name}_models"], rtol=0, atol=2e-8)
    if name == "reject":
        assert 0 < sum(a for _, a in ch.proposals) < len(ch.proposals)  # both Metropolis branches
    if name in ("Damping", "MS"):
        hit = (np.isclose(xs, 0.0) | np.isclose(xs, 0.3 * g["small_wm"][None, :])).any()
        assert hit  # the clamp-and-flip branch fired


def test_config1_chain_first_samples(golden, tmp_path):
    """KA5: the reference's example/uniformgrid chain (Damping, seed 100), first 8 samples."""
    p, c = golden["prism"], golden["config1"]
    o = p["c1_obs"]
    model = potential.GravMagModule(p["c1_dobs"], (0, 2000, 0, 3000, 0, 1000), (100, 100, 100),
                                    (o[:, 0], o[:, 1], o[:, 2]), verbose=False)
    M = model.M
    b = np.ones((M, 2))
    b[:, 0] = 0
    x = model.Wm @ (0.5 * np.linspace(0, 1, M))
    x0 = model.Wm @ (0.001 * np.ones(M))
    for reg in ("Damping", "MS", "Smoothness", "TV"):
        U, g_, dpre, Ud, Um = model.misfit_and_grad(x, x0, None, None, "mandatory", 1000, 1,
                                                    regulization=reg, beta=0.001)
        ref = c[f"c1_mg_{reg}_scalars"]
        assert np.allclose([U, Ud, Um, g_[0], g_[3000], np.linalg.norm(g_)], ref, rtol=1e-9)
    ch = hmc.HMCSample(model, 8, 0, 0.01, [5, 20], np.ones(M) * 0.001, np.ones(M) * 0.001, b,
                       "mandatory", 1000, p["c1_dobs"], "Fixed", 0.8, 1, "Damping", 0.001, 100, 0.001,
                       save_folder=str(tmp_path / "c"), quiet=True)
    mis = np.loadtxt(tmp_path / "c0" / "misfit.dat", ndmin=2)
    assert [(L, int(a)) for L, a in ch.proposals] == [(int(L), int(a)) for L, a in c["c1_chain_prop_log"][:, :2]]
    assert np.allclose(mis, c["c1_chain_misfit"], rtol=1e-8, atol=2e-8)
    mod = np.loadtxt(tmp_path / "c0" / "model.dat", ndmin=2)
    assert np.allclose(mod[-1], c["c1_chain_last_model"], rtol=0, atol=2e-8)


def test_philox_chain_runs_and_is_reproducible(golden, tmp_path):
    g = golden["potential_hmc"]
    model, dobs = small_model(g)
    M = model.M
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = 0.0, 1.0
    out = []
    for k in range(2):
        ch = hmc.HMCSample(model, 10, 0, 0.02, [3, 8], np.ones(M) * 0.001, np.ones(M) * 0.001, b,
                           "mandatory", 1000, dobs, "Fixed", 0.8, 1.0, "Damping", 0.001, 7, 0.05,
                           save_folder=str(tmp_path / f"p{k}"), quiet=True, rng="philox",
                           max_proposals=60)
        out.append((ch.proposals, ch.x_final.copy()))
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1])
    assert sum(a for _, a in out[0][0]) >= 5
    # the device normal generator has the right first two moments
    L = _lib.lib()
    n = 1 << 20
    A = torch.zeros((4, n), dtype=torch.float64, device="cuda")
    A[:, 0] = 1.0
    cfg = _lib.HmcConfig(4, n, n, 0, 0, _lib.RegParams(0, 0, 1, 1, n, 0, 1.0, 0.01, 1.0))
    z = np.zeros(n)
    hi = np.full(n, 1e9)
    h = C.c_void_p()
    _lib.check(L.gi_hmc_create(C.byref(cfg), _lib.ptr(A), _lib.ptr(np.zeros(4)), None, _lib.ptr(-hi),
                               _lib.ptr(hi), _lib.ptr(z), None, None, C.byref(h)))
    _lib.check(L.gi_hmc_set_state(h, _lib.ptr(z)))
    res = _lib.HmcResult()
    _lib.check(L.gi_hmc_propose_philox(h, 1234, 0, 1.0, 1, 1e-9, C.byref(res)))
    # K0 = 0.5 * sum p^2 with p ~ N(0,1): Hcur - U ~ n/2
    assert abs((res.Hcur - 0.0) / (0.5 * n) - 1.0) < 0.01
    L.gi_hmc_destroy(h)
