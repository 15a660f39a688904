"""GPU parity of the batched-chain path (C chains as columns, DMMA contractions): the GEMM kernels
against numpy on ragged shapes, and every chain of a batch against the CPU oracle run as the
reference would run it (`mpiexec -n C`: chain c = process rank c, seed + c).

Tolerances (BASELINE.json north_star): per-leapfrog positions and potentials 1e-9 relative, accept
decisions identical; contractions 1e-13 normwise."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from gravinv3dhmc_b200 import _lib  # noqa: E402
from gravinv3dhmc_b200.inversion import batched, hmc, potential  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402


def normwise(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("n,m,c", [(1, 1, 2), (5, 40, 3), (17, 300, 8), (130, 1000, 9), (257, 2100, 16),
                                   (600, 6000, 33), (1000, 5003, 64), (2049, 4100, 64), (300, 2000, 20), (300, 2000, 48),
                                   (300, 2000, 51)])
def test_gemm_fwd_adj_vs_numpy(n, m, c):
    L = _lib.lib()
    rng = np.random.RandomState(n + 3 * m + c)
    ld = _lib.padded_ld(m)
    A = rng.standard_normal((n, m))
    X = rng.standard_normal((c, m))
    R = rng.standard_normal((c, n))
    plan = C.c_void_p()
    _lib.check(L.gi_plan_create(n, m, ld, c, C.byref(plan)))
    cp, npad = C.c_int32(), C.c_int64()
    _lib.check(L.gi_plan_batch_info(plan, C.byref(cp), C.byref(npad)))
    cp, npad = cp.value, npad.value
    assert cp >= c and cp % 8 == 0 and npad >= n and npad % 16 == 0
    f64 = dict(dtype=torch.float64, device="cuda")
    Ad = torch.zeros((n, ld), **f64)
    Ad[:, :m] = torch.as_tensor(A)
    Xd = torch.zeros((cp, ld), **f64)
    Xd[:c, :m] = torch.as_tensor(X)
    Rd = torch.zeros((cp, npad), **f64)
    Rd[:c, :n] = torch.as_tensor(R)
    D = torch.full((cp, n), np.nan, **f64)
    Gt = torch.full((cp, ld), np.nan, **f64)
    s = _lib.stream_ptr()
    _lib.check(L.gi_gemm_fwd(plan, _lib.ptr(Ad), _lib.ptr(Xd), _lib.ptr(D), s))
    _lib.check(L.gi_gemm_adj(plan, _lib.ptr(Ad), _lib.ptr(Rd), _lib.ptr(Gt), s))
    D1, G1 = D.cpu().numpy(), Gt.cpu().numpy()
    assert normwise(D1[:c], X @ A.T) < 1e-13
    assert normwise(G1[:c, :m], R @ A) < 1e-13
    assert np.all(D1[c:] == 0) and np.all(G1[c:] == 0) and np.all(G1[:, m:] == 0)
    # deterministic: bitwise identical on a second run
    _lib.check(L.gi_gemm_fwd(plan, _lib.ptr(Ad), _lib.ptr(Xd), _lib.ptr(D), s))
    _lib.check(L.gi_gemm_adj(plan, _lib.ptr(Ad), _lib.ptr(Rd), _lib.ptr(Gt), s))
    assert np.array_equal(D.cpu().numpy(), D1) and np.array_equal(Gt.cpu().numpy(), G1)
    # data sums / residual per chain
    dobs_c = rng.standard_normal(n)
    fix = rng.standard_normal(n)
    sums = torch.zeros((cp, 8), **f64)
    Rout = torch.full((cp, npad), np.nan, **f64)
    _lib.check(L.gi_data_sum_batched(plan, _lib.ptr(D), _lib.ptr(torch.as_tensor(fix).cuda()),
                                     _lib.ptr(sums), s))
    fixd, dcd = torch.as_tensor(fix).cuda(), torch.as_tensor(dobs_c).cuda()
    _lib.check(L.gi_residual_batched(plan, _lib.ptr(D), _lib.ptr(fixd), _lib.ptr(dcd), n,
                                     _lib.ptr(Rout), _lib.ptr(sums), s))
    sm, Ro = sums.cpu().numpy(), Rout.cpu().numpy()
    dinv = D1[:c] + fix[None, :]
    rref = (dinv - dinv.mean(axis=1, keepdims=True)) - dobs_c[None, :]
    assert np.allclose(sm[:c, 0], dinv.sum(axis=1), rtol=1e-12, atol=1e-12)
    assert normwise(Ro[:c, :n], rref) < 1e-12 and np.all(Ro[:, n:] == 0)
    assert np.allclose(sm[:c, 1], (rref ** 2).sum(axis=1), rtol=1e-12)
    L.gi_plan_destroy(plan)


def small_model(g, fixed=False):
    o = g["small_obs"]
    kw = dict(fixed=True, grav_fix=g["fixed_gravfix"]) if fixed else {}
    dobs = g["fixed_dobs"] if fixed else g["small_dobs"]
    return potential.GravMagModule(dobs, (0, 400, 0, 600, 0, 500), (100, 100, 100),
                                   (o[:, 0], o[:, 1], o[:, 2]), verbose=False, **kw), dobs


CASES = {
    # name: (regularization, constraint, alpha, beta, delta, Sigma, Lrange, lo, hi, init, fixed)
    # -- the parameter sets of the golden single-chain traces (oracle/make_golden.py)
    "Damping": ("Damping", "mandatory", 1.0, 0.001, 0.02, 0.05, [3, 8], 0.0, 0.3, 0.001, False),
    "MS": ("MS", "mandatory", 0.5, 0.001, 0.02, 0.05, [3, 8], 0.0, 0.3, 0.001, False),
    "Smoothness": ("Smoothness", "mandatory", 2.0, 0.001, 0.02, 0.05, [3, 8], 0.0, 0.3, 0.001, False),
    "TV": ("TV", "mandatory", 0.05, 0.001, 0.02, 0.05, [3, 8], 0.0, 0.3, 0.001, False),
    "reject": ("Damping", "mandatory", 1.0, 0.001, 0.1, 1.0, [4, 9], -5.0, 5.0, 0.001, False),
    "fixed": ("Damping", "mandatory", 1.0, 0.001, 0.02, 0.05, [3, 8], 0.0, 1.0, 0.001, True),
    "log": ("Damping", "logarithmic", 1.0, 0.001, 1e-5, 1e-4, [3, 6], -0.5, 1.5, 0.3, False),
}


@pytest.mark.parametrize("name,nchains", [("Damping", 5), ("MS", 3), ("Smoothness", 2), ("TV", 9),
                                          ("reject", 6), ("fixed", 2), ("log", 3)])
def test_batched_chains_match_oracle_ranks(golden, name, nchains, tmp_path):
    g = golden["potential_hmc"]
    reg, constraint, alpha, beta, delta, Sigma, Lrange, lo, hi, init, fixed = CASES[name]
    model, dobs = small_model(g, fixed)
    M = model.M
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = lo, hi
    nprops, seed = 6, 41
    bt = batched.HMCBatch(model, nchains, delta, Lrange, np.ones(M) * init, np.ones(M) * init, b,
                          constraint, 1000, dobs, alpha, reg, beta, seed, Sigma,
                          save_folder=str(tmp_path / "chain"), quiet=True)
    traces = []
    for _ in range(nprops):
        tr = {}
        bt.propose(trace=tr)
        traces.append(tr)
    om = onp.OracleModel(g["small_Aw"], g["small_wm"], dobs, tuple(g["small_mshape"]), fixed=fixed,
                         grav_fix=g["fixed_gravfix"] if fixed else None)
    n_acc = n_rej = 0
    for c in range(nchains):
        otr = []
        onp.hmc_sample(om, 10 ** 6, 0, delta, Lrange, np.ones(M) * init, np.ones(M) * init, b,
                       constraint, 1000, alpha, reg, beta, seed, Sigma, myrank=c,
                       max_proposals=nprops, trace=otr)
        assert [(L, bool(a)) for L, a in bt.proposals[c]] == [(t["L"], bool(t["accept"])) for t in otr]
        for k, t in enumerate(otr):
            L = t["L"]
            ref_x = np.array([x for x, _ in t["steps"]])
            ref_U = np.array([U for _, U in t["steps"]])
            got_x, got_U = traces[k]["x"][: L + 1, c], traces[k]["U"][: L + 1, c]
            scale = np.max(np.abs(ref_x), axis=1, keepdims=True)
            assert np.max(np.abs(got_x - ref_x) / scale) < 1e-9
            assert np.max(np.abs(got_U - ref_U) / np.abs(ref_U)) < 1e-9
            assert abs(traces[k]["Hnew"][c] - t["Hnew"]) < 1e-9 * abs(t["Hnew"])
            n_acc += bool(t["accept"])
            n_rej += not t["accept"]
    if name == "reject":
        assert n_acc > 0 and n_rej > 0  # both Metropolis branches, per chain independent
    bt.close()


@pytest.mark.parametrize("mode", ["stream", "lockstep"])
def test_batch_sample_files_match_oracle(golden, tmp_path, mode):
    """HMCSampleBatch writes, for every chain c, the files the reference process `myrank=c` writes;
    chains that reach nsamples stop being recorded while the others continue.  Both schedulers:
    lockstep rounds and the streaming sampler (chains restart inside the step they finish in)."""
    g = golden["potential_hmc"]
    model, dobs = small_model(g)
    M = model.M
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = -5.0, 5.0
    args = dict(delta=0.1, Lrange=[4, 9], Sigma=1.0, alpha=1.0, beta=0.001, seed=3)
    nchains, nsamples = 4, 5
    bt = batched.HMCSampleBatch(model, nchains, nsamples, 0, args["delta"], args["Lrange"],
                                np.ones(M) * 0.001, np.ones(M) * 0.001, b, "mandatory", 1000, dobs,
                                "Fixed", 0.8, args["alpha"], "Damping", args["beta"], args["seed"],
                                args["Sigma"], save_folder=str(tmp_path / "run"), quiet=True, mode=mode)
    om = onp.OracleModel(g["small_Aw"], g["small_wm"], dobs, tuple(g["small_mshape"]))
    lens = set()
    for c in range(nchains):
        ref = onp.hmc_sample(om, nsamples, 0, args["delta"], args["Lrange"], np.ones(M) * 0.001,
                             np.ones(M) * 0.001, b, "mandatory", 1000, args["alpha"], "Damping",
                             args["beta"], args["seed"], args["Sigma"], myrank=c)
        mis = np.loadtxt(tmp_path / f"run{c}" / "misfit.dat", ndmin=2)
        mod = np.loadtxt(tmp_path / f"run{c}" / "model.dat", ndmin=2)
        assert mis.shape == (nsamples, 7) and mod.shape == (nsamples, M)
        assert np.allclose(mis, ref["misfit"], rtol=0, atol=2e-8)
        assert np.allclose(mod, ref["models"], rtol=0, atol=2e-8)
        assert [(L, bool(a)) for L, a in bt.proposals[c]] == [(L, bool(a)) for L, a in ref["log"]]
        lens.add(len(ref["log"]))
    assert len(lens) > 1  # the chains needed different numbers of proposals
    if mode == "stream":
        # no idling: the batch ran about as many gradient evaluations as the busiest chain needed
        busiest = max(sum(L for L, _ in bt.proposals[c]) for c in range(nchains))
        assert busiest <= bt.stream_steps <= busiest + 2 * args["Lrange"][1]
    bt.close()


@pytest.mark.parametrize("name", ["MS", "TV", "fixed", "log"])
def test_streaming_chains_match_oracle(golden, name, tmp_path):
    """the streaming scheduler reproduces every chain's proposal log and final state"""
    g = golden["potential_hmc"]
    reg, constraint, alpha, beta, delta, Sigma, Lrange, lo, hi, init, fixed = CASES[name]
    model, dobs = small_model(g, fixed)
    M = model.M
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = lo, hi
    nchains, nprops, seed = 7, 5, 11
    bt = batched.HMCBatch(model, nchains, delta, Lrange, np.ones(M) * init, np.ones(M) * init, b,
                          constraint, 1000, dobs, alpha, reg, beta, seed, Sigma,
                          save_folder=str(tmp_path / "st"), quiet=True)
    recs = []
    bt.stream(10 ** 6, 0, max_proposals=nprops, write=False,
              on_record=lambda c, r, acc: recs.append((c, int(r.seq), r.U, r.Hnew)))
    om = onp.OracleModel(g["small_Aw"], g["small_wm"], dobs, tuple(g["small_mshape"]), fixed=fixed,
                         grav_fix=g["fixed_gravfix"] if fixed else None)
    for c in range(nchains):
        otr = []
        ref = onp.hmc_sample(om, 10 ** 6, 0, delta, Lrange, np.ones(M) * init, np.ones(M) * init, b,
                             constraint, 1000, alpha, reg, beta, seed, Sigma, myrank=c,
                             max_proposals=nprops, trace=otr)
        assert [(L, bool(a)) for L, a in bt.proposals[c]] == [(L, bool(a)) for L, a in ref["log"]]
        assert np.max(np.abs(bt.x[c] - ref["x"])) < 1e-9 * np.max(np.abs(ref["x"]))
        mine = sorted((seq, Hn) for cc, seq, U, Hn in recs if cc == c)
        assert [s for s, _ in mine] == list(range(nprops))
        for (seq, Hn), t in zip(mine, otr):
            assert abs(Hn - t["Hnew"]) < 1e-9 * abs(t["Hnew"])
    bt.close()


def test_batch_matches_single_chain_product_path(golden, tmp_path):
    """chain 0 of a batch == hmc.HMCSample with myrank=0 (GEMV path), same draws."""
    g = golden["potential_hmc"]
    model, dobs = small_model(g)
    M = model.M
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = 0.0, 0.3
    common = (0.01, [3, 9], np.ones(M) * 0.001, np.ones(M) * 0.001, b, "mandatory", 1000, dobs)
    ch = hmc.HMCSample(model, 4, 0, *common, "Fixed", 0.8, 1.0, "MS", 0.001, 5, 0.01, myrank=0,
                       save_folder=str(tmp_path / "s"), quiet=True)
    bt = batched.HMCSampleBatch(model, 2, 4, 0, *common, "Fixed", 0.8, 1.0, "MS", 0.001, 5, 0.01,
                                save_folder=str(tmp_path / "b"), quiet=True)
    a = np.loadtxt(tmp_path / "s0" / "model.dat", ndmin=2)
    bb = np.loadtxt(tmp_path / "b0" / "model.dat", ndmin=2)
    assert ch.proposals == bt.proposals[0]
    assert np.allclose(a, bb, rtol=0, atol=2e-8)
    bt.close()


def test_batch_philox_reproducible_and_errors(golden, tmp_path):
    g = golden["potential_hmc"]
    model, dobs = small_model(g)
    M = model.M
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = 0.0, 1.0
    out = []
    for k in range(2):
        bt = batched.HMCSampleBatch(model, 8, 6, 0, 0.02, [3, 8], np.ones(M) * 0.001,
                                    np.ones(M) * 0.001, b, "mandatory", 1000, dobs, "Fixed", 0.8, 1.0,
                                    "Damping", 0.001, 7, 0.05, save_folder=str(tmp_path / f"p{k}"),
                                    quiet=True, rng="philox", max_proposals=40)
        out.append((bt.proposals, bt.x.copy()))
        bt.close()
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1])
    assert not np.array_equal(out[0][1][0], out[0][1][1])  # chains differ (key = seed + c)
    with pytest.raises(ValueError, match="2..64"):
        batched.HMCBatch(model, 65, 0.01, [1, 2], np.ones(M), np.ones(M), b, "mandatory", 1000, dobs,
                         1.0, "Damping", 0.001, 1, 0.01)
    with pytest.raises(ValueError, match="regularization"):
        batched.HMCBatch(model, 2, 0.01, [1, 2], np.ones(M), np.ones(M), b, "mandatory", 1000, dobs,
                         1.0, "L1", 0.001, 1, 0.01)


@pytest.mark.parametrize("npieces", [1, 2])
def test_shard_hook_machinery_single_rank(tmp_path, npieces):
    """the row-sharded device loop (exchange hooks, piece-major adjoint output reduced piece by piece)
    with ONE rank, where every reduction is the identity: it must reproduce the plain batched chain
    (to rounding: the centred data come from numpy's mean here and from a long-double mean there;
    the multi-rank runs are in tests/multi_gpu_check.py)"""
    rng = np.random.RandomState(4)
    mesh_range, spacing = (0, 1600, 0, 1600, 0, 200), (100, 100, 100)   # 2 x 16 x 16 = 512 cells
    xs, ys = np.meshgrid(np.linspace(50, 1550, 7), np.linspace(50, 1550, 6))
    obs = (xs.ravel(), ys.ravel(), np.full(xs.size, -20.0))
    model = potential.GravMagModule(rng.standard_normal(xs.size), mesh_range, spacing, obs, verbose=False)
    M = model.M
    assert model.ld == 512
    b = np.zeros((M, 2))
    b[:, 1] = 1.0
    out = {}
    for driver in ("device", ("device-hooks", npieces)):
        bt = batched.HMCBatch(model, 5, 0.02, [3, 8], np.ones(M) * 0.001, np.ones(M) * 0.001, b,
                              "mandatory", 1000, model.dobs, 0.5, "TV", 0.001, 9, 0.05,
                              save_folder=str(tmp_path / "h"), quiet=True, driver=driver)
        bt.stream(10 ** 6, 0, max_proposals=4, write=False)
        out[driver if isinstance(driver, str) else driver[0]] = (bt.proposals, bt.x.copy())
        bt.close()
    assert out["device"][0] == out["device-hooks"][0]
    assert np.allclose(out["device"][1], out["device-hooks"][1], rtol=1e-10, atol=1e-14)


def test_mid_size_contractions_use_the_production_tiling():
    """bench.py's `mid` shape (4096 observations x 131 072 voxels, 64 chains): large enough for the
    production tile chooser -- whole waves of 148 CTAs, several k-chunks, 512 adjoint strips -- so the
    tiling the c5 runs use is exercised under pytest, against a torch FP64 product
    on sampled rows / columns and the linearity of both passes."""
    L = _lib.lib()
    n, m, nch = 4096, 131072, 64
    ld = _lib.padded_ld(m)
    f64 = dict(dtype=torch.float64, device="cuda")
    gen = torch.Generator(device="cuda")
    gen.manual_seed(42)
    A = torch.randn((n, ld), generator=gen, **f64)
    plan = C.c_void_p()
    _lib.check(L.gi_plan_create(n, m, ld, nch, C.byref(plan)))
    cp, npad = C.c_int32(), C.c_int64()
    _lib.check(L.gi_plan_batch_info(plan, C.byref(cp), C.byref(npad)))
    assert cp.value == 64 and npad.value == n
    X = torch.randn((cp.value, ld), generator=gen, **f64)
    R = torch.randn((cp.value, npad.value), generator=gen, **f64)
    D = torch.zeros((cp.value, n), **f64)
    G = torch.zeros((cp.value, ld), **f64)
    s = _lib.stream_ptr()
    _lib.check(L.gi_gemm_fwd(plan, _lib.ptr(A), _lib.ptr(X), _lib.ptr(D), s))
    _lib.check(L.gi_gemm_adj(plan, _lib.ptr(A), _lib.ptr(R), _lib.ptr(G), s))
    rows = torch.tensor([0, 1, 127, 128, 2047, 2048, 4094, 4095], device="cuda")
    ref = X @ A[rows].T
    assert float((D[:, rows] - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    cols = torch.unique(torch.cat([torch.linspace(0, m - 1, 48, device="cuda").long(),
                                   torch.tensor([0, 255, 256, 65535, 65536, m - 257, m - 256, m - 1], device="cuda")]))
    ref = R @ A[:, cols]
    assert float((G[:, cols] - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    # deterministic and linear: a second run gives the same bits, doubling the input doubles the output
    D2, G2 = torch.zeros_like(D), torch.zeros_like(G)
    _lib.check(L.gi_gemm_fwd(plan, _lib.ptr(A), _lib.ptr(X), _lib.ptr(D2), s))
    _lib.check(L.gi_gemm_adj(plan, _lib.ptr(A), _lib.ptr(R), _lib.ptr(G2), s))
    assert torch.equal(D, D2) and torch.equal(G, G2)
    X.mul_(2.0)
    _lib.check(L.gi_gemm_fwd(plan, _lib.ptr(A), _lib.ptr(X), _lib.ptr(D2), s))
    assert torch.equal(D2, 2.0 * D)
    L.gi_plan_destroy(plan)
