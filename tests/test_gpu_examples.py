"""GPU parity on BASELINE.json configs 2-4: the product, through the reference's Python API, against
what the UNMODIFIED reference produced on the shipped example data (tests/golden/examples.npz).

c2 example/segmentgrid (segmented-z prisms; Smoothness and MS; 2 chains; + wavelet='3D' vs oracle),
c3 example/realdata (spherical, segmented, carved, grav_fix, prior model; Damping and MS; 2 chains),
c4 example/global (10x60x120 tesseroids; kernel rows).
Tolerances: masks bit-exact; kernel 1e-10 normwise; weights 1e-11; chains: identical (L, accept)
logs, misfit.dat / model.dat rows to the files' 1e-8 print precision."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from gravinv3dhmc_b200 import mesher, utils  # noqa: E402
from gravinv3dhmc_b200.gravmag import tesseroid  # noqa: E402
from gravinv3dhmc_b200.inversion import batched, hmc, potential  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402


def normwise(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


def check_chains(g, key, model, dobs, nsamples, delta, init, apr, bounds, alpha, reg, beta, Sigma, tmp_path,
                 rtol=1e-8):
    """both ranks at once through the batched sampler, rank 0 also through the single-chain path"""
    M = model.M
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = bounds
    nprop = max(len(g[f"{key}_{reg}_r{r}_log"]) for r in (0, 1))
    bt = batched.HMCSampleBatch(model, 2, nsamples, 0, delta, [5, 20], init, apr, b, "mandatory", 1000,
                                dobs, "Fixed", 0.8, alpha, reg, beta, 100, Sigma,
                                save_folder=str(tmp_path / f"{key}_{reg}_b"), quiet=True,
                                max_proposals=nprop)
    for rank in (0, 1):
        log = g[f"{key}_{reg}_r{rank}_log"]
        assert [(L, int(a)) for L, a in bt.proposals[rank]] == [(int(L), int(a)) for L, a in log]
        mis = np.loadtxt(tmp_path / f"{key}_{reg}_b{rank}" / "misfit.dat", ndmin=2)
        mod = np.loadtxt(tmp_path / f"{key}_{reg}_b{rank}" / "model.dat", ndmin=2)
        assert np.allclose(mis, g[f"{key}_{reg}_r{rank}_misfit"], rtol=rtol, atol=2e-8)
        assert np.allclose(mod[-1], g[f"{key}_{reg}_r{rank}_last_model"], rtol=0, atol=max(2e-8, rtol))
    bt.close()
    ch = hmc.HMCSample(model, nsamples, 0, delta, [5, 20], init, apr, b, "mandatory", 1000, dobs,
                       "Fixed", 0.8, alpha, reg, beta, 100, Sigma, myrank=1,
                       save_folder=str(tmp_path / f"{key}_{reg}_s"), quiet=True)
    mis = np.loadtxt(tmp_path / f"{key}_{reg}_s1" / "misfit.dat", ndmin=2)
    assert np.allclose(mis, g[f"{key}_{reg}_r1_misfit"], rtol=rtol, atol=2e-8)
    ch.close()


def test_c2_segmentgrid(golden, tmp_path):
    g = golden["examples"]
    o, dobs = g["c2_obs"], g["c2_dobs"]
    args = (dobs, (0, 2000, 0, 3000, 0, 2100), ([100, 200, 300], 100, 100), (o[:, 0], o[:, 1], o[:, 2]))
    kw = dict(mseg=True, mdivisionsection=[0, 300, 900, 2100], coordinate="cartesian", njobs=5,
              field="gravity", verbose=False)
    model = potential.GravMagModule(*args, **kw)
    assert model.mshape == tuple(g["c2_mshape"])
    assert np.allclose(model.Wm.diagonal(), g["c2_wm"], rtol=1e-12)
    assert normwise(model.Aw.cpu().numpy()[g["c2_rows"]], g["c2_Aw_rows"]) < 1e-10
    M = model.M
    for reg in ("Smoothness", "MS"):
        check_chains(g, "c2", model, dobs, 4, 0.01, np.ones(M) * 0.001, np.ones(M) * 0.001, (0.0, 1.0),
                     1.0, reg, 0.001, 0.001, tmp_path)
    # the shipped driver runs this grid with wavelet='3D' (main_seg.py:37); PyWavelets is absent, so
    # that path is checked against the oracle restatement (parity unpinned upstream)
    wmodel = potential.GravMagModule(*args, wavelet="3D", **kw)
    assert wmodel.Awcp.shape == (600, 11 * 31 * 20)
    om = onp.OracleModel(model.Aw.cpu().numpy(), g["c2_wm"], dobs, model.mshape, wavelet="3D")
    b = np.zeros((M, 2))
    b[:, 1] = 1.0
    ch = hmc.HMCSample(wmodel, 3, 0, 0.01, [5, 20], np.ones(M) * 0.001, np.ones(M) * 0.001, b,
                       "mandatory", 1000, dobs, "Fixed", 0.8, 1.0, "MS", 0.001, 100, 0.001,
                       save_folder=str(tmp_path / "c2w"), quiet=True, max_proposals=12)
    ref = onp.hmc_sample(om, 3, 0, 0.01, [5, 20], np.ones(M) * 0.001, np.ones(M) * 0.001, b,
                         "mandatory", 1000, 1.0, "MS", 0.001, 100, 0.001, max_proposals=12)
    assert [(L, bool(a)) for L, a in ch.proposals] == [(L, bool(a)) for L, a in ref["log"]]
    assert np.max(np.abs(ch.x_final - ref["x"])) < 1e-9 * np.max(np.abs(ref["x"]))


def test_c3_realdata(golden, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)  # carvetopo writes carve_topo_interp.txt into the CWD like the reference
    g = golden["examples"]
    o, t, dobs = g["c3_obs"], g["c3_topo"], g["c3_dobs"]
    model = potential.GravMagModule(dobs, (106.5, 118.5, 16, 28, 2000, -60000),
                                    ([-1000, -2000, -5000], 0.5, 0.5), (o[:, 0], o[:, 1], o[:, 2]),
                                    fixed=True, grav_fix=g["c3_grav_sea"], mseg=True,
                                    mdivisionsection=[2000, -5000, -15000, -60000],
                                    coordinate="spherical", njobs=5, field="gravity", wavelet=False,
                                    mtopo=(t[:, 0], t[:, 1], t[:, 2]), verbose=False)
    assert model.mshape == tuple(g["c3_mshape"])
    assert np.array_equal(np.array(model.mask), g["c3_mask"])  # bit-exact bookkeeping
    # The observations of this data set sit ON the 0.5-degree cell corners and inside the top layer,
    # so ~1.4 % of the pairs subdivide and some reach 100+ leaves a few hundred metres from the
    # observation.  There l^2 = r^2 + rc^2 - 2 r rc cos(psi) cancels ~(r/l)^2 ~ 1e9-fold: a 1-ulp
    # difference between CUDA's and glibc's cos() moves those leaf values by up to ~1e-6 relative --
    # in the reference itself as much as here.  Parity is therefore stated as: subdivision DECISIONS
    # bit-exact for every pair, values 1e-10 wherever the pair has < 9 leaves (98.6 % of the entries
    # that subdivide at all), and <= 1e-5 relative on the deeply subdivided near-field pairs -- where
    # tests/test_gpu_nearfield.py shows against a binary128 evaluation that the reference's own FP64
    # values are up to 9e-4 from the exact quadrature and the GPU's are no further.
    tab = model.mesh.bounds_table()
    Kraw, _ = tesseroid.assemble(o[:, 0], o[:, 1], o[:, 2], tab)
    K = Kraw[:, : model.M].cpu().numpy()
    Ko, _ = onp.tess_gz(o[:, 0], o[:, 1], o[:, 2], tab, threads=8)
    leaves = tesseroid.leaf_counts(o[:, 0], o[:, 1], o[:, 2], tab)
    assert np.array_equal(leaves, onp.tess_leaves(o[:, 0], o[:, 1], o[:, 2], tab, threads=8))
    rel = np.abs(K - Ko) / np.abs(Ko)
    assert rel[leaves < 9].max() < 1e-10
    assert rel.max() < 1e-5 and (rel > 1e-10).sum() < 2e-4 * rel.size
    assert np.allclose(model.Wm.diagonal(), g["c3_wm"], rtol=1e-6)
    assert normwise(model.Aw.cpu().numpy()[g["c3_rows"]], g["c3_Aw_rows"]) < 1e-6
    M = model.M
    init = utils.rho2carve(np.ones(int(np.prod(model.mshape))) * 0.01, model.mask)
    apr = utils.rho2carve(g["c3_apr_mesh"], model.mask)
    U, gr, dpre, Ud, Um = model.misfit_and_grad(model.Wm @ init, model.Wm @ apr, None, None, "mandatory",
                                                1000, 1, regulization="MS", beta=0.01)
    assert np.allclose([U, Ud, Um, gr[0], gr[M // 2], np.linalg.norm(gr)], g["c3_mg_MS"], rtol=1e-6)
    assert normwise(dpre, g["c3_mg_dpre"]) < 1e-6
    for reg in ("Damping", "MS"):
        check_chains(g, "c3", model, dobs, 3, float(g[f"c3_{reg}_delta"]), init, apr, (-0.5, 0.5), 1,
                     reg, 0.01, 0.01, tmp_path, rtol=1e-6)
    with pytest.raises(ValueError, match="Smoothness/TV"):  # carved model: no full grid
        b = np.ones((M, 2))
        hmc.HMCSample(model, 1, 0, 0.01, [5, 20], init, apr, b, "mandatory", 1000, dobs, "Fixed", 0.8,
                      1, "TV", 0.01, 100, 0.01, save_folder=str(tmp_path / "x"), quiet=True)


def test_c4_global_kernel_rows(golden):
    g = golden["examples"]
    o = g["c4_obs_rows"]
    mesh = mesher.TesseroidMesh((-180, 180, -90, 90, 0, -3000000), (-300000, 3, 3))
    assert mesh.shape == tuple(g["c4_shape"])
    tab, _ = tesseroid._check_table(mesh.bounds_table())
    Gd, M = tesseroid.assemble(o[:, 0], o[:, 1], o[:, 2], tab)
    K, ref = Gd[:, :M].cpu().numpy(), g["c4_kernel_rows"]
    bad = np.argwhere(np.abs(K - ref) > 1e-10 * np.max(np.abs(ref)))
    assert len(bad) <= 1, bad[:5]  # a split-threshold tie would show up here
