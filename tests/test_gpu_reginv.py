"""GPU parity of the regularised conjugate gradient and its bootstrap (gi_cg_*)
against (a) golden vectors from the UNMODIFIED reference's inversion/reginv.py and (b) the CPU oracle
on larger seeded problems.

Tolerance: every per-iteration quantity (regularisation factor, normed data / model error) and the
final model / forward data to 1e-9 relative (the north_star's bar for per-step sampler quantities);
iteration counts and early-stop decisions identical."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from gravinv3dhmc_b200.inversion import reginv  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402

MRANGE, MSPACING, MSHAPE = (0, 800, 0, 600, 0, 400), (100, 100, 100), (4, 6, 8)
REGS = ("Damping", "MS", "Smoothness", "TV")
TOL = 1e-9


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def obs_of(g, key="obs"):
    o = g[key]
    return o[:, 0].copy(), o[:, 1].copy(), o[:, 2].copy()


@pytest.fixture(scope="module")
def cg_small(golden):
    g = golden["reginv"]
    return g, reginv.ConjugateGradient(g["dobs"], MRANGE, MSPACING, obs_of(g), verbose=False)


@pytest.mark.parametrize("reg", REGS)
def test_cg_vs_reference_golden(cg_small, reg):
    g, cg = cg_small
    assert cg.mshape == MSHAPE and cg.dsize == 80 and cg.msize == 192
    assert rel(cg.Wm.diagonal(), g["cg_wm"]) < 1e-12
    assert rel(cg.Aw[[0, 17, 79]].cpu().numpy(), g["cg_Aw_rows"]) < 1e-10
    m, d, dm, mm, rf = cg.CG(g["initial"], g["aprior"], g["boundary"], regularization=reg,
                             beta=float(g["cg_%s_beta" % reg]), q=0.9, maxk=14)
    assert len(rf) == len(dm) == len(mm) == 14
    assert rel(rf, g["cg_%s_regul" % reg]) < TOL
    assert rel(dm, g["cg_%s_data_misfit" % reg]) < TOL
    assert rel(mm, g["cg_%s_model_misfit" % reg]) < TOL
    assert rel(m, g["cg_%s_model" % reg]) < TOL
    assert rel(d, g["cg_%s_data" % reg]) < TOL
    assert cg.last_launches > 0
    # bitwise reproducible
    m2 = cg.CG(g["initial"], g["aprior"], g["boundary"], regularization=reg,
               beta=float(g["cg_%s_beta" % reg]), q=0.9, maxk=14)[0]
    assert np.array_equal(m, m2)


def test_cg_early_stop(golden):
    g = golden["reginv"]
    cg = reginv.ConjugateGradient(g["cgstop_dobs"], MRANGE, MSPACING, obs_of(g), verbose=False)
    m, d, dm, mm, rf = cg.CG(g["initial"], np.zeros(192), (-5.0, 5.0), regularization="Damping",
                             beta=0.01, q=0.5, maxk=50)
    assert len(rf) == len(dm) == len(mm) == len(g["cgstop_regul"]) == 2
    assert rel(rf, g["cgstop_regul"]) < TOL and rel(dm, g["cgstop_data_misfit"]) < TOL
    assert rel(mm, g["cgstop_model_misfit"]) < TOL
    assert rel(m, g["cgstop_model"]) < TOL and rel(d, g["cgstop_data"]) < TOL


def test_cg_single_evaluations_and_errors(cg_small):
    g, cg = cg_small
    A = (cg.A).cpu().numpy()
    ocg = onp.OracleCG(A, g["dobs"], MSHAPE)
    mw = cg.Wm @ (0.3 * np.linspace(0, 1, 192))
    apr = cg.Wm @ g["aprior"]
    assert abs(cg.data(mw) - ocg.data(mw)) < 1e-10 * ocg.data(mw)
    assert rel(cg.data_gfun(mw), ocg.data_gfun(mw)) < 1e-10
    for reg, beta in (("MS", 0.003), ("Damping", 0.0), ("Smoothness", 0.0), ("TV", 0.002)):
        args = (mw, apr) + ((beta,) if reg in ("MS", "TV") else ())
        v = getattr(cg, "model_" + reg)(*args)
        gv = getattr(cg, "model_gfun_" + reg)(*args)
        assert abs(v - ocg.model(reg, mw, apr, beta)) < 1e-11 * abs(ocg.model(reg, mw, apr, beta))
        assert rel(gv, ocg.model_gfun(reg, mw, apr, beta)) < 1e-11
    with pytest.raises(ValueError):
        cg.CG(g["initial"], g["aprior"], g["boundary"], regularization="L1")
    with pytest.raises(ValueError):
        reginv.ConjugateGradient(g["dobs"], MRANGE, MSPACING, obs_of(g), field="electric", verbose=False)
    # the Cartesian magnetic branch (reginv.py:75-92): same weighted tf kernel as GravMagModule
    mag = reginv.ConjugateGradient(g["dobs"], MRANGE, MSPACING, obs_of(g), field="magnetic",
                                   mangle=(55.0, -8.0), verbose=False)
    xp, yp, zp = obs_of(g)
    tab, _ = onp.OracleMesh(MRANGE, MSPACING).active_bounds()
    _, K = onp.prism_field("tf", xp, yp, zp, tab, inc=55.0, dec=-8.0)
    assert rel(mag.Aw.cpu().numpy(), onp.sensitivity_weighting(K)[0]) < 1e-10


def test_cg_spherical_vs_reference_golden(golden):
    g = golden["reginv"]
    cg = reginv.ConjugateGradient(g["t_dobs"], (-10, 10, -10, 10, 0, -300000), (-100000, 5, 5),
                                  obs_of(g, "t_obs"), coordinate="spherical", verbose=False)
    m, d, dm, mm, rf = cg.CG(np.full(cg.msize, 0.001), np.zeros(cg.msize), (0.0, 0.4),
                             regularization="Damping", beta=0.01, q=0.9, maxk=10)
    assert rel(rf, g["t_regul"]) < TOL and rel(dm, g["t_data_misfit"]) < TOL
    assert rel(mm, g["t_model_misfit"]) < TOL
    assert rel(m, g["t_model"]) < TOL and rel(d, g["t_data"]) < TOL


@pytest.mark.parametrize("batch", [64, 2, 1])
def test_bootstrap_vs_reference_golden(golden, batch):
    """batch=64: the 5 replicates share one DMMA batch; 2: batches 2+2+1 (DMMA, DMMA, GEMV); 1: GEMV"""
    g = golden["reginv"]
    bs = reginv.BootStrap(MRANGE, MSPACING, obs_of(g), g["dobs"], tuple(g["boundary"]), samples=5,
                          beta=float(g["bs_beta"]), maxk=9, batch=batch, verbose=False)
    mi, dmi, mmi, rfi = bs.BSCG(g["initial"])
    assert mi.shape == (5, 192) and dmi.shape == (5, 8) and rfi.shape == (5, 9)
    assert rel(rfi, g["bs_regul"]) < TOL
    assert rel(dmi, g["bs_data_misfit"]) < TOL
    assert rel(mmi, g["bs_model_misfit"]) < TOL
    assert rel(mi, g["bs_models"]) < TOL
    assert bs.last_launches > 0


def test_bootstrap_cg_on_gathered_rows_and_early_stop(golden):
    g = golden["reginv"]
    bs = reginv.BootStrap(MRANGE, MSPACING, obs_of(g), g["dobs"], tuple(g["boundary"]), samples=5,
                          beta=float(g["bs_beta"]), maxk=9, verbose=False)
    idx = g["bs_index"][3]
    AwS = bs.Aw[torch.as_tensor(idx, device=bs.Aw.device)]
    m, dm, mm, rf = bs.CG(AwS, g["dobs"][idx], g["initial"])
    assert len(dm) == 8 and len(rf) == 9
    assert rel(m, g["bs_models"][3]) < TOL and rel(dm, g["bs_data_misfit"][3]) < TOL
    assert rel(rf, g["bs_regul"][3]) < TOL
    # reginv.py:693-696 + 744-746: a replicate that stops early makes BSCG raise ValueError
    bs2 = reginv.BootStrap(MRANGE, MSPACING, obs_of(g), 0.02 * g["dobs"], (-5.0, 5.0), samples=2,
                           beta=0.05, maxk=6, verbose=False)
    with pytest.raises(ValueError):
        bs2.BSCG(np.zeros(192))


@pytest.mark.parametrize("reg,beta", [("Damping", 0.01), ("MS", 0.002), ("Smoothness", 0.01), ("TV", 0.001)])
def test_cg_vs_oracle_larger(reg, beta):
    """20 x 24 x 10 voxels x 30 x 20 observations (ragged: M = 4800 is not a multiple of the strips,
    N = 600 not of the row tiles), 12 iterations against the oracle."""
    mrange, msp = (0, 2000, 0, 2400, 0, 1000), (100, 100, 100)
    xs, ys = np.linspace(40, 1960, 30), np.linspace(60, 2340, 20)
    X, Y = np.meshgrid(xs, ys)
    xp, yp, zp = X.ravel(), Y.ravel(), np.full(X.size, -2.0)
    mesh = onp.OracleMesh(mrange, msp)
    _, A = onp.prism_gz(xp, yp, zp, mesh.active_bounds()[0], threads=4)
    rho = np.zeros((10, 24, 20))
    rho[2:5, 8:14, 6:12] = 0.8
    rng = np.random.default_rng(7)
    d0 = A @ rho.ravel()
    dobs = d0 + rng.normal(0, 0.02 * np.abs(d0).max(), d0.shape)
    cg = reginv.ConjugateGradient(dobs, mrange, msp, (xp, yp, zp), verbose=False)
    ocg = onp.OracleCG(A, dobs, (10, 24, 20))
    init, apr = np.full(4800, 0.01), 0.02 + 0.01 * np.cos(np.arange(4800) * 0.11)
    got = cg.CG(init, apr, (0.0, 0.7), regularization=reg, beta=beta, q=0.9, maxk=12)
    ref = ocg.CG(init, apr, (0.0, 0.7), reg, beta, 0.9, 12)
    assert len(got[4]) == len(ref[4])
    for a, b in zip(got, ref):
        assert rel(a, b) < TOL


def test_bootstrap_vs_oracle_larger():
    """17 replicates (Cp = 24) on 16 x 12 x 6 voxels x 150 observations against the oracle's gathered rows"""
    mrange, msp = (0, 1600, 0, 1200, 0, 600), (100, 100, 100)
    xs, ys = np.linspace(30, 1570, 15), np.linspace(30, 1170, 10)
    X, Y = np.meshgrid(xs, ys)
    xp, yp, zp = X.ravel(), Y.ravel(), np.full(X.size, -1.0)
    mesh = onp.OracleMesh(mrange, msp)
    _, A = onp.prism_gz(xp, yp, zp, mesh.active_bounds()[0], threads=4)
    rho = np.zeros((6, 12, 16))
    rho[1:4, 4:8, 5:11] = 0.6
    rng = np.random.default_rng(11)
    d0 = A @ rho.ravel()
    dobs = d0 + rng.normal(0, 0.03 * np.abs(d0).max(), d0.shape)
    bs = reginv.BootStrap(mrange, msp, (xp, yp, zp), dobs, (0.0, 0.5), samples=17, beta=0.04, maxk=7,
                          verbose=False)
    obs_ = onp.OracleBootStrap(A, dobs, (6, 12, 16), (0.0, 0.5), 17, 0.04, 7)
    got = bs.BSCG(np.full(1152, 0.005))
    ref = obs_.BSCG(np.full(1152, 0.005))
    for a, b in zip(got, ref):
        assert rel(a, b) < TOL


@pytest.mark.parametrize("wavelet", ["1D", "3D"])
def test_cg_wavelet_compressed_forward_vs_oracle(golden, wavelet):
    """reginv.py:107-117, 250-264: data terms through the wavelet-compressed kernel (PyWavelets
    conventions restated in the oracle -- parity unpinned), Aw @ Iw dense"""
    g = golden["reginv"]
    cg = reginv.ConjugateGradient(g["dobs"], MRANGE, MSPACING, obs_of(g), wavelet=wavelet, verbose=False)
    ocg = onp.OracleCG(cg.A.cpu().numpy(), g["dobs"], MSHAPE, wavelet=wavelet)
    mw = cg.Wm @ (0.3 * np.linspace(0, 1, 192))
    assert abs(cg.data(mw) - ocg.data(mw)) < 1e-9 * ocg.data(mw)
    assert rel(cg.data_gfun(mw), ocg.data_gfun(mw)) < 1e-9
    got = cg.CG(g["initial"], g["aprior"], g["boundary"], regularization="Damping", beta=0.01, q=0.9, maxk=8)
    ref = ocg.CG(g["initial"], g["aprior"], g["boundary"], "Damping", 0.01, 0.9, 8)
    for a, b in zip(got, ref):
        assert rel(a, b) < 1e-8
    # the compressed forward differs from the dense one (threshold 1e-3): the path is really taken
    dense = reginv.ConjugateGradient(g["dobs"], MRANGE, MSPACING, obs_of(g), verbose=False)
    assert abs(dense.data(mw) - cg.data(mw)) > 1e-9 * dense.data(mw)
