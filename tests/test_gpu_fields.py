"""GPU parity of the other prism fields (gi_prism_field_assemble) against golden
vectors from the UNMODIFIED reference and against the CPU oracle on random prisms.

Tolerance (north_star: G entries and forward data 1e-10): kernel matrices and forward results 1e-10
normwise (max |diff| / max |ref|) -- the corner terms (up to x*y*log r ~ 1e7 for the potential)
cancel in the 8-corner sum and CUDA's libm differs from glibc's by an ulp, exactly like the gz
kernel; measured: 2e-12 for the potential, < 1e-13 for the gradient components."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from gravinv3dhmc_b200 import mesher, utils  # noqa: E402
from gravinv3dhmc_b200.gravmag import prism  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402
from tests.helpers import synthetic_topo  # noqa: E402

MRANGE, MSPACING = (0, 400, 0, 600, 0, 500), (100, 100, 100)
GRAV = ("potential", "geoid", "gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz")


def nrm(a, b):
    assert a.shape == b.shape
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def obs_of(g):
    o = g["obs"]
    return o[:, 0].copy(), o[:, 1].copy(), o[:, 2].copy()


@pytest.mark.parametrize("field", GRAV)
def test_gravity_fields_vs_reference_golden(golden, field):
    g = golden["fields"]
    xp, yp, zp = obs_of(g)
    mesh = mesher.PrismMesh(MRANGE, MSPACING)
    mesh.addprop("density", g["dens"])
    res, K = getattr(prism, field)(xp, yp, zp, mesh)
    assert K.shape == g[field + "_kernel"].shape
    assert np.isfinite(K).all()
    assert nrm(K, g[field + "_kernel"]) < 1e-10
    assert nrm(res, g[field + "_result"]) < 1e-10
    # sensitivity use: dens overrides the property (prism.py:138-141)
    res1, K1 = getattr(prism, field)(xp, yp, zp, mesh, dens=1.0)
    assert np.array_equal(K1, K) and nrm(res1, g[field + "_kernel"].sum(axis=1)) < 1e-10


def test_magnetic_fields_vs_reference_golden(golden):
    g = golden["fields"]
    xp, yp, zp = obs_of(g)
    inc, dec = g["inc_dec"]
    mesh = mesher.PrismMesh(MRANGE, MSPACING)
    mesh.addprop("magnetization", g["mag"])
    res, K = prism.tf(xp, yp, zp, mesh, inc, dec)
    assert nrm(K, g["tf_kernel"]) < 1e-10 and nrm(res, g["tf_result"]) < 1e-10
    res, K = prism.tf(xp, yp, zp, mesh, inc, dec, pmag=2.5)
    assert nrm(K, g["tf_scalar_kernel"]) < 1e-10 and nrm(res, g["tf_scalar_result"]) < 1e-10
    for comp in ("bx", "by", "bz"):
        assert nrm(getattr(prism, comp)(xp, yp, zp, mesh), g[comp + "_result"]) < 1e-10
    assert nrm(prism.bx(xp, yp, zp, mesh, pmag=[0.3, -1.2, 2.0]), g["bx_pmag_result"]) < 1e-10
    # utils.ang2vec / dircos (utils.py:420-474)
    assert np.allclose(utils.ang2vec(1.5 + 0.01 * np.arange(mesh.size), 40.0, 25.0), g["mag"], rtol=1e-15)


def test_carved_mesh_and_errors(golden):
    g = golden["fields"]
    xp, yp, zp = obs_of(g)
    mesh = mesher.PrismMesh(MRANGE, MSPACING)
    t = g["carved_topo"]
    mesh.carvetopo(t[:, 0], t[:, 1], t[:, 2])
    mesh.addprop("density", np.zeros(mesh.size))
    _, K = prism.gzz(xp, yp, zp - 200.0, mesh)
    assert K.shape == g["carved_gzz_kernel"].shape and nrm(K, g["carved_gzz_kernel"]) < 1e-10
    with pytest.raises(ValueError):
        prism.gxx(xp[:-1], yp, zp, mesh)
    # a mesh without the property and no dens: nothing to compute (prism.py:136-137)
    bare = mesher.PrismMesh(MRANGE, MSPACING)
    res, K = prism.gx(xp, yp, zp, bare)
    assert K.shape == (xp.size, 0) and not res.any()


@pytest.mark.parametrize("field", ["potential", "gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz", "tf"])
def test_fields_vs_oracle_ragged(field):
    """1000 x 37-cell segmented mesh rows x 333 observations (ragged against the 128-column CTAs and
    8-row tiles), observations on and off the mesh nodes"""
    mesh = mesher.PrismMeshSegment((0, 3700, 0, 300, 0, 2100), ([100, 200, 300], 100, 100), [0, 300, 900, 2100])
    rng = np.random.RandomState(9)
    n = 333
    xp = np.concatenate([rng.uniform(-100, 3800, n - 40), np.arange(40) * 100.0])
    yp = np.concatenate([rng.uniform(-50, 350, n - 40), np.full(40, 100.0)])
    zp = np.concatenate([rng.uniform(-200, -1, n - 40), np.full(20, 0.0), np.full(20, 300.0)])
    tab = mesh.bounds_table()
    dens = rng.uniform(-1, 1, mesh.size)
    if field == "tf":
        mag = utils.ang2vec(np.abs(dens) + 0.1, 30.0, -40.0)
        mesh.addprop("magnetization", mag)
        res, K = prism.tf(xp, yp, zp, mesh, 60.0, 10.0)
        ores, oK = onp.prism_field("tf", xp, yp, zp, tab, inc=60.0, dec=10.0, mag=mag, threads=4)
    else:
        mesh.addprop("density", dens)
        res, K = getattr(prism, field)(xp, yp, zp, mesh)
        ores, oK = onp.prism_field(field, xp, yp, zp, tab, dens=dens, threads=4)
    assert nrm(K, oK) < 1e-10
    assert nrm(res, ores) < 1e-10


def test_magnetic_module_and_chain_vs_reference_golden(golden, tmp_path):
    """GravMagModule(coordinate="cartesian", field="magnetic") (potential.py:125-149): weighted tf
    kernel, misfit_and_grad and a short chain against the unmodified reference"""
    from gravinv3dhmc_b200.inversion import hmc, potential

    g = golden["magnetic"]
    o = g["obs"]
    inc, dec = g["mangle"]
    model = potential.GravMagModule(g["dobs"], (0, 400, 0, 600, 0, 500), (100, 100, 100),
                                    (o[:, 0], o[:, 1], o[:, 2]), coordinate="cartesian", field="magnetic",
                                    mangle=(inc, dec), verbose=False)
    assert nrm(model.Aw.cpu().numpy(), g["Aw"]) < 1e-10
    assert np.allclose(model.Wm.diagonal(), g["wm"], rtol=1e-10)
    U, grad, dpre, Ud, Um = model.misfit_and_grad(g["mg_x"], g["mg_x0"], None, None, "mandatory", 1000, 0.7,
                                                  regulization="MS", beta=0.001)
    assert np.allclose([U, Ud, Um], g["mg_scalars"], rtol=1e-9)
    assert nrm(grad, g["mg_grad"]) < 1e-9 and nrm(dpre, g["mg_dpre"]) < 1e-9
    M = model.M
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = 0.0, 3.0
    ch = hmc.HMCSample(model, 8, 0, 0.02, [3, 8], np.ones(M) * 0.001, np.ones(M) * 0.001, b, "mandatory",
                       1000, g["dobs"], "Fixed", 0.8, 1.0, "Damping", 0.001, 21, 0.05, myrank=0,
                       save_folder=str(tmp_path / "mag"), quiet=True)
    log = g["chain_prop_log"]
    assert [(L, int(a)) for L, a in ch.proposals] == [(int(L), int(a)) for L, a in log[:, :2]]
    mis = np.loadtxt(tmp_path / "mag0" / "misfit.dat", ndmin=2)
    mod = np.loadtxt(tmp_path / "mag0" / "model.dat", ndmin=2)
    assert np.allclose(mis, g["chain_misfit"], rtol=0, atol=2e-8)
    assert np.allclose(mod, g["chain_models"], rtol=0, atol=2e-8)
    with pytest.raises(ValueError):
        potential.GravMagModule(g["dobs"], (0, 10, 0, 10, 0, -1000), (-500, 5, 5), (o[:, 0], o[:, 1], o[:, 2]),
                                coordinate="spherical", field="magnetic", verbose=False)


def test_field_edge_cases():
    """empty and degenerate inputs of the field builders (prism.py:136-137 skips cells without the
    property; zero observations / zero cells must not launch anything)"""
    from gravinv3dhmc_b200.gravmag import tesseroid

    mesh = mesher.PrismMesh(MRANGE, MSPACING)
    mesh.addprop("density", np.ones(mesh.size))
    e = np.zeros(0)
    res, K = prism.gyz(e, e, e, mesh)
    assert res.shape == (0,) and K.shape == (0, mesh.size)
    res, K = prism.tf(e, e, e, mesh, 10.0, 20.0, pmag=1.0)
    assert res.shape == (0,) and K.shape == (0, mesh.size)
    one = np.array([123.0]), np.array([77.0]), np.array([-3.0])
    res, K = prism.potential(*one, [None, None])          # a model of masked cells only
    assert K.shape == (1, 0) and res[0] == 0.0
    tm = mesher.TesseroidMesh((-2, 2, -2, 2, 0, -20000), (-10000, 2, 2))
    res, K = tesseroid.gxz(e, e, e, tm, dens=1.0)
    assert res.shape == (0,) and K.shape == (0, tm.size)
    # a single (observation, cell) pair straight above the cell centre: gxz = gyz = 0 by symmetry
    cell = mesher.PrismMesh((0, 100, 0, 100, 0, 100), (100, 100, 100))
    cell.addprop("density", np.ones(1))
    c = np.array([50.0]), np.array([50.0]), np.array([-40.0])
    for f in ("gxz", "gyz", "gxy", "gx", "gy"):
        assert abs(getattr(prism, f)(*c, cell)[0][0]) < 1e-9 * abs(prism.gzz(*c, cell)[0][0])
    # Laplace: gxx + gyy + gzz = 0 outside the body
    lap = sum(getattr(prism, f)(*c, cell)[0][0] for f in ("gxx", "gyy", "gzz"))
    assert abs(lap) < 1e-9 * abs(prism.gzz(*c, cell)[0][0])
