"""CPU: the oracle restatement of the other prism fields (oracle/csrc/oracle_fields.c) against golden
vectors produced by the UNMODIFIED reference (oracle/make_golden_fields.py) -- and, where the compiled
reference extension is present (oracle/_ref), against it live on random prisms."""
import numpy as np
import pytest

from oracle import oracle_np as onp, ref_harness

MRANGE, MSPACING = (0, 400, 0, 600, 0, 500), (100, 100, 100)
GRAV = ("potential", "geoid", "gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz")


def nrm(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module")
def setup(golden):
    g = golden["fields"]
    tab, _ = onp.OracleMesh(MRANGE, MSPACING).active_bounds()
    o = g["obs"]
    return g, tab, (o[:, 0].copy(), o[:, 1].copy(), o[:, 2].copy())


@pytest.mark.parametrize("field", GRAV)
def test_gravity_fields_bitwise(setup, field):
    g, tab, (xp, yp, zp) = setup
    res, K = onp.prism_field(field, xp, yp, zp, tab, dens=g["dens"])
    # same operations in the same order as the compiled reference: identical bits
    assert np.array_equal(K, g[field + "_kernel"])
    assert np.array_equal(res, g[field + "_result"])


def test_magnetic_fields(setup):
    g, tab, (xp, yp, zp) = setup
    inc, dec = g["inc_dec"]
    res, K = onp.prism_field("tf", xp, yp, zp, tab, inc=inc, dec=dec, mag=g["mag"])
    assert np.array_equal(K, g["tf_kernel"]) and np.array_equal(res, g["tf_result"])
    f = np.array(onp.dircos(inc, dec))
    res, K = onp.prism_field("tf", xp, yp, zp, tab, inc=inc, dec=dec, mag=np.tile(2.5 * f, (tab.shape[0], 1)))
    assert np.array_equal(K, g["tf_scalar_kernel"]) and np.array_equal(res, g["tf_scalar_result"])
    for comp in ("bx", "by", "bz"):
        res, K = onp.prism_field(comp, xp, yp, zp, tab, mag=g["mag"])
        assert K is None and np.array_equal(res, g[comp + "_result"])
    res, _ = onp.prism_field("bx", xp, yp, zp, tab, mag=np.tile([0.3, -1.2, 2.0], (tab.shape[0], 1)))
    assert np.array_equal(res, g["bx_pmag_result"])


def test_carved_columns(setup, golden):
    g, _, (xp, yp, zp) = setup
    m = onp.OracleMesh(MRANGE, MSPACING)
    t = g["carved_topo"]
    m.carvetopo(t[:, 0], t[:, 1], t[:, 2])
    tab, _ = m.active_bounds()
    _, K = onp.prism_field("gzz", xp, yp, zp - 200.0, tab)
    assert K.shape == g["carved_gzz_kernel"].shape and np.array_equal(K, g["carved_gzz_kernel"])


@pytest.mark.skipif(ref_harness.prism_so_path() is None, reason="compiled reference _prism not built")
def test_against_compiled_reference_random_prisms():
    ref = ref_harness.load_prism_ext()
    rng = np.random.RandomState(4)
    n = 40
    xp, yp, zp = rng.uniform(-500, 500, n), rng.uniform(-500, 500, n), rng.uniform(-300, 100, n)
    for _ in range(6):
        lo = rng.uniform(-300, 200, 3)
        hi = lo + rng.uniform(10, 300, 3)
        b = np.array([[lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]]])
        for field in ("potential", "gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz"):
            res, k1 = np.zeros(n), np.zeros(n)
            getattr(ref, field)(xp, yp, zp, b[0, 0], b[0, 1], b[0, 2], b[0, 3], b[0, 4], b[0, 5], 1.0, res, k1)
            _, K = onp.prism_field(field, xp, yp, zp, b)
            assert np.array_equal(K[:, 0], k1 * onp.prism_field_scale(field)), field
