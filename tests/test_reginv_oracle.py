"""CPU: the oracle restatement of inversion/reginv.py (oracle_np.OracleCG / OracleBootStrap) against
the golden vectors produced by the UNMODIFIED reference (oracle/make_golden_reginv.py)."""
import numpy as np
import pytest

from oracle import oracle_np as onp

MRANGE, MSPACING, MSHAPE = (0, 800, 0, 600, 0, 400), (100, 100, 100), (4, 6, 8)
REGS = ("Damping", "MS", "Smoothness", "TV")


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module")
def setup(golden):
    g = golden["reginv"]
    obs = g["obs"]
    mesh = onp.OracleMesh(MRANGE, MSPACING)
    _, A = onp.prism_gz(obs[:, 0], obs[:, 1], obs[:, 2], mesh.active_bounds()[0])
    return g, A


@pytest.mark.parametrize("reg", REGS)
def test_oracle_cg_vs_reference(setup, reg):
    g, A = setup
    cg = onp.OracleCG(A, g["dobs"], MSHAPE)
    assert rel(cg.wm, g["cg_wm"]) < 1e-14
    assert rel(cg.Aw[[0, 17, 79]], g["cg_Aw_rows"]) < 1e-13
    m, d, dm, mm, rf = cg.CG(g["initial"], g["aprior"], g["boundary"], reg,
                             float(g["cg_%s_beta" % reg]), 0.9, 14)
    assert len(rf) == 14
    assert rel(m, g["cg_%s_model" % reg]) < 1e-11
    assert rel(d, g["cg_%s_data" % reg]) < 1e-11
    assert rel(dm, g["cg_%s_data_misfit" % reg]) < 1e-11
    assert rel(mm, g["cg_%s_model_misfit" % reg]) < 1e-11
    assert rel(rf, g["cg_%s_regul" % reg]) < 1e-11


def test_oracle_cg_early_stop(setup):
    g, A = setup
    cg = onp.OracleCG(A, g["cgstop_dobs"], MSHAPE)
    m, d, dm, mm, rf = cg.CG(g["initial"], np.zeros(A.shape[1]), (-5.0, 5.0), "Damping", 0.01, 0.5, 50)
    assert len(rf) == len(g["cgstop_regul"]) == 2 and len(dm) == 2
    assert rel(m, g["cgstop_model"]) < 1e-12 and rel(dm, g["cgstop_data_misfit"]) < 1e-12


def test_oracle_bootstrap_vs_reference(setup):
    g, A = setup
    bs = onp.OracleBootStrap(A, g["dobs"], MSHAPE, g["boundary"], 5, float(g["bs_beta"]), 9)
    mi, dmi, mmi, rfi = bs.BSCG(g["initial"])
    assert rel(mi, g["bs_models"]) < 1e-10
    assert rel(dmi, g["bs_data_misfit"]) < 1e-10
    assert rel(mmi, g["bs_model_misfit"]) < 1e-10
    assert rel(rfi, g["bs_regul"]) < 1e-12
    # the resampling indices are the legacy global-RNG stream the reference draws
    for s in range(5):
        idx = np.random.RandomState(s).choice(np.arange(80), size=80, replace=True)
        assert np.array_equal(idx, g["bs_index"][s])


def test_oracle_bootstrap_early_stop_raises(setup):
    g, A = setup
    assert int(g["bs_stop_raises"]) == 1
    bs = onp.OracleBootStrap(A, 0.02 * g["dobs"], MSHAPE, (-5.0, 5.0), 2, 0.05, 6)
    with pytest.raises(ValueError):
        bs.BSCG(np.zeros(A.shape[1]))
