"""CPU: the oracle restatement of the tesseroid fields other than gz and of the forward-only module
(oracle/csrc/oracle_tess.c: kernel_field / oracle_tess_field) against golden vectors produced by the
UNMODIFIED reference (oracle/make_golden_tessfields.py)."""
import numpy as np
import pytest

from oracle import oracle_np as onp

FIELDS = ("potential", "geoid", "gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz")


def nrm(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("field", FIELDS)
def test_tess_fields_vs_reference(golden, field):
    g = golden["tessfields"]
    o = g["obs"]
    tab, _ = onp.OracleMesh((-10, 10, -10, 10, 0, -300000), (-100000, 5, 5), zdown=False).active_bounds()
    res, K, err = onp.tess_field(field, o[:, 0], o[:, 1], o[:, 2], tab, dens=g["dens"])
    assert np.array_equal(K, g[field + "_kernel"])       # same operations, same order: same bits
    assert nrm(res, g[field + "_result"]) < 1e-13        # (per-leaf vs per-cell density factor)
    if field != "geoid":
        fres, _, _ = onp.tess_field(field, o[:, 0], o[:, 1], o[:, 2], tab, dens=g["dens"], forward=True)
        assert nrm(fres, g["fwd_" + field]) < 1e-13
