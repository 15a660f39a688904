"""CPU: gi_legacy_randn_scaled continues a numpy legacy RandomState bit for bit -- the momentum draws of
the samplers must be the reference's (`randn(n) * Sigma` between `randint` and `rand`,
inversion/hmc.py:297, 95, 165), whichever side generates them."""
import numpy as np
import pytest

from gravinv3dhmc_b200 import _lib


@pytest.mark.parametrize("seed", [0, 1, 100, 163, 2 ** 31 + 5])
def test_stream_is_numpys(seed):
    a, b = np.random.RandomState(seed), np.random.RandomState(seed)
    for n, scale in ((1, 1.0), (7, 0.001), (1000, 0.5), (625, 2.0), (3, 1.0), (0, 1.0), (100001, 0.001)):
        assert a.randint(5, 21) == b.randint(5, 21)
        ref = a.randn(n) * scale
        out = np.full(n, np.nan)
        _lib.legacy_randn_scaled(b, n, scale, out)
        assert np.array_equal(ref, out)          # odd n leaves a cached deviate behind: also handled
        assert a.rand() == b.rand()
    assert np.array_equal(a.randn(5), b.randn(5))  # numpy itself continues from the written-back state
    sa, sb = a.get_state(), b.get_state()
    assert np.array_equal(sa[1], sb[1]) and sa[2:] == sb[2:]


def test_bad_arguments():
    rs = np.random.RandomState(0)
    with pytest.raises(ValueError):
        _lib.check(_lib.lib().gi_legacy_randn_scaled(None, None, None, None, 4, 1.0, None), "x")
