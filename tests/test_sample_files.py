"""CPU: the sample-file sinks of inversion/sink.py -- the reference's text format (hmc.py:241-249:
one "%.8f" row per accepted sample, model.dat removed at the start of a run, misfit.dat appended to)
and the raw float64 format, and the reader / posterior helpers built on them
(example/uniformgrid/plot_uniform.py:44-54, 103-104)."""
import json
import os

import numpy as np
import pytest

from gravinv3dhmc_b200.inversion.sink import SampleWriter, posterior_from_samples, read_samples


def rows(n, M, seed=0):
    rs = np.random.RandomState(seed)
    return rs.rand(n, 7) * 100.0, rs.randn(n, M)


@pytest.mark.parametrize("mode", ["text", "binary"])
def test_writer_reader_roundtrip(tmp_path, mode):
    M = 37
    mis, mod = rows(6, M)
    folder = str(tmp_path / "chain0")
    w = SampleWriter(folder, mode, M)
    for a, b in zip(mis, mod):
        w.append(a, b)
    got_mis, got_mod = read_samples(folder)
    assert got_mis.shape == (6, 7) and got_mod.shape == (6, M)
    if mode == "text":
        assert np.allclose(got_mis, mis, rtol=0, atol=5e-9) and np.allclose(got_mod, mod, rtol=0, atol=5e-9)
        # byte-compatible with the reference's writer: np.savetxt(fmt="%.8f", delimiter=" ")
        line = open(os.path.join(folder, "model.dat")).readline().split()
        assert line == ["%.8f" % v for v in mod[0]]
    else:
        assert np.array_equal(got_mis, mis) and np.array_equal(got_mod, mod)
        meta = json.load(open(os.path.join(folder, "samples.json")))
        assert meta["M"] == M and meta["dtype"] == "<f8" and len(meta["misfit_columns"]) == 7
    last_mis, last_mod = read_samples(folder, last=2)
    assert np.array_equal(last_mod, got_mod[-2:]) and np.array_equal(last_mis, got_mis[-2:])
    mean, std = posterior_from_samples(got_mod)
    assert np.allclose(mean, got_mod.mean(axis=0)) and np.allclose(std, got_mod.std(axis=0))


def test_rerun_semantics_of_the_reference(tmp_path):
    """hmc.py:257-258: a new run removes model.dat but APPENDS to a stale misfit.dat"""
    M = 5
    folder = str(tmp_path / "c")
    mis, mod = rows(3, M, 1)
    w = SampleWriter(folder, "text", M)
    for a, b in zip(mis, mod):
        w.append(a, b)
    w2 = SampleWriter(folder, "text", M)
    w2.append(mis[0], mod[0])
    assert np.loadtxt(os.path.join(folder, "model.dat"), ndmin=2).shape == (1, M)
    assert np.loadtxt(os.path.join(folder, "misfit.dat"), ndmin=2).shape == (4, 7)


def test_none_mode_and_errors(tmp_path):
    w = SampleWriter(str(tmp_path / "n"), "none", 4)
    w.append(np.zeros(7), np.zeros(4))
    assert not os.path.exists(str(tmp_path / "n"))
    with pytest.raises(ValueError):
        SampleWriter(str(tmp_path / "x"), "parquet", 4)
