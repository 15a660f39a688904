"""tools/run_example.py: the SetPMTS.txt-driven runner mirrors what the reference's example drivers do
with their parameter file (example/uniformgrid/main_uniform.py:98-119; main_global.py:22-28 reorder).
CPU part: parsing and the per-example geometry; the GPU part runs config 1 through it."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import run_example as rx  # noqa: E402

LINE = ('{"set": "model01_singlecube", "test": "T1", "rhomin": 0, "rhomax": 1, "mspacing": [100, 100, 100], '
        '"Lrange": [5, 20], "delta": 0.01, "Sigma": 0.001, "RegulFactor": 1, "regularization": "MS", '
        '"beta": 0.001, "nsamples": 500}')


def stage_c1(tmp_path, g):
    d = tmp_path / "uniformgrid"
    (d / "modeldata").mkdir(parents=True)
    (d / "SetPMTS.txt").write_text(LINE + "\n\n" + LINE.replace('"T1"', '"T2"').replace('"MS"', '"Damping"') + "\n")
    np.savetxt(d / "modeldata" / "model01_singlecube_gz_noise.txt", np.c_[g["obs"], g["dobs"]], fmt="%.18e")
    return str(d)


def test_parse_and_geometry(tmp_path):
    from tests import chains200 as c2h

    d = stage_c1(tmp_path, c2h.load("c1_MS"))
    pm = rx.parse_setpmts(os.path.join(d, "SetPMTS.txt"))
    assert len(pm) == 2 and pm[0]["regularization"] == "MS" and pm[1]["test"] == "T2"
    assert pm[0]["Lrange"] == [5, 20] and pm[0]["nsamples"] == 500
    assert rx.example_kind(d) == "uniformgrid"
    with pytest.raises(ValueError):
        rx.example_kind(str(tmp_path))
    # the spacing each driver hands to GravMagModule
    assert rx.model_spacing("uniformgrid", [100, 100, 100]) == (100, 100, 100)
    assert rx.model_spacing("global", [3, 3, -300000]) == (-300000, 3, 3)           # main_global.py:22-28
    assert rx.model_spacing("segmentgrid", [100, 100, [100, 200, 300]]) == ([100, 200, 300], 100, 100)
    assert rx.model_spacing("realdata", [[-1000, -2000, -5000], 0.5, 0.5]) == ([-1000, -2000, -5000], 0.5, 0.5)
    (tmp_path / "bad.txt").write_text('{"set": "x"}\n')
    with pytest.raises(ValueError, match="lacks"):
        rx.parse_setpmts(str(tmp_path / "bad.txt"))
    (tmp_path / "evil.txt").write_text('__import__("os").system("true")\n')
    with pytest.raises(ValueError):
        rx.parse_setpmts(str(tmp_path / "evil.txt"))  # literal_eval, not eval


@pytest.mark.gpu
def test_runs_config1_like_the_reference(tmp_path):
    """config 1 as shipped (MS, SetPMTS line 0) minus the wavelet (PyWavelets absent upstream): the
    first 12 samples equal the unmodified reference's (tests/golden/chains200_c1_MS.npz)"""
    from tests import chains200 as c2h

    g = c2h.load("c1_MS")
    d = stage_c1(tmp_path, g)
    r = rx.run(d, 0, nsamples=12, wavelet="none", quiet=True)
    assert r["shape"] == (10, 30, 20) and r["observations"] == 600
    mis = np.loadtxt(os.path.join(r["save_folder"] + "0", "misfit.dat"), ndmin=2)
    assert mis.shape == (12, 7) and np.allclose(mis, g["misfit"][:12], rtol=0, atol=2e-8)
    assert [(L, bool(a)) for L, a in r["proposals"][0]] == [(int(L), bool(a)) for L, a in g["log"][:12, :2]]
    # two chains as one batch: rank 0 is the same chain
    r2 = rx.run(d, 0, nsamples=5, chains=2, wavelet="none", quiet=True, out=str(tmp_path / "b"))
    mis2 = np.loadtxt(os.path.join(r2["save_folder"] + "0", "misfit.dat"), ndmin=2)
    assert np.allclose(mis2, g["misfit"][:5], rtol=0, atol=2e-8)
