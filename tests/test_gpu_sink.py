"""GPU parity of the sample sinks: the on-device running mean / std of the accepted
models (gi_stats_*) against np.mean / np.std of (a) the CPU oracle's unrounded samples (1e-10
relative to the model scale) and (b) the "%.8f" rows of the reference-format model.dat (2e-8
absolute: the text rounding), and the binary sample files against the text files."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from gravinv3dhmc_b200.inversion import batched, hmc, potential, sink  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402


def small_model(g):
    o = g["small_obs"]
    return potential.GravMagModule(g["small_dobs"], (0, 400, 0, 600, 0, 500), (100, 100, 100),
                                   (o[:, 0], o[:, 1], o[:, 2]), verbose=False)


def bounds(M, lo=-5.0, hi=5.0):
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = lo, hi
    return b


def test_stats_accumulator_vs_numpy(golden):
    """Welford accumulators: window (skip / take), per-slot and pooled results, reset"""
    model = small_model(golden["potential_hmc"])
    M = model.M
    rng = np.random.default_rng(5)
    wminv = model.WmInv.diagonal()
    samples = [rng.normal(0.3, 0.2, (n, M)) * (1.0 + np.arange(M))[None, :] for n in (9, 14, 1)]
    sk = sink.SampleSink(model, nslots=3, skip=2, take=10)
    for slot, S in enumerate(samples):
        for row in S:
            sk.add(row, slot)
    kept = [S[2:12] for S in samples]
    for slot, K in enumerate(kept):
        mean, std, cnt = sk.result(slot)
        assert cnt == K.shape[0] and sk.seen == samples[slot].shape[0]
        if cnt == 0:
            assert not mean.any() and not std.any()
            continue
        ref = K * wminv[None, :]
        scale = np.abs(ref).max()
        assert np.max(np.abs(mean - ref.mean(axis=0))) < 1e-13 * scale
        assert np.max(np.abs(std - ref.std(axis=0))) < 1e-12 * scale
    mean, std, cnt = sk.result()  # pooled over the chains
    allk = np.concatenate(kept) * wminv[None, :]
    assert cnt == allk.shape[0]
    assert np.max(np.abs(mean - allk.mean(axis=0))) < 1e-13 * np.abs(allk).max()
    assert np.max(np.abs(std - allk.std(axis=0))) < 1e-12 * np.abs(allk).max()
    dm, ds = sk.forward()
    A = (model.Aw * model.wm_dev[:M][None, :]).cpu().numpy()
    assert np.max(np.abs(dm - A @ allk.mean(axis=0))) < 1e-11 * np.abs(dm).max()
    assert np.max(np.abs(ds - A @ allk.std(axis=0))) < 1e-11 * np.abs(ds).max()
    sk.reset()
    assert sk.result()[2] == 0
    assert sk.launches() > 0
    sk.close()


@pytest.mark.parametrize("output", ["text", "binary", "none"])
def test_single_chain_sink_and_file_formats(golden, tmp_path, output):
    g = golden["potential_hmc"]
    model = small_model(g)
    M = model.M
    nsamples, ndraws = 9, 3
    args = (0.1, [4, 9], np.ones(M) * 0.001, np.ones(M) * 0.001, bounds(M), "mandatory", 1000,
            g["small_dobs"], "Fixed", 0.8, 1.0, "Damping", 0.001, 3, 1.0)
    chain = hmc.setup_chain(model, *args, save_folder=str(tmp_path / "c"), quiet=True)
    chain.output = output
    chain.sink = sink.SampleSink(model)
    chain.sample(nsamples, ndraws)
    om = onp.OracleModel(g["small_Aw"], g["small_wm"], g["small_dobs"], tuple(g["small_mshape"]))
    ref = onp.hmc_sample(om, nsamples, ndraws, args[0], args[1], args[2], args[3], args[4], "mandatory",
                         1000, 1.0, "Damping", 0.001, 3, 1.0)
    assert ref["models"].shape == (nsamples, M)
    mean, std, cnt = chain.sink.result(0)
    assert cnt == nsamples and chain.sink.seen == nsamples + ndraws
    scale = np.abs(ref["models"]).max()
    assert np.max(np.abs(mean - ref["models"].mean(axis=0))) < 1e-10 * scale
    assert np.max(np.abs(std - ref["models"].std(axis=0))) < 1e-10 * scale
    folder = tmp_path / "c0"
    if output == "none":
        assert not (folder / "model.dat").exists() and not (folder / "model.f64").exists()
    else:
        mis, mod = sink.read_samples(str(folder))
        assert mis.shape == (nsamples, 7) and mod.shape == (nsamples, M)
        tol = 2e-8 if output == "text" else 1e-9 * scale
        assert np.allclose(mod, ref["models"], rtol=0, atol=tol)
        assert np.allclose(mis, ref["misfit"], rtol=1e-9 if output == "binary" else 0, atol=2e-8)
        fmean, fstd = sink.posterior_from_samples(mod)  # plot_uniform.py:103-104 on the files
        assert np.allclose(mean, fmean, rtol=0, atol=2e-8) and np.allclose(std, fstd, rtol=0, atol=2e-8)
        last_mis, last_mod = sink.read_samples(str(folder), last=4)
        assert np.array_equal(last_mod, mod[-4:]) and np.array_equal(last_mis, mis[-4:])
    # plot_uniform.py:117-131 files
    mean2, std2, dmean, dstd = chain.sink.save(str(tmp_path / "post"), 0)
    im = np.loadtxt(tmp_path / "post" / "inversion_model.dat")
    ia = np.loadtxt(tmp_path / "post" / "inversion_anomaly.dat")
    assert im.shape == (M, 5) and ia.shape == (g["small_dobs"].size, 6)
    assert np.allclose(im[:, 3], mean, atol=1e-8, rtol=0) and np.allclose(im[:, 4], std, atol=1e-8, rtol=0)
    A = g["small_Aw"] * g["small_wm"][None, :]
    assert np.allclose(ia[:, 3], A @ ref["models"].mean(axis=0), atol=2e-8, rtol=0)
    assert np.allclose(ia[:, 5], g["small_dobs"] - ia[:, 3], atol=2e-8, rtol=0)
    assert np.array_equal(im[:4, 0], [0.0, 100.0, 200.0, 300.0])  # x of the first cells
    chain.close()


@pytest.mark.parametrize("mode,output", [("stream", "none"), ("stream", "text"), ("lockstep", "binary")])
def test_batch_sink_matches_oracle_chains(golden, tmp_path, mode, output):
    """every chain's device statistics equal those of the oracle chain `myrank = c`; with output
    'none' the streaming sampler copies no position back"""
    g = golden["potential_hmc"]
    model = small_model(g)
    M = model.M
    nchains, nsamples, ndraws = 5, 6, 2
    bt = batched.HMCBatch(model, nchains, 0.1, [4, 9], np.ones(M) * 0.001, np.ones(M) * 0.001, bounds(M),
                          "mandatory", 1000, g["small_dobs"], 1.0, "Damping", 0.001, 3, 1.0,
                          save_folder=str(tmp_path / "b"), quiet=True)
    bt.output = output
    bt.sink = sink.SampleSink(model, nslots=nchains)
    (bt.stream if mode == "stream" else bt.sample)(nsamples, ndraws)
    om = onp.OracleModel(g["small_Aw"], g["small_wm"], g["small_dobs"], tuple(g["small_mshape"]))
    allm = []
    for c in range(nchains):
        ref = onp.hmc_sample(om, nsamples, ndraws, 0.1, [4, 9], np.ones(M) * 0.001, np.ones(M) * 0.001,
                             bounds(M), "mandatory", 1000, 1.0, "Damping", 0.001, 3, 1.0, myrank=c)
        allm.append(ref["models"])
        mean, std, cnt = bt.sink.result(c)
        scale = np.abs(ref["models"]).max()
        assert cnt == nsamples
        assert np.max(np.abs(mean - ref["models"].mean(axis=0))) < 1e-10 * scale
        assert np.max(np.abs(std - ref["models"].std(axis=0))) < 1e-10 * scale
        assert [(L, bool(a)) for L, a in bt.proposals[c]] == [(L, bool(a)) for L, a in ref["log"]]
        if mode == "stream" and output != "none":
            # (output 'none': bt.x is the device's current state, which may already include a queued
            # proposal that ran after the chain reached its target)
            assert np.max(np.abs(bt.x[c] - ref["x"])) < 1e-9 * np.max(np.abs(ref["x"]))
        if output != "none":
            mis, mod = sink.read_samples(str(tmp_path / f"b{c}"))
            assert np.allclose(mod, ref["models"], rtol=0, atol=2e-8)
            assert np.allclose(mis, ref["misfit"], rtol=0, atol=2e-8)
    allm = np.concatenate(allm)
    mean, std, cnt = bt.sink.result()
    assert cnt == nchains * nsamples
    assert np.max(np.abs(mean - allm.mean(axis=0))) < 1e-10 * np.abs(allm).max()
    assert np.max(np.abs(std - allm.std(axis=0))) < 1e-10 * np.abs(allm).max()
    bt.close()
