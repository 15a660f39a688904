"""GPU parity of `JointModule` (joint gz + total-field inversion on one mesh,
inversion/potential.py:847-1812) against golden vectors from the UNMODIFIED reference
(tests/golden/joint.npz, oracle/make_golden_joint.py): the weighted block kernel, `weightKDM`'s model
and data weights, `misfit_and_grad` for the regularisers the reference's class supports (MS, MS1,
MStry, Damping), its failure modes (Smoothness / TV: AttributeError, spherical: UnboundLocalError) and
8-sample chains through the single-chain and the batched samplers.  Tolerances: kernel and weights
1e-10, misfit_and_grad 1e-9, chains: identical decisions, per-leapfrog x and U 1e-9."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from gravinv3dhmc_b200.inversion import batched, hmc, potential  # noqa: E402


def nrm(a, b):
    return np.max(np.abs(np.asarray(a) - b)) / np.max(np.abs(b))


@pytest.fixture(scope="module")
def joint(golden):
    g = golden["joint"]
    o = g["obs"]
    inc, dec = g["mangle"]
    model = potential.JointModule(g["dobs_gz"], g["dobs_tf"], (0, 400, 0, 600, 0, 500), (100, 100, 100),
                                  (o[:, 0], o[:, 1], o[:, 2]), coordinate="cartesian", njobs=1,
                                  mangle=(inc, dec), wavelet=False, verbose=False)
    return g, model


def test_block_kernel_and_weights(joint):
    g, model = joint
    n = g["obs"].shape[0]
    assert model.Aw.shape == g["Aw"].shape == (2 * n, 2 * model.Mcells)
    Aw = model.Aw.cpu().numpy()
    assert nrm(Aw, g["Aw"]) < 1e-10
    assert np.all(Aw[:n, model.Mcells:] == 0) and np.all(Aw[n:, : model.Mcells] == 0)  # block structure
    assert np.allclose(model.Wm.diagonal(), g["wm"], rtol=1e-10)
    assert np.allclose(model.Wb.diagonal(), g["wb"], rtol=1e-10)
    assert np.allclose(model.dobsw, g["dobsw"], rtol=1e-10)
    assert nrm(model.forward(g["m_true"]), g["forward_true"]) < 1e-10
    Awk, WmInv, Wm = model.kernelw()
    assert Awk is model.Aw and Wm is model.Wm


@pytest.mark.parametrize("reg", ["MS", "MS1", "MStry", "Damping"])
def test_misfit_and_grad(joint, reg):
    g, model = joint
    U, grad, dpre, Ud, Um = model.misfit_and_grad(g["mg_x"], g["mg_x0"], None, None, "mandatory", 1000, 0.7,
                                                  regulization=reg, beta=0.001)
    assert np.allclose([U, Ud, Um], g["mg_%s_scalars" % reg], rtol=1e-9)
    assert nrm(grad, g["mg_%s_grad" % reg]) < 1e-9 and nrm(dpre, g["mg_%s_dpre" % reg]) < 1e-9


def test_reference_failure_modes(joint):
    g, model = joint
    assert str(g["err_Smoothness"]) == str(g["err_TV"]) == "AttributeError"
    for reg in ("Smoothness", "TV"):
        with pytest.raises(AttributeError):
            model.misfit_and_grad(g["mg_x"], g["mg_x0"], None, None, "mandatory", 1000, 0.7, regulization=reg)
    with pytest.raises(ValueError, match="regularization"):
        model.misfit_and_grad(g["mg_x"], g["mg_x0"], None, None, "mandatory", 1000, 0.7, regulization="L1")
    with pytest.raises(ValueError, match="boundary constraint"):
        model.misfit_and_grad(g["mg_x"], g["mg_x0"], None, None, "periodic", 1000, 0.7)
    assert str(g["err_spherical"]) == "UnboundLocalError"
    o = g["obs"]
    with pytest.raises(UnboundLocalError):
        potential.JointModule(g["dobs_gz"], g["dobs_tf"], (0, 10, 0, 10, 0, -1000), (-500, 5, 5),
                              (o[:, 0], o[:, 1], o[:, 2]), coordinate="spherical", verbose=False)
    with pytest.raises(ValueError):
        potential.JointModule(g["dobs_gz"], g["dobs_tf"], (0, 400, 0, 600, 0, 500), (100, 100, 100),
                              (o[:, 0], o[:, 1], o[:, 2]), coordinate="polar", verbose=False)


@pytest.mark.parametrize("reg,alpha", [("Damping", 1.0), ("MS", 0.5), ("MStry", 0.5)])
def test_chain_matches_reference(joint, reg, alpha, tmp_path):
    g, model = joint
    M = model.M
    dobs = np.append(g["dobs_gz"], g["dobs_tf"])
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = 0.0, 3.0
    traces = []
    real = hmc.HamitonianMC._leapfrog

    def traced(self, xcur, dt, L, a, fignum=0, trace=None):
        tr = {}
        out = real(self, xcur, dt, L, a, fignum, trace=tr)
        traces.append(tr)
        return out

    hmc.HamitonianMC._leapfrog = traced
    try:
        ch = hmc.HMCSample(model, 8, 0, 0.02, [3, 8], np.ones(M) * 0.001, np.ones(M) * 0.001, b, "mandatory",
                           1000, dobs, "Fixed", 0.8, alpha, reg, 0.001, 21, 0.05, myrank=0,
                           save_folder=str(tmp_path / "j"), quiet=True)
    finally:
        hmc.HamitonianMC._leapfrog = real
    log = g["chain_%s_prop_log" % reg]
    assert [(L, int(a)) for L, a in ch.proposals] == [(int(L), int(a)) for L, a in log[:, :2]]
    xs = np.concatenate([t["x"] for t in traces])
    Us = np.concatenate([t["U"] for t in traces])
    rx, rU = g["chain_%s_steps_x" % reg], g["chain_%s_steps_U" % reg]
    assert np.max(np.abs(xs - rx) / np.max(np.abs(rx), axis=1, keepdims=True)) < 1e-9
    assert np.max(np.abs(Us - rU) / np.abs(rU)) < 1e-9
    mis = np.loadtxt(tmp_path / "j0" / "misfit.dat", ndmin=2)
    assert np.allclose(mis, g["chain_%s_misfit" % reg], rtol=0, atol=2e-8)
    ch.close()
    # the same chain as rank 0 of a batch (DMMA contractions, no mean removal in the batched misfit)
    bt = batched.HMCSampleBatch(model, 2, 8, 0, 0.02, [3, 8], np.ones(M) * 0.001, np.ones(M) * 0.001, b,
                                "mandatory", 1000, dobs, "Fixed", 0.8, alpha, reg, 0.001, 21, 0.05,
                                save_folder=str(tmp_path / "jb"), quiet=True, max_proposals=len(log))
    assert [(L, int(a)) for L, a in bt.proposals[0]][: len(log)] == [(int(L), int(a)) for L, a in log[:, :2]]
    mis = np.loadtxt(tmp_path / "jb0" / "misfit.dat", ndmin=2)
    assert np.allclose(mis, g["chain_%s_misfit" % reg], rtol=0, atol=2e-8)
    bt.close()
    with pytest.raises(AttributeError):
        hmc.HMCSample(model, 1, 0, 0.02, [3, 8], np.ones(M) * 0.001, np.ones(M) * 0.001, b, "mandatory", 1000,
                      dobs, "Fixed", 0.8, alpha, "TV", 0.001, 21, 0.05, save_folder=str(tmp_path / "e"), quiet=True)
