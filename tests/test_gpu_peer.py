"""The peer-memory exchange of row-sharded batches (csrc/peer.cu, gi_hmcb_set_peer) on ONE GPU: the
rank is its own peer, so the slot-table scalar reduction, the staged gradient partials, the
column-slice update and the epoch bookkeeping all run (only the NVLink copies are empty).  The
multi-rank behaviour is checked by tests/multi_gpu_check.py (`gpurun --gpus 2/4`)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from gravinv3dhmc_b200 import _lib  # noqa: E402
from gravinv3dhmc_b200.inversion import batched, peer, potential  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402


def test_allreduce_small_single_rank():
    pb = peer.PeerBuffer(1 << 20)
    t = torch.arange(64, dtype=torch.float64, device="cuda")
    for _ in range(5):  # the slot table alternates between its two halves
        pb.allreduce_small(t)
    assert torch.equal(t.cpu(), torch.arange(64, dtype=torch.float64))
    pb.close()


@pytest.mark.parametrize("reg,constraint", [("TV", "mandatory"), ("MS", "mandatory"), ("Damping", "logarithmic")])
def test_peer_path_matches_oracle_single_rank(golden, reg, constraint, tmp_path):
    g = golden["potential_hmc"]
    o, dobs = g["small_obs"], g["small_dobs"]
    model = potential.GravMagModule(dobs, (0, 400, 0, 600, 0, 500), (100, 100, 100),
                                    (o[:, 0], o[:, 1], o[:, 2]), verbose=False)
    om = onp.OracleModel(g["small_Aw"], g["small_wm"], dobs, tuple(g["small_mshape"]))
    M = model.M
    b = np.ones((M, 2))
    if constraint == "logarithmic":
        b[:, 0], b[:, 1] = -0.5, 1.5
        args = (1e-5, [3, 6], np.ones(M) * 0.3, np.ones(M) * 0.3, b, constraint, 1000)
        alpha, Sigma = 1.0, 1e-4
    else:
        b[:, 0], b[:, 1] = 0.0, 0.3
        args = (0.02, [3, 8], np.ones(M) * 0.001, np.ones(M) * 0.001, b, constraint, 1000)
        alpha, Sigma = 0.05, 0.05
    for mode in ("lockstep", "stream"):
        bt = batched.HMCBatch(model, 5, *args, dobs, alpha, reg, 0.001, 3, Sigma,
                              save_folder=str(tmp_path / (reg + mode)), quiet=True, driver="device-peer")
        assert bt.exchange == "peer"
        if mode == "lockstep":
            for _ in range(6):
                bt.propose()
        else:
            bt.stream(10 ** 6, 0, max_proposals=6, write=False)
        for c in range(5):
            ref = onp.hmc_sample(om, 10 ** 6, 0, *args, alpha, reg, 0.001, 3, Sigma, myrank=c, max_proposals=6)
            assert [(L, bool(a)) for L, a in bt.proposals[c]] == [(L, bool(a)) for L, a in ref["log"]]
            assert np.max(np.abs(bt.x[c] - ref["x"])) < 1e-9 * np.max(np.abs(ref["x"]))
        bt.close()
