"""world_size-2 gloo worker for tests/test_sharded_host_logic.py (CPU, fake kernel backend)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from gravinv3dhmc_b200 import _lib
    from gravinv3dhmc_b200.gravmag._common import split_rows
    from gravinv3dhmc_b200.inversion import batched, hmc, potential
    from oracle import oracle_np as onp
    from tests import fake_backend

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MASTER_PORT"],
                            rank=rank, world_size=world)
    fake_backend.install(lambda mod, name, val: setattr(mod, name, val))
    g = np.load(os.path.join(ROOT, "tests", "golden", "potential_hmc.npz"))
    A_w, wm_ref, dobs = g["small_Aw"], g["small_wm"], g["small_dobs"]
    A = A_w * wm_ref[None, :]                      # the unweighted kernel
    N, M = A.shape
    lo, hi = split_rows(N, world)[rank]
    ld = _lib.padded_ld(M)
    # a model object over this rank's rows (assembly itself is a CUDA kernel: GPU tests)
    model = potential.GravMagModule.__new__(potential.GravMagModule)
    model.verbose, model.group, model.rank, model.world = False, dist.group.WORLD, rank, world
    model.rows, model.n_total, model.M, model.ld = (lo, hi), N, M, ld
    model.dobs, model.fixed, model.grav_fix, model.wavelet = dobs, False, [], False
    model.weightfactor, model.mshape = 0.5, tuple(int(v) for v in g["small_mshape"])
    model._engine, model.timing = None, {}
    Apad = torch.zeros((hi - lo, ld), dtype=torch.float64)
    Apad[:, :M] = torch.as_tensor(A[lo:hi])
    model.Aw_pad = Apad
    model.sensitivityWeighting()                   # colsumsq -> all_reduce -> weights -> scale
    assert np.allclose(model.Wm.diagonal(), wm_ref, rtol=1e-13), "weights"
    assert np.allclose(model.Aw.numpy(), A_w[lo:hi], rtol=1e-12, atol=1e-15), "Aw shard"
    om = onp.OracleModel(A_w, wm_ref, dobs, model.mshape)
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = -5.0, 5.0
    b_tv = np.ones((M, 2))
    b_tv[:, 0], b_tv[:, 1] = 0.0, 0.3
    out = os.environ["GI_TEST_OUT"]
    cases = {"Damping": (1.0, 0.1, 1.0, [4, 9], b),           # frequent rejections
             "TV": (0.05, 0.02, 0.05, [3, 8], b_tv)}             # the golden TV chain's parameters
    for reg, (alpha, delta, Sigma, Lr, bb) in cases.items():
        ch = hmc.HMCSample(model, 3, 0, delta, Lr, np.ones(M) * 0.001, np.ones(M) * 0.001, bb,
                           "mandatory", 1000, dobs, "Fixed", 0.8, alpha, reg, 0.001, 3, Sigma, myrank=0,
                           save_folder=os.path.join(out, "s_%s_r%d_" % (reg, rank)), quiet=True,
                           max_proposals=30)
        ref = onp.hmc_sample(om, 3, 0, delta, Lr, np.ones(M) * 0.001, np.ones(M) * 0.001, bb,
                             "mandatory", 1000, alpha, reg, 0.001, 3, Sigma, max_proposals=30)
        assert [(L, bool(a)) for L, a in ch.proposals] == [(L, bool(a)) for L, a in ref["log"]], reg
        assert np.max(np.abs(ch.x_final - ref["x"])) < 1e-9 * np.max(np.abs(ref["x"])), reg
        assert any(a for _, a in ch.proposals), reg
        # only rank 0 writes the chain files
        assert os.path.exists(os.path.join(out, "s_%s_r%d_0" % (reg, rank), "misfit.dat")) == (rank == 0)
    nch, nprops = 3, 4
    bt = batched.HMCBatch(model, nch, 0.1, [4, 9], np.ones(M) * 0.001, np.ones(M) * 0.001, b,
                          "mandatory", 1000, dobs, 1.0, "MS", 0.001, 3, 1.0,
                          save_folder=os.path.join(out, "b_r%d_" % rank), quiet=True)
    for _ in range(nprops):
        bt.propose()
    for c in range(nch):
        ref = onp.hmc_sample(om, 10 ** 6, 0, 0.1, [4, 9], np.ones(M) * 0.001, np.ones(M) * 0.001, b,
                             "mandatory", 1000, 1.0, "MS", 0.001, 3, 1.0, myrank=c, max_proposals=nprops)
        assert [(L, bool(a)) for L, a in bt.proposals[c]] == [(L, bool(a)) for L, a in ref["log"]], c
        assert np.max(np.abs(bt.x[c] - ref["x"])) < 1e-9 * np.max(np.abs(ref["x"])), c
    xs = [torch.zeros_like(bt._sh.x_cur) for _ in range(world)]
    dist.all_gather(xs, bt._sh.x_cur)
    assert all(torch.equal(xs[0], x) for x in xs), "replicated state diverged"
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
