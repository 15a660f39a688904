"""Row-sharded (multi-GPU) parity check, run under torchrun on N >= 2 GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank assembles its observation rows, the chains run through the NCCL all-reduce path
(single chain: inversion/sharded.py; batch: inversion/batched.py) and rank 0 compares positions,
potentials and accept decisions with the CPU oracle (1e-9) and the weights with 1e-12.
Also driven by tests/test_gpu_multi.py when >= 2 GPUs are visible."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import faulthandler

    import torch
    import torch.distributed as dist

    # a hang must not outlive the caller's patience: dump every thread's stack and exit
    faulthandler.dump_traceback_later(int(os.environ.get("GI_CHECK_TIMEOUT", "240")), exit=True)

    from gravinv3dhmc_b200.inversion import batched, hmc, potential
    from oracle import oracle_np as onp

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    g = np.load(os.path.join(ROOT, "tests", "golden", "potential_hmc.npz"))
    o, dobs = g["small_obs"], g["small_dobs"]
    model = potential.GravMagModule(dobs, (0, 400, 0, 600, 0, 500), (100, 100, 100),
                                    (o[:, 0], o[:, 1], o[:, 2]), verbose=False,
                                    shard=(rank, world), group=dist.group.WORLD)
    M = model.M
    lo, hi = model.rows
    assert np.allclose(model.Wm.diagonal(), g["small_wm"], rtol=1e-12)
    err = np.max(np.abs(model.Aw.cpu().numpy() - g["small_Aw"][lo:hi])) / np.max(np.abs(g["small_Aw"]))
    assert err < 1e-10, err
    om = onp.OracleModel(g["small_Aw"], g["small_wm"], dobs, tuple(g["small_mshape"]))
    b = np.ones((M, 2))
    b[:, 0], b[:, 1] = -5.0, 5.0
    b_tv = np.ones((M, 2))
    b_tv[:, 0], b_tv[:, 1] = 0.0, 0.3
    tmp = tempfile.mkdtemp()
    # ---- single chain, row-sharded ----
    cases = {"Damping": (1.0, 0.1, 1.0, [4, 9], b),           # frequent rejections
             "TV": (0.05, 0.02, 0.05, [3, 8], b_tv)}             # the golden TV chain's parameters
    for reg, (alpha, delta, Sigma, Lr, bb) in cases.items():
        ch = hmc.HMCSample(model, 4, 0, delta, Lr, np.ones(M) * 0.001, np.ones(M) * 0.001, bb,
                           "mandatory", 1000, dobs, "Fixed", 0.8, alpha, reg, 0.001, 3, Sigma, myrank=0,
                           save_folder=os.path.join(tmp, "s_%s_r%d_" % (reg, rank)), quiet=True,
                           max_proposals=30)
        ref = onp.hmc_sample(om, 4, 0, delta, Lr, np.ones(M) * 0.001, np.ones(M) * 0.001, bb,
                             "mandatory", 1000, alpha, reg, 0.001, 3, Sigma, max_proposals=30)
        assert [(L, bool(a)) for L, a in ch.proposals] == [(L, bool(a)) for L, a in ref["log"]], reg
        assert np.max(np.abs(ch.x_final - ref["x"])) < 1e-9 * np.max(np.abs(ref["x"])), reg
        assert any(a for _, a in ch.proposals), reg
    # ---- batch of chains, row-sharded: device loop with all-reduce hooks, and the host-driven one ----
    nch, nprops = 5, 5
    nacc = nrej = 0
    finals = {}
    for driver in ("device", "device-nccl", "host"):  # peer memory (default), NCCL hooks, host-driven
        bt = batched.HMCBatch(model, nch, 0.02, [3, 8], np.ones(M) * 0.001, np.ones(M) * 0.001, b_tv,
                              "mandatory", 1000, dobs, 0.05, "MS", 0.001, 3, 0.05,
                              save_folder=os.path.join(tmp, "b%s%d_" % (driver, rank)), quiet=True,
                              driver=driver)
        assert (bt._sh is None) == (driver != "host")
        assert bt.exchange == {"device": "peer", "device-nccl": "nccl", "host": "nccl"}[driver]
        traces = []
        for _ in range(nprops):
            tr = {}
            bt.propose(trace=tr)
            traces.append(tr)
        for c in range(nch):
            otr = []
            onp.hmc_sample(om, 10 ** 6, 0, 0.02, [3, 8], np.ones(M) * 0.001, np.ones(M) * 0.001, b_tv,
                           "mandatory", 1000, 0.05, "MS", 0.001, 3, 0.05, myrank=c, max_proposals=nprops,
                           trace=otr)
            assert [(L, bool(a)) for L, a in bt.proposals[c]] == [(t["L"], bool(t["accept"])) for t in otr]
            for k, t in enumerate(otr):
                L = t["L"]
                rx = np.array([x for x, _ in t["steps"]])
                rU = np.array([U for _, U in t["steps"]])
                gx, gU = traces[k]["x"][: L + 1, c], traces[k]["U"][: L + 1, c]
                assert np.max(np.abs(gx - rx) / np.max(np.abs(rx), axis=1, keepdims=True)) < 1e-9
                assert np.max(np.abs(gU - rU) / np.abs(rU)) < 1e-9
                nacc += bool(t["accept"])
                nrej += not t["accept"]
        finals[driver] = bt.x.copy()
        if driver != "host":
            # the replicated state is bitwise identical on every rank
            xs = [torch.zeros(nch, M, dtype=torch.float64, device="cuda") for _ in range(world)]
            dist.all_gather(xs, torch.as_tensor(bt.x).cuda())
            assert all(torch.equal(xs[0], x) for x in xs)
            bt.close()
    assert np.allclose(finals["device"], finals["host"], rtol=1e-9, atol=0)
    assert np.allclose(finals["device-nccl"], finals["host"], rtol=1e-9, atol=0)
    assert nacc > 0 and nrej >= 0  # the commit path ran (all drivers)
    # ---- a batch with REAL rejections (large step: the golden 'reject' chain's parameters), both
    #      device exchanges: the reject / commit-skip branch of the sharded batch path ----
    for driver in ("device", "device-nccl"):
        bt = batched.HMCBatch(model, 4, 0.1, [4, 9], np.ones(M) * 0.001, np.ones(M) * 0.001, b, "mandatory",
                              1000, dobs, 1.0, "Damping", 0.001, 3, 1.0,
                              save_folder=os.path.join(tmp, "rj%s%d_" % (driver, rank)), quiet=True, driver=driver)
        for _ in range(12):
            bt.propose()
        rej = 0
        for c in range(4):
            ref = onp.hmc_sample(om, 10 ** 6, 0, 0.1, [4, 9], np.ones(M) * 0.001, np.ones(M) * 0.001, b,
                                 "mandatory", 1000, 1.0, "Damping", 0.001, 3, 1.0, myrank=c, max_proposals=12)
            assert [(L, bool(a)) for L, a in bt.proposals[c]] == [(L, bool(a)) for L, a in ref["log"]]
            assert np.max(np.abs(bt.x[c] - ref["x"])) < 1e-9 * np.max(np.abs(ref["x"]))
            rej += sum(1 for _, a in ref["log"] if not a)
        assert rej > 0
        nrej += rej
        bt.close()
    # ---- streaming sampler over the row-sharded kernel (device driver), TV on the full grid ----
    for driver in ("device", "device-nccl"):
        bs = batched.HMCBatch(model, 6, 0.02, [3, 8], np.ones(M) * 0.001, np.ones(M) * 0.001, b_tv,
                              "mandatory", 1000, dobs, 0.05, "TV", 0.001, 3, 0.05,
                              save_folder=os.path.join(tmp, "st%s%d_" % (driver, rank)), quiet=True, driver=driver)
        bs.stream(10 ** 6, 0, max_proposals=4, write=(rank == 0))
        for c in range(6):
            ref = onp.hmc_sample(om, 10 ** 6, 0, 0.02, [3, 8], np.ones(M) * 0.001, np.ones(M) * 0.001, b_tv,
                                 "mandatory", 1000, 0.05, "TV", 0.001, 3, 0.05, myrank=c, max_proposals=4)
            assert [(L, bool(a)) for L, a in bs.proposals[c]] == [(L, bool(a)) for L, a in ref["log"]]
            assert np.max(np.abs(bs.x[c] - ref["x"])) < 1e-9 * np.max(np.abs(ref["x"]))
        bs.close()
    # ---- a wider grid (12 x 10 x 8 = 960 voxels -> several 256-column strips per rank): the column
    #      slices of the peer path really are spread over the ranks; lockstep (with rejections) and
    #      streaming, Damping and TV, against the oracle on the oracle-assembled kernel ----
    xs, ys = np.meshgrid(np.linspace(40, 1160, 8), np.linspace(30, 970, 6))
    xo, yo, zo = xs.ravel(), ys.ravel(), np.full(xs.size, -2.0)
    rng = np.random.RandomState(11)
    wmesh = onp.OracleMesh((0, 1200, 0, 1000, 0, 800), (100, 100, 100))
    wtab, _ = wmesh.active_bounds()
    _, Aor = onp.prism_gz(xo, yo, zo, wtab)
    Awo, wmo, _, _ = onp.sensitivity_weighting(Aor)
    rho = np.zeros(wmesh.shape)
    rho[2:5, 3:7, 4:9] = 0.2
    wd = Aor @ rho.ravel() + 0.01 * rng.randn(xo.size)
    wmodel = potential.GravMagModule(wd, (0, 1200, 0, 1000, 0, 800), (100, 100, 100), (xo, yo, zo),
                                     verbose=False, shard=(rank, world), group=dist.group.WORLD)
    wom = onp.OracleModel(Awo, wmo, wd, wmesh.shape)
    Mw = wmodel.M
    assert Mw == 960
    bw = np.zeros((Mw, 2))
    bw[:, 1] = 0.3
    for reg, delta, Sigma, alpha in (("TV", 0.02, 0.05, 0.05), ("Damping", 0.25, 1.0, 1.0)):
        bb = bw if reg == "TV" else np.c_[np.full(Mw, -5.0), np.full(Mw, 5.0)]
        for mode in ("lockstep", "stream"):
            bt = batched.HMCBatch(wmodel, 7, delta, [3, 8], np.ones(Mw) * 0.001, np.ones(Mw) * 0.001, bb,
                                  "mandatory", 1000, wd, alpha, reg, 0.001, 3, Sigma,
                                  save_folder=os.path.join(tmp, "w%s%s%d_" % (reg, mode, rank)), quiet=True)
            assert bt.exchange == "peer"
            if mode == "lockstep":
                for _ in range(8):
                    bt.propose()
            else:
                bt.stream(10 ** 6, 0, max_proposals=8, write=False)
            for c in range(7):
                ref = onp.hmc_sample(wom, 10 ** 6, 0, delta, [3, 8], np.ones(Mw) * 0.001, np.ones(Mw) * 0.001,
                                     bb, "mandatory", 1000, alpha, reg, 0.001, 3, Sigma, myrank=c,
                                     max_proposals=8)
                assert [(L, bool(a)) for L, a in bt.proposals[c]] == [(L, bool(a)) for L, a in ref["log"]], \
                    (reg, mode, c)
                assert np.max(np.abs(bt.x[c] - ref["x"])) < 1e-9 * np.max(np.abs(ref["x"])), (reg, mode, c)
                if reg == "Damping":
                    nrej += sum(1 for _, a in ref["log"] if not a)
            xs_all = [torch.zeros(7, Mw, dtype=torch.float64, device="cuda") for _ in range(world)]
            dist.all_gather(xs_all, torch.as_tensor(bt.x).cuda())
            assert all(torch.equal(xs_all[0], x) for x in xs_all)  # bitwise identical replicas
            bt.close()
    # ---- row-sharded regularised CG and bootstrap (gi_cg_set_shard) against the oracle ----
    from gravinv3dhmc_b200.inversion import reginv

    gr = np.load(os.path.join(ROOT, "tests", "golden", "reginv.npz"))
    ob = gr["obs"]
    cg = reginv.ConjugateGradient(gr["dobs"], (0, 800, 0, 600, 0, 400), (100, 100, 100),
                                  (ob[:, 0].copy(), ob[:, 1].copy(), ob[:, 2].copy()), verbose=False,
                                  shard=(rank, world), group=dist.group.WORLD)
    rel = lambda a, b: np.max(np.abs(np.asarray(a) - b)) / np.max(np.abs(b))
    for reg in ("MS", "TV"):
        m, d, dm, mm, rf = cg.CG(gr["initial"], gr["aprior"], gr["boundary"], regularization=reg,
                                 beta=float(gr["cg_%s_beta" % reg]), q=0.9, maxk=14)
        assert rel(rf, gr["cg_%s_regul" % reg]) < 1e-9 and rel(dm, gr["cg_%s_data_misfit" % reg]) < 1e-9
        assert rel(m, gr["cg_%s_model" % reg]) < 1e-9 and rel(d, gr["cg_%s_data" % reg]) < 1e-9
    bsr = reginv.BootStrap((0, 800, 0, 600, 0, 400), (100, 100, 100),
                           (ob[:, 0].copy(), ob[:, 1].copy(), ob[:, 2].copy()), gr["dobs"],
                           tuple(gr["boundary"]), samples=5, beta=float(gr["bs_beta"]), maxk=9,
                           verbose=False, shard=(rank, world), group=dist.group.WORLD)
    mi, dmi, mmi, rfi = bsr.BSCG(gr["initial"])
    assert rel(rfi, gr["bs_regul"]) < 1e-9 and rel(dmi, gr["bs_data_misfit"]) < 1e-9
    assert rel(mi, gr["bs_models"]) < 1e-9
    dist.barrier()
    if rank == 0:
        print("multi_gpu_check ok: world=%d, %d accepted / %d rejected batch proposals match the oracle"
              % (world, nacc, nrej))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
