#!/usr/bin/env python
"""cg_bench.py -- throughput of the regularised CG (gi_cg_*) and its batched bootstrap on the
bench.py workloads, against the roofline that bounds each:

  * ConjugateGradient.CG, 1 column: 3 streaming passes over Aw per iteration (2 forward, 1 adjoint)
    -> HBM bound, algorithmic bytes 3 * 8 * N * M per iteration;
  * BootStrap.BSCG, C replicates as columns: 3 DMMA contractions per iteration -> FP64 tensor pipe,
    3 * 2 * N * M * C flop per iteration.

    python tools/cg_bench.py [--workload mid] [--iters 10] [--replicates 64] [--cpu]

Prints one JSON line per mode.  `--cpu` also times the oracle port of the reference loop
(oracle_np.OracleCG, numpy, all BLAS threads) on a bounded row sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (workload table, peaks)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mid", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--replicates", type=int, default=64)
    ap.add_argument("--reg", default="MS")
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--cpu-rows", type=int, default=128)
    args = ap.parse_args()

    import torch

    from gravinv3dhmc_b200 import _lib
    from gravinv3dhmc_b200.inversion import reginv

    hbm_peak, peak_src = bench.peaks()
    mrange, msp, obs, rho = bench.workload_geometry(args.workload)
    (nz, ny, nx), _, side = bench.WORKLOADS[args.workload]
    N, M = side * side, nz * ny * nx
    t0 = time.perf_counter()
    cg = reginv.ConjugateGradient(np.zeros(N), mrange, msp, obs, verbose=False)
    mod = cg._mod
    dobs = mod.forward_local(cg.Wm @ rho).cpu().numpy()
    rng = np.random.default_rng(12345)
    dobs = dobs + rng.normal(0.0, 0.02 * np.abs(dobs).max(), dobs.shape)
    cg.dobs = dobs
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    init, apr = np.full(M, 0.001), np.full(M, 0.001)

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        return out, e0.elapsed_time(e1) * 1e-3

    out = []
    # ---- single CG: warm-up run (2 iterations) then the timed run; tolerance 0.001 is never hit
    # on this noisy workload, so exactly `iters` iterations execute
    cg.CG(init, apr, (0.0, 1.0), regularization=args.reg, beta=0.001, q=0.9, maxk=2)
    res, sec = timed(lambda: cg.CG(init, apr, (0.0, 1.0), regularization=args.reg, beta=0.001, q=0.9,
                                   maxk=args.iters))
    n_it = len(res[4])
    # the same run split into its parts: handle creation, the device loop, result read-back
    mw0, mwapr = cg.Wm @ init, cg.Wm @ apr
    t0 = time.perf_counter()
    h = reginv._CgHandle(mod.Aw_pad, M, cg.dobs, mod.wm_dev, mod.wminv_dev, mod.wmsq_dev, _lib.CG_REGINV,
                         reginv._reg(args.reg, cg.mshape, 0.001), 0.9, 0.001, (0.0, 1.0), ncols=1, mwapr=mwapr)
    torch.cuda.synchronize()
    t_create = time.perf_counter() - t0
    _, t_loop = timed(lambda: h.run(mw0, args.iters))
    t0 = time.perf_counter()
    h.result()
    t_result = time.perf_counter() - t0
    h.close()
    loop_gbs = (3 * args.iters + 1) * 8.0 * N * M / t_loop / 1e9
    passes = 3 * n_it + 2  # + start-point forward and the final data_inv forward
    gbs = passes * 8.0 * N * M / sec / 1e9
    out.append({"mode": "cg", "workload": args.workload, "voxels": M, "observations": N, "reg": args.reg,
                "iterations": n_it, "seconds": sec, "iters_per_s": n_it / sec,
                "passes_over_Aw": passes, "hbm_gbs": gbs, "hbm_peak_gbs": hbm_peak, "frac": gbs / hbm_peak,
                "peak_src": peak_src, "setup_s": setup_s, "launches": cg.last_launches,
                "data_misfit_first_last": [res[2][0], res[2][-1]],
                "parts_s": {"create": t_create, "loop": t_loop, "result": t_result},
                "loop_hbm_gbs": loop_gbs, "loop_frac": loop_gbs / hbm_peak})
    print(json.dumps(out[-1]), flush=True)

    # ---- bootstrap: C replicates per pass
    C_ = args.replicates
    bs = reginv.BootStrap.__new__(reginv.BootStrap)  # share the assembled kernel with `cg`
    bs.__dict__.update(cg.__dict__)
    bs.boundary, bs.samples, bs.maxk, bs.beta, bs.batch = (0.0, 1.0), C_, 3, 0.05, 64
    bs.BSCG(init)
    bs.maxk = args.iters
    res, sec = timed(lambda: bs.BSCG(init))
    nb = (C_ + 63) // 64
    flops = (3 * args.iters + 1) * 2.0 * N * M * C_
    out.append({"mode": "bootstrap", "workload": args.workload, "replicates": C_, "iterations": args.iters,
                "seconds": sec, "replicate_iters_per_s": C_ * args.iters / sec, "batches": nb,
                "fp64_tflops": flops / sec / 1e12, "fp64_peak_tflops": bench.FP64_PEAK_TFLOPS,
                "frac": flops / sec / 1e12 / bench.FP64_PEAK_TFLOPS, "launches": bs.last_launches,
                "incl": "host RNG resampling, H2D of row multiplicities, D2H of the models"})
    print(json.dumps(out[-1]), flush=True)

    if args.cpu:
        from oracle import oracle_np as onp

        rows = np.linspace(0, N - 1, args.cpu_rows).astype(np.int64)
        mesh = onp.OracleMesh(mrange, msp)
        _, A = onp.prism_gz(obs[0][rows], obs[1][rows], obs[2][rows], mesh.active_bounds()[0],
                            threads=os.cpu_count())
        ocg = onp.OracleCG(A, dobs[rows], (nz, ny, nx))
        k = 3
        t0 = time.perf_counter()
        ocg.CG(init, apr, (0.0, 1.0), args.reg, 0.001, 0.9, k)
        sec = time.perf_counter() - t0
        out.append({"mode": "cg_cpu_port", "rows": int(rows.size), "iterations": k, "seconds": sec,
                    "iters_per_s_scaled_to_N": k / sec * rows.size / N, "cores": os.cpu_count(),
                    "kind": "port (oracle_np.OracleCG, numpy BLAS threads)"})
        print(json.dumps(out[-1]), flush=True)


if __name__ == "__main__":
    main()
