#!/usr/bin/env python
"""Measure the wavelet-compressed forward path (gravmag/compressor3D.py:47-68 per
evaluation = one level-2 db4 DWT of the model + one CSR matvec `Awcp @ coef`) on a B200:

    python tools/wavelet_bench.py [--workload mid|c2|tiny] [--reps 50]      # one JSON line per workload

For each workload: the compression itself (kernelcompressor: batched DWT of every kernel row,
threshold 1e-3, CSR packing), then per evaluation the two kernels timed alone with CUDA events --
`gi_dwt_db4_l2_3d` and `gi_csr_spmv` -- against their ALGORITHMIC bytes:
    DWT : 8 M read + 8 Ncoef written                               (every voxel once, every coefficient once)
    SpMV: nnz (8 B value + 4 B column) + 8 (N + 1) row pointers + 8 Ncoef (x, gathered) + 8 N (y)
and the dense pass it replaces (`gi_gemv_fwd`, 8 N M bytes).  Fractions are of the measured HBM copy
bandwidth (MEASURED_PEAKS.json).  PARITY UNPINNED upstream (PyWavelets absent): this measures the
kernels, the conventions are checked against the oracle restatement in tests/test_gpu_wavelet.py."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (workload geometry, peaks)


def c2_geometry():
    """example/segmentgrid's shape: 10 x 30 x 20 voxels (segmented z), 600 observations on a regular grid"""
    xs, ys = np.meshgrid(np.linspace(0, 2000, 20), np.linspace(0, 3000, 30))
    return ((0, 2000, 0, 3000, 0, 2100), ([100, 200, 300], 100, 100),
            (xs.ravel(), ys.ravel(), np.zeros(xs.size)),
            dict(mseg=True, mdivisionsection=[0, 300, 900, 2100]))


def run(workload, reps):
    import torch

    from gravinv3dhmc_b200 import _lib
    from gravinv3dhmc_b200.gravmag import compressor3D as cp3D
    from gravinv3dhmc_b200.inversion import potential

    L = _lib.lib()
    if workload == "c2":
        mrange, mspacing, obs, kw = c2_geometry()
    else:
        mrange, mspacing, obs, _ = bench.workload_geometry(workload)
        kw = {}
    N = obs[0].size
    ev = lambda: torch.cuda.Event(enable_timing=True)
    model = potential.GravMagModule(np.zeros(N), mrange, mspacing, obs, coordinate="cartesian",
                                    verbose=False, **kw)
    M, mshape = model.M, tuple(int(v) for v in model.mshape)
    a, b = ev(), ev()
    a.record()
    Awcp = cp3D.kernelcompressor(model.Aw, mshape)
    b.record()
    torch.cuda.synchronize()
    t_comp = a.elapsed_time(b) * 1e-3
    ncoef, nnz = Awcp.shape[1], Awcp.nnz
    s = _lib.stream_ptr()
    rng = np.random.RandomState(1)
    m = torch.as_tensor(rng.uniform(0, 0.3, M)).cuda()
    mp = torch.zeros(model.ld, dtype=torch.float64, device="cuda")
    mp[:M] = m
    coef = torch.empty(ncoef, dtype=torch.float64, device="cuda")
    y = torch.empty(N, dtype=torch.float64, device="cuda")

    def timeit(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / reps

    t_dwt = timeit(lambda: _lib.check(L.gi_dwt_db4_l2_3d(_lib.ptr(m), *mshape, _lib.ptr(coef), None, s)))
    t_spmv = timeit(lambda: _lib.check(L.gi_csr_spmv(_lib.ptr(Awcp.indptr), _lib.ptr(Awcp.indices),
                                                     _lib.ptr(Awcp.data), N, _lib.ptr(coef), _lib.ptr(y), s)))
    eng = model.engine()
    t_dense = timeit(lambda: _lib.check(L.gi_gemv_fwd(eng.plan, _lib.ptr(eng.Aw), _lib.ptr(mp), _lib.ptr(eng.d), s)))
    # consistency: compressed forward vs dense (threshold 1e-3 on unit-norm columns; exact only when the
    # lengths nest -- see tests/test_gpu_wavelet.py)
    dd = eng.d.clone()
    rel = float((y - dd).abs().max() / dd.abs().max())
    hbm, src = bench.peaks()
    b_dwt = 8.0 * M + 8.0 * ncoef
    b_spmv = 12.0 * nnz + 8.0 * (N + 1) + 8.0 * ncoef + 8.0 * N
    b_dense = 8.0 * N * M
    return {
        "workload": workload, "grid": mshape, "voxels": M, "observations": N, "coefficients": ncoef,
        "nnz": nnz, "density": nnz / (N * ncoef), "compress_seconds": t_comp,
        "compress_GBps": (8.0 * N * M + 8.0 * N * ncoef) / t_comp / 1e9,
        "dwt": {"us": t_dwt * 1e6, "algorithmic_bytes": b_dwt, "GBps": b_dwt / t_dwt / 1e9,
                "frac_of_hbm_peak": b_dwt / t_dwt / 1e9 / hbm},
        "spmv": {"us": t_spmv * 1e6, "algorithmic_bytes": b_spmv, "GBps": b_spmv / t_spmv / 1e9,
                 "frac_of_hbm_peak": b_spmv / t_spmv / 1e9 / hbm},
        "dense_gemv_fwd": {"us": t_dense * 1e6, "algorithmic_bytes": b_dense, "GBps": b_dense / t_dense / 1e9,
                           "frac_of_hbm_peak": b_dense / t_dense / 1e9 / hbm},
        "per_evaluation_speedup_vs_dense": t_dense / (t_dwt + t_spmv),
        "compressed_vs_dense_forward_rel_diff": rel, "hbm_peak_GBps": hbm, "peak_source": src,
        "parity": "unpinned upstream (PyWavelets absent); kernels == oracle restatement 1e-12",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mid,c2")
    ap.add_argument("--reps", type=int, default=50)
    a = ap.parse_args()
    for w in a.workload.split(","):
        print(json.dumps(run(w, a.reps)))


if __name__ == "__main__":
    main()
