#!/usr/bin/env python
"""fused_bench.py -- the single-pass gradient evaluation (gi_fused_pass, csrc/fused.cu) against the
two-pass kernels (gi_gemv_fwd + gi_gemv_adj): results and time per evaluation on random matrices.

    python tools/fused_bench.py [--rows 4096 --cols 1048576 --reps 5]
"""
import argparse
import ctypes as C
import json
import os
import sys


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4096)
    ap.add_argument("--cols", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch

    from gravinv3dhmc_b200 import _lib

    L = _lib.lib()
    n, m = args.rows, args.cols
    ld = _lib.padded_ld(m)
    f64 = dict(dtype=torch.float64, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.zeros((n, ld), **f64)
    A[:, :m] = torch.rand((n, m), generator=g, **f64) * 1e-3
    x = torch.zeros(ld, **f64)
    x[:m] = torch.rand(m, generator=g, **f64)
    dobs_c = torch.randn(n, generator=g, **f64)
    fix = torch.randn(n, generator=g, **f64) * 0.1
    s = _lib.stream_ptr()
    plan = C.c_void_p()
    _lib.check(L.gi_plan_create(n, m, ld, 1, C.byref(plan)))
    d2, r2, g2, sums = torch.zeros(n, **f64), torch.zeros(n, **f64), torch.zeros(ld, **f64), torch.zeros(8, **f64)

    def two_pass():
        _lib.check(L.gi_gemv_fwd(plan, _lib.ptr(A), _lib.ptr(x), _lib.ptr(d2), s))
        _lib.check(L.gi_data_sum(plan, _lib.ptr(d2), _lib.ptr(fix), _lib.ptr(sums), s))
        _lib.check(L.gi_residual(plan, _lib.ptr(d2), _lib.ptr(fix), _lib.ptr(dobs_c), n, _lib.ptr(r2), _lib.ptr(sums), s))
        _lib.check(L.gi_gemv_adj(plan, _lib.ptr(A), _lib.ptr(r2), _lib.ptr(g2), s))

    fh = C.c_void_p()
    _lib.check(L.gi_fused_create(n, m, ld, _lib.ptr(A), s, C.byref(fh)), "gi_fused_create")
    d1, g1 = torch.zeros(n, **f64), torch.zeros(ld, **f64)

    def fused():
        _lib.check(L.gi_fused_pass(fh, _lib.ptr(x), _lib.ptr(dobs_c), _lib.ptr(fix), 1, _lib.ptr(d1), _lib.ptr(g1), s))

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.reps

    t2, t1 = timed(two_pass), timed(fused)
    nrm = lambda a, b: float((a - b).abs().max() / b.abs().max())
    gb = 8.0 * n * ld / 1e9
    out = {"rows": n, "cols": m, "two_pass_ms": t2, "fused_ms": t1, "speedup": t2 / t1,
           "two_pass_GBps_algorithmic": 2 * gb / t2 * 1e3, "fused_GBps_algorithmic": 2 * gb / t1 * 1e3,
           "fused_GBps_dram": gb / t1 * 1e3, "err_d": nrm(d1, d2), "err_g": nrm(g1[:m], g2[:m])}
    fused()
    torch.cuda.synchronize()
    d1b, g1b = d1.clone(), g1.clone()
    fused()
    torch.cuda.synchronize()
    out["deterministic"] = bool(torch.equal(d1, d1b) and torch.equal(g1, g1b))
    print(json.dumps(out))
    L.gi_fused_destroy(fh)
    L.gi_plan_destroy(plan)


if __name__ == "__main__":
    main()
