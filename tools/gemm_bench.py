#!/usr/bin/env python
"""Time the batched contraction kernels alone (random matrix, no assembly):
    python tools/gemm_bench.py [--rows 8192 --cols 262144 --chains 64 --reps 10]
Prints TFLOP/s (FP64) of gi_gemm_fwd / gi_gemm_adj and, for chains == 1, GB/s of the GEMV passes."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from gravinv3dhmc_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=8192)
ap.add_argument("--cols", type=int, default=262144)
ap.add_argument("--chains", type=int, default=64)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
L = _lib.lib()
n, m, c = a.rows, a.cols, a.chains
ld = _lib.padded_ld(m)
f64 = dict(dtype=torch.float64, device="cuda")
A = torch.empty((n, ld), **f64).uniform_(-1, 1)
plan = C.c_void_p()
_lib.check(L.gi_plan_create(n, m, ld, c, C.byref(plan)))
s = _lib.stream_ptr()


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / a.reps


if c > 1:
    cp, npad = C.c_int32(), C.c_int64()
    _lib.check(L.gi_plan_batch_info(plan, C.byref(cp), C.byref(npad)))
    X = torch.empty((cp.value, ld), **f64).uniform_(-1, 1)
    R = torch.empty((cp.value, npad.value), **f64).uniform_(-1, 1)
    D = torch.empty((cp.value, n), **f64)
    Gt = torch.empty((cp.value, ld), **f64)
    tf = timeit(lambda: _lib.check(L.gi_gemm_fwd(plan, _lib.ptr(A), _lib.ptr(X), _lib.ptr(D), s)))
    ta = timeit(lambda: _lib.check(L.gi_gemm_adj(plan, _lib.ptr(A), _lib.ptr(R), _lib.ptr(Gt), s)))
    fl = 2.0 * n * m * c
    print("rows=%d cols=%d chains=%d  fwd %.3f ms %.2f TFLOP/s | adj %.3f ms %.2f TFLOP/s"
          % (n, m, c, tf * 1e3, fl / tf / 1e12, ta * 1e3, fl / ta / 1e12))
else:
    x = torch.empty(ld, **f64).uniform_(-1, 1)
    r = torch.empty(n, **f64).uniform_(-1, 1)
    d = torch.empty(n, **f64)
    g = torch.empty(ld, **f64)
    tf = timeit(lambda: _lib.check(L.gi_gemv_fwd(plan, _lib.ptr(A), _lib.ptr(x), _lib.ptr(d), s)))
    ta = timeit(lambda: _lib.check(L.gi_gemv_adj(plan, _lib.ptr(A), _lib.ptr(r), _lib.ptr(g), s)))
    by = 8.0 * n * m
    print("rows=%d cols=%d chains=1  fwd %.3f ms %.0f GB/s | adj %.3f ms %.0f GB/s"
          % (n, m, tf * 1e3, by / tf / 1e9, ta * 1e3, by / ta / 1e9))
L.gi_plan_destroy(plan)
