"""Device-side evaluation engine shared by `potential.GravMagModule` and the samplers.

`BlockEngine` drives the building-block C-ABI entry points (gi_gemv_fwd, gi_data_sum, gi_residual,
gi_gemv_adj, gi_update) for one row shard of the weighted kernel.  With a process group the two
exchange steps of the row-sharded path are `torch.distributed.all_reduce` calls (NCCL on GPUs,
gloo in the CPU tests of the host logic):

    local d = Aw_g x  ->  all_reduce(sum d)  ->  local r  ->  local g = Aw_g^T r
                      ->  all_reduce(g, Ud)  ->  replicated update (bitwise identical on all ranks)

All heavy arithmetic happens in libgravinv_b200.so; torch owns memory, streams and collectives.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib


def reg_params(regularization, constraint, mshape, alpha, beta, log_factor):
    if constraint not in _lib.CONSTRAINTS:  # potential.py:824, hmc.py:278
        raise ValueError("Please choose right boundary constraint(mandatory, logarithmic)!")
    if regularization not in _lib.REG_KINDS:  # potential.py:836
        raise ValueError("Please choose regularization from 'MS','Damping', 'Smoothness', 'TV'.")
    nz, ny, nx = (int(v) for v in mshape)
    return _lib.RegParams(_lib.REG_KINDS[regularization], _lib.CONSTRAINTS[constraint], nz, ny, nx,
                          0, float(alpha), float(beta), float(log_factor))


class BlockEngine:
    """One row shard: Aw_pad is the [n_local, ld] CUDA tensor (padding columns zero)."""

    def __init__(self, Aw_pad, M, dobs_local, dobs_mean, n_total, gravfix_local=None, group=None):
        self.torch = torch = _lib.require_cuda()
        self.L = _lib.lib()
        self.Aw = Aw_pad
        self.dev = Aw_pad.device
        self.n_local, self.ld = (int(v) for v in Aw_pad.shape)
        self.M = int(M)
        self.n_total = int(n_total)
        self.group = group
        self.world = 1
        if group is not None:
            import torch.distributed as dist

            self.world = dist.get_world_size(group)
        f64 = dict(dtype=torch.float64, device=self.dev)
        dl = np.asarray(dobs_local, dtype=np.float64)
        self.dobs_c = torch.as_tensor(dl - dobs_mean, **f64)  # potential.py:706
        self.fix = None
        if gravfix_local is not None:
            self.fix = torch.as_tensor(np.asarray(gravfix_local, dtype=np.float64), **f64)
        self.d = torch.zeros(self.n_local, **f64)
        self.r = torch.zeros(self.n_local, **f64)
        # gradient partial + 8 trailing scalar slots so that one all-reduce carries both
        self.gext = torch.zeros(self.ld + 8, **f64)
        self.g = self.gext[: self.ld]
        self.sums = torch.zeros(8, **f64)
        self.plan = C.c_void_p()
        _lib.check(self.L.gi_plan_create(self.n_local, self.M, self.ld, 1, C.byref(self.plan)),
                   "gi_plan_create")
        self.launches = 0
        self._fused = None  # single-pass evaluation (csrc/fused.cu): None = not tried, False = unavailable

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                self.L.gi_plan_destroy(self.plan)
                self.plan = None
            if getattr(self, "_fused", None):
                self.L.gi_fused_destroy(self._fused)
                self._fused = None
        except Exception:
            pass

    def _fused_handle(self):
        """gi_fused handle over this rank's rows for kernels >= 1 GB (GI_FUSED_GEMV=1 forces it for
        any shape that fits, =0 disables), plus s = Aw_local^T 1 for the row-sharded mean correction"""
        if self._fused is None:
            import os

            self._fused = False
            env = os.environ.get("GI_FUSED_GEMV", "")
            # rows of >= 4 MB in kernels of >= 1 GB: below that the per-row hand-off of the single-pass
            # kernel costs more than the second pass it saves (csrc/leapfrog.cu: fused_ready)
            big = self.n_local * self.ld * 8.0 >= 1e9 and self.M >= (1 << 19)
            if env != "0" and (big or env == "1"):
                h = C.c_void_p()
                if self.L.gi_fused_create(self.n_local, self.M, self.ld, _lib.ptr(self.Aw),
                                          _lib.stream_ptr(), C.byref(h)) == _lib.GI_OK:
                    self._fused = h
                    if self.world > 1:
                        ones = self.torch.ones(self.n_local, dtype=self.torch.float64, device=self.dev)
                        self._s_local = self.torch.zeros(self.ld, dtype=self.torch.float64, device=self.dev)
                        _lib.check(self.L.gi_gemv_adj(self.plan, _lib.ptr(self.Aw), _lib.ptr(ones),
                                                      _lib.ptr(self._s_local), _lib.stream_ptr()), "gi_gemv_adj")
        return self._fused

    def vec(self, a=None):
        """zero-padded device M-vector (ld entries)"""
        t = self.torch.zeros(self.ld, dtype=self.torch.float64, device=self.dev)
        if a is not None:
            t[: self.M] = self.torch.as_tensor(np.asarray(a, dtype=np.float64), device=self.dev)
        return t

    def _all_reduce(self, t):
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def data_pass(self, mw, dpre=None):
        """d, r, Ud and gdata = Aw^T r for the padded device vector `mw`; leaves the results in
        self.d / self.r / self.g / self.sums[0:2].  `dpre` (device, this rank's rows) overrides the
        dense forward product (wavelet-compressed forward, potential.py:693-696)."""
        L, s, p = self.L, _lib.stream_ptr(), _lib.ptr
        fused = self._fused_handle() if dpre is None else False
        if fused:
            # d and Aw_local^T (e - mean_local) in ONE pass over this rank's rows
            _lib.check(L.gi_fused_pass(fused, p(mw), p(self.dobs_c), p(self.fix), 1 if self.n_total > 0 else 0,
                                       p(self.d), p(self.g), s),
                       "gi_fused_pass")
        elif dpre is not None:
            self.d.copy_(dpre)
        else:
            _lib.check(L.gi_gemv_fwd(self.plan, p(self.Aw), p(mw), p(self.d), s), "gi_gemv_fwd")
        _lib.check(L.gi_data_sum(self.plan, p(self.d), p(self.fix), p(self.sums), s), "gi_data_sum")
        if fused and self.world > 1:
            s0_local = self.sums[0].clone()
        self._all_reduce(self.sums[0:1])
        _lib.check(L.gi_residual(self.plan, p(self.d), p(self.fix), p(self.dobs_c), self.n_total,
                                 p(self.r), p(self.sums), s), "gi_residual")
        if fused:
            if self.world > 1:
                # the kernel centred e on the LOCAL mean: move to the global one,
                # Aw_l^T (e - mean_g) = Aw_l^T (e - mean_l) + (mean_l - mean_g) * Aw_l^T 1
                delta = s0_local / self.n_local - self.sums[0] / self.n_total
                self.g.addcmul_(self._s_local, delta.expand_as(self._s_local))
        else:
            _lib.check(L.gi_gemv_adj(self.plan, p(self.Aw), p(self.r), p(self.g), s), "gi_gemv_adj")
        if self.world > 1:
            # one exchange: the gradient partials and the partial sum r^2 ride together
            self.gext[self.ld] = self.sums[1]
            self._all_reduce(self.gext)
            self.sums[1] = self.gext[self.ld]
        self.launches += 6

    def update(self, reg, x_in, mw_in, mwapr, wmsq, low, high, pm, x_out, mw_out, grad_out, pcoef,
               dt, advance):
        p = _lib.ptr
        _lib.check(self.L.gi_update(self.plan, C.byref(reg), p(self.g), p(x_in), p(mw_in), p(mwapr),
                                    p(wmsq), p(low), p(high), p(pm), p(x_out), p(mw_out),
                                    p(grad_out), float(pcoef), float(dt), int(advance),
                                    p(self.sums), _lib.stream_ptr()), "gi_update")
        self.launches += 1

    def to_mw(self, x, low, high, constraint, log_factor):
        """potential.py:819-820 on device (elementwise transform of an M-vector: plumbing)."""
        if constraint == "mandatory":
            return x
        t = self.torch
        e = t.pow(t.tensor(np.e, dtype=t.float64, device=self.dev), log_factor * x)
        mw = (low + high * e) / (1 + e)
        mw[self.M:] = 0
        return mw


class BatchEngine:
    """One row shard, C chains batched as columns (gi_gemm_fwd / gi_data_sum_batched /
    gi_residual_batched / gi_gemm_adj / gi_update_batched).  Per gradient evaluation of the batch:

        D_g = Aw_g X          local rows, all chains        (DMMA contraction)
        all-reduce(sum D)     Cp doubles
        R_g, |R_g|^2          local
        Gt_g = Aw_g^T R_g     local partial [Cp][ld]        (DMMA contraction)
        all-reduce(Gt, |R|^2) Cp*ld + Cp doubles            (one NCCL call)
        update                replicated on every rank (NCCL returns identical bits everywhere)
    """

    def __init__(self, Aw_pad, M, nchains, dobs_local, dobs_mean, n_total, gravfix_local=None,
                 group=None):
        self.torch = torch = _lib.require_cuda()
        self.L = _lib.lib()
        self.Aw = Aw_pad
        self.dev = Aw_pad.device
        self.n_local, self.ld = (int(v) for v in Aw_pad.shape)
        self.M, self.nchains, self.n_total, self.group = int(M), int(nchains), int(n_total), group
        self.world = 1
        if group is not None:
            import torch.distributed as dist

            self.world = dist.get_world_size(group)
        self.plan = C.c_void_p()
        _lib.check(self.L.gi_plan_create(self.n_local, self.M, self.ld, self.nchains,
                                         C.byref(self.plan)), "gi_plan_create")
        cp, npad = C.c_int32(), C.c_int64()
        _lib.check(self.L.gi_plan_batch_info(self.plan, C.byref(cp), C.byref(npad)),
                   "gi_plan_batch_info")
        self.Cp, self.npad = cp.value, npad.value
        f64 = dict(dtype=torch.float64, device=self.dev)
        dl = np.asarray(dobs_local, dtype=np.float64)
        self.dobs_c = torch.as_tensor(dl - dobs_mean, **f64)
        self.fix = None
        if gravfix_local is not None:
            self.fix = torch.as_tensor(np.asarray(gravfix_local, dtype=np.float64), **f64)
        self.d = torch.zeros((self.Cp, self.n_local), **f64)
        self.r = torch.zeros((self.Cp, self.npad), **f64)
        # gradient partials + Cp trailing slots: one all-reduce carries Gt and sum r^2
        self.gext = torch.zeros(self.Cp * self.ld + self.Cp, **f64)
        self.g = self.gext[: self.Cp * self.ld].view(self.Cp, self.ld)
        self.sums = torch.zeros((self.Cp, 8), **f64)
        self.s0 = torch.zeros(self.Cp, **f64)
        self.launches = 0

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                self.L.gi_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    def mat(self, a=None):
        """zero-padded device [Cp][ld] matrix; `a` is [nchains][M] (numpy)"""
        t = self.torch.zeros((self.Cp, self.ld), dtype=self.torch.float64, device=self.dev)
        if a is not None:
            t[: self.nchains, : self.M] = self.torch.as_tensor(
                np.ascontiguousarray(a, dtype=np.float64), device=self.dev)
        return t

    def vec(self, a=None):
        t = self.torch.zeros(self.ld, dtype=self.torch.float64, device=self.dev)
        if a is not None:
            t[: self.M] = self.torch.as_tensor(np.asarray(a, dtype=np.float64), device=self.dev)
        return t

    def _all_reduce(self, t):
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def data_pass(self, MW):
        """D, R, Ud and Gt = Aw^T R for the padded [Cp][ld] device matrix MW"""
        L, s, p = self.L, _lib.stream_ptr(), _lib.ptr
        _lib.check(L.gi_gemm_fwd(self.plan, p(self.Aw), p(MW), p(self.d), s), "gi_gemm_fwd")
        _lib.check(L.gi_data_sum_batched(self.plan, p(self.d), p(self.fix), p(self.sums), s),
                   "gi_data_sum_batched")
        if self.world > 1:
            self.s0.copy_(self.sums[:, 0])
            self._all_reduce(self.s0)
            self.sums[:, 0] = self.s0
        _lib.check(L.gi_residual_batched(self.plan, p(self.d), p(self.fix), p(self.dobs_c),
                                         self.n_total, p(self.r), p(self.sums), s),
                   "gi_residual_batched")
        _lib.check(L.gi_gemm_adj(self.plan, p(self.Aw), p(self.r), p(self.g), s), "gi_gemm_adj")
        if self.world > 1:
            self.gext[self.Cp * self.ld:] = self.sums[:, 1]
            self._all_reduce(self.gext)
            self.sums[:, 1] = self.gext[self.Cp * self.ld:]
        self.launches += 6

    def update(self, reg, grad_in, x_in, mw_in, mwapr, wmsq, low, high, pm, x_out, mw_out, grad_out,
               dt, L_dev, step, mode):
        p = _lib.ptr
        _lib.check(self.L.gi_update_batched(self.plan, C.byref(reg), p(grad_in),
                                            None if grad_in is not None else p(self.g), p(x_in),
                                            p(mw_in), p(mwapr), p(wmsq), p(low), p(high), p(pm),
                                            p(x_out), p(mw_out), p(grad_out), float(dt), p(L_dev),
                                            int(step), int(mode), p(self.sums), _lib.stream_ptr()),
                   "gi_update_batched")
        self.launches += 1
