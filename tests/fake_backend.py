"""TEST DOUBLE of libgravinv_b200's building-block entry points on CPU tensors (numpy arithmetic),
so that the HOST logic of the row-sharded path -- row splitting, which quantities are all-reduced,
global means, replicated Metropolis decisions -- can run under `gloo` with world_size 2 on a box
without a GPU.  It is never importable from the product package; the CUDA kernels themselves are
covered by the `-m gpu` parity tests."""
import ctypes as C

import numpy as np

from oracle import oracle_np as onp


def _arr(ptr, n, dtype=np.float64):
    if ptr is None or (hasattr(ptr, "value") and ptr.value is None):
        return None
    addr = ptr.value if hasattr(ptr, "value") else int(ptr)
    ct = {np.float64: C.c_double, np.int32: C.c_int32}[dtype]
    return np.ctypeslib.as_array((ct * int(n)).from_address(addr))


class FakeLib:
    def __init__(self):
        self.plans = {}
        self.err = b""

    # ---- plumbing -----------------------------------------------------------------------------
    def gi_last_error(self):
        return self.err

    def gi_plan_create(self, nrows, M, ld, nchains, out):
        hid = len(self.plans) + 1
        cp = 1 if nchains == 1 else (int(nchains) + 7) // 8 * 8
        self.plans[hid] = dict(n=int(nrows), M=int(M), ld=int(ld), C=int(nchains), Cp=cp,
                               npad=(int(nrows) + 15) // 16 * 16)
        out._obj.value = hid
        return 0

    def gi_plan_destroy(self, plan):
        return 0

    def gi_plan_batch_info(self, plan, cp, npad):
        p = self.plans[plan.value]
        cp._obj.value, npad._obj.value = p["Cp"], p["npad"]
        return 0

    # ---- weighting ------------------------------------------------------------------------------
    def gi_colsumsq(self, G, nrows, M, ld, out, accumulate, stream):
        A = _arr(G, nrows * ld).reshape(nrows, ld)
        o = _arr(out, ld)
        s = o.copy() if accumulate else np.zeros(ld)
        for r in range(nrows):
            s += A[r] * A[r]
        o[:] = s
        return 0

    def gi_weights_from_sumsq(self, sumsq, M, wf, wm, wminv, wmsq, stream):
        s = _arr(sumsq, M)
        d = np.sqrt(s) if wf == 0.5 else np.power(s, wf)
        _arr(wm, M)[:], _arr(wminv, M)[:], _arr(wmsq, M)[:] = d, 1.0 / d, d * d
        return 0

    def gi_scale_columns(self, G, nrows, M, ld, cs, stream):
        A = _arr(G, nrows * ld).reshape(nrows, ld)
        A[:, :M] *= _arr(cs, M)[None, :]
        return 0

    # ---- single chain ---------------------------------------------------------------------------
    def gi_gemv_fwd(self, plan, G, x, d, stream):
        p = self.plans[plan.value]
        A = _arr(G, p["n"] * p["ld"]).reshape(p["n"], p["ld"])
        _arr(d, p["n"])[:] = A @ _arr(x, p["ld"])
        return 0

    def gi_data_sum(self, plan, d, fix, sums, stream):
        p = self.plans[plan.value]
        dv, f = _arr(d, p["n"]), _arr(fix, p["n"])
        _arr(sums, 8)[0] = np.sum(dv + f) if f is not None else np.sum(dv)
        return 0

    def gi_residual(self, plan, d, fix, dobs_c, n_total, r, sums, stream):
        p = self.plans[plan.value]
        dv, f, s = _arr(d, p["n"]), _arr(fix, p["n"]), _arr(sums, 8)
        dinv = dv + f if f is not None else dv
        rr = (dinv - s[0] / n_total) - _arr(dobs_c, p["n"])
        _arr(r, p["n"])[:] = rr
        s[1] = np.sum(rr * rr)
        return 0

    def gi_gemv_adj(self, plan, G, r, g, stream):
        p = self.plans[plan.value]
        A = _arr(G, p["n"] * p["ld"]).reshape(p["n"], p["ld"])
        _arr(g, p["ld"])[:] = A.T @ _arr(r, p["n"])
        return 0

    @staticmethod
    def _reg_grad(reg, mw, apr, wmsq, M):
        r = reg._obj if hasattr(reg, "_obj") else reg
        om = onp.OracleModel.__new__(onp.OracleModel)
        om.wmsq, om.mshape, om._R3d = wmsq, (r.nz, r.ny, r.nx), None
        kind = r.reg_kind
        if kind == 0:
            return om.model_Damping_all(mw, apr)
        if kind == 1:
            return om.model_MS_all(mw, apr, r.beta)
        if kind == 2:
            return om.model_Smoothness_all(mw, apr)
        return om.model_TV_all(mw, apr, r.beta)

    def _update_one(self, reg, grad_in, gdata, x_in, mw_in, apr, wmsq, low, high, pm, x_out, mw_out,
                    grad_out, pcoef, dt, advance, sums, M, copy_x=False, save_k0=False):
        r = reg._obj if hasattr(reg, "_obj") else reg
        if grad_in is not None:
            grad = grad_in[:M].copy()
        else:
            um, gm = self._reg_grad(reg, mw_in[:M], apr[:M], None if wmsq is None else wmsq[:M], M)
            grad = 2.0 * gdata[:M] + r.alpha * gm
            sums[2] = um
        if grad_out is not None:
            grad_out[:M] = grad
        p = pm[:M]
        sums[4] = 0.5 * np.dot(p, p)
        if save_k0:
            sums[5] = sums[4]
        p -= pcoef * grad
        if advance:
            x = x_in[:M] + dt * p
            if r.constraint == 0:
                hi, lo = x > high[:M], x < low[:M]
                x[hi], x[lo] = high[:M][hi], low[:M][lo]
                p[hi | lo] = -p[hi | lo]
                mw = x
            else:
                e = np.e ** (r.log_factor * x)
                mw = (low[:M] + high[:M] * e) / (1 + e)
            x_out[:M] = x
            if mw_out is not x_out:
                mw_out[:M] = mw
        elif copy_x:
            x_out[:M] = x_in[:M]
            if mw_out is not x_out:
                mw_out[:M] = mw_in[:M]
        sums[3] = 0.5 * np.dot(p, p)

    def gi_update(self, plan, reg, gdata, x_in, mw_in, mwapr, wmsq, low, high, pm, x_out, mw_out,
                  grad_out, pcoef, dt, advance, sums, stream):
        p = self.plans[plan.value]
        ld, M = p["ld"], p["M"]
        a = lambda q: _arr(q, ld)
        xo = a(x_out)
        mo = xo if (mw_out is not None and x_out is not None and mw_out.value == x_out.value) else a(mw_out)
        self._update_one(reg, None, a(gdata), a(x_in), a(mw_in), a(mwapr), a(wmsq), a(low), a(high),
                         a(pm), xo, mo, a(grad_out), pcoef, dt, advance, _arr(sums, 8), M)
        return 0

    # ---- batched --------------------------------------------------------------------------------
    def gi_gemm_fwd(self, plan, G, X, D, stream):
        p = self.plans[plan.value]
        A = _arr(G, p["n"] * p["ld"]).reshape(p["n"], p["ld"])
        _arr(D, p["Cp"] * p["n"]).reshape(p["Cp"], p["n"])[:] = \
            _arr(X, p["Cp"] * p["ld"]).reshape(p["Cp"], p["ld"]) @ A.T
        return 0

    def gi_data_sum_batched(self, plan, D, fix, sums, stream):
        p = self.plans[plan.value]
        Dv, f = _arr(D, p["Cp"] * p["n"]).reshape(p["Cp"], p["n"]), _arr(fix, p["n"])
        _arr(sums, 8 * p["Cp"]).reshape(p["Cp"], 8)[:, 0] = (Dv + f[None] if f is not None else Dv).sum(1)
        return 0

    def gi_residual_batched(self, plan, D, fix, dobs_c, n_total, R, sums, stream):
        p = self.plans[plan.value]
        Dv, f = _arr(D, p["Cp"] * p["n"]).reshape(p["Cp"], p["n"]), _arr(fix, p["n"])
        s = _arr(sums, 8 * p["Cp"]).reshape(p["Cp"], 8)
        dinv = Dv + f[None] if f is not None else Dv
        rr = (dinv - s[:, :1] / n_total) - _arr(dobs_c, p["n"])[None]
        Rv = _arr(R, p["Cp"] * p["npad"]).reshape(p["Cp"], p["npad"])
        Rv[:] = 0
        Rv[:, : p["n"]] = rr
        s[:, 1] = (rr * rr).sum(1)
        return 0

    def gi_gemm_adj(self, plan, G, R, Gt, stream):
        p = self.plans[plan.value]
        A = _arr(G, p["n"] * p["ld"]).reshape(p["n"], p["ld"])
        Rv = _arr(R, p["Cp"] * p["npad"]).reshape(p["Cp"], p["npad"])[:, : p["n"]]
        _arr(Gt, p["Cp"] * p["ld"]).reshape(p["Cp"], p["ld"])[:] = Rv @ A
        return 0

    def gi_update_batched(self, plan, reg, grad_in, gdata, x_in, mw_in, mwapr, wmsq, low, high, pm,
                          x_out, mw_out, grad_out, dt, L_dev, step, mode, sums, stream):
        p = self.plans[plan.value]
        ld, M, Cp = p["ld"], p["M"], p["Cp"]
        mat = lambda q: None if _arr(q, 1) is None else _arr(q, Cp * ld).reshape(Cp, ld)
        vec = lambda q: _arr(q, ld)
        gi, gd, xi, mwi, pmv, xo, go = (mat(q) for q in (grad_in, gdata, x_in, mw_in, pm, x_out, grad_out))
        same = mw_out.value == x_out.value
        mo = xo if same else mat(mw_out)
        Ls = _arr(L_dev, Cp, np.int32)
        S = _arr(sums, 8 * Cp).reshape(Cp, 8)
        for c in range(Cp):
            m = mode
            if Ls is not None:
                m = 2 if Ls[c] == 0 else 3 if step == 0 else (0 if step < Ls[c] else 1 if step == Ls[c] else 2)
            pcoef, adv = {0: (dt, 1), 1: (0.5 * dt, 0), 2: (0.0, 0), 3: (0.5 * dt, 1)}[m]
            self._update_one(reg, None if gi is None else gi[c], None if gd is None else gd[c], xi[c],
                             mwi[c], vec(mwapr), vec(wmsq), vec(low), vec(high), pmv[c], xo[c],
                             xo[c] if same else mo[c], go[c] if (go is not None and m == 1) else None,
                             pcoef, dt, adv, S[c], M, copy_x=True, save_k0=(m == 3))
        return 0


def install(monkeypatch_setattr):
    """route gravinv3dhmc_b200._lib to the fake backend (CPU tensors)"""
    import torch

    from gravinv3dhmc_b200 import _lib

    fake = FakeLib()
    monkeypatch_setattr(_lib, "lib", lambda: fake)
    monkeypatch_setattr(_lib, "require_cuda", lambda: torch)
    monkeypatch_setattr(_lib, "stream_ptr", lambda torch=None: None)
    monkeypatch_setattr(_lib, "sync", lambda: None)
    monkeypatch_setattr(_lib, "check", lambda rc, what="": None if rc == 0 else (_ for _ in ()).throw(
        ValueError(what)))
    return fake
