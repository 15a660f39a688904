"""Row-sharded (multi-GPU) HMC proposal: G is partitioned by observation rows across ranks
(contiguous chunks as in the reference's worker pool, gravmag/prism.py:986-996); all M-vectors are
replicated.  Per gradient evaluation:

    d_g = Aw_g mw          local rows          (gi_gemv_fwd)
    sum(d) all-reduce      1 double            -> mean removal needs the global sum
    r_g, |r_g|^2           local               (gi_residual)
    g_g = Aw_g^T r_g       local partial       (gi_gemv_adj)
    all-reduce(g, |r|^2)   M + 1 doubles       (one NCCL call)
    update / clamp / K     replicated          (gi_update) -- NCCL returns the same bits on every
                                                rank, so the replicated state stays identical

Every rank draws the same momentum and uniform (same seed), so accept decisions agree without
communication.  The arithmetic of `hmc.py:85-177` is unchanged; see leapfrog.cu.
"""
from __future__ import annotations

import numpy as np


class _ShardState:
    """replicated device vectors of one chain on one rank"""

    def __init__(self, chain, alpha):
        m = chain.model
        eng = m.engine()
        self.eng = eng
        self.low, self.high = eng.vec(chain.low), eng.vec(chain.high)
        self.apr = eng.vec(chain.aprior_model)
        self.x_cur, self.g_cur = eng.vec(), eng.vec()
        self.xa, self.xb, self.p, self.gnew = eng.vec(), eng.vec(), eng.vec(), eng.vec()
        self.logc = chain.constraint == "logarithmic"
        self.mw_cur = eng.vec() if self.logc else self.x_cur
        self.mwa = eng.vec() if self.logc else self.xa
        self.mwb = eng.vec() if self.logc else self.xb
        self.d_cur = eng.torch.zeros_like(eng.d)
        self.U = self.Ud = self.Um = None
        self.synced = None


def _grad_eval(st, chain, reg, x_in, mw_in, x_out, mw_out, grad_out, pcoef, dt, advance):
    eng = st.eng
    eng.data_pass(mw_in, chain.model._forward(mw_in))
    eng.update(reg, x_in, mw_in, st.apr, chain.model.wmsq_dev, st.low, st.high, st.p, x_out, mw_out,
               grad_out, pcoef, dt, advance)


def _set_state(st, chain, reg, xcur):
    eng = st.eng
    st.x_cur.copy_(eng.vec(xcur))
    if st.logc:
        st.mw_cur.copy_(eng.to_mw(st.x_cur, st.low, st.high, chain.constraint, chain.log_factor))
    st.p.zero_()
    _grad_eval(st, chain, reg, st.x_cur, st.mw_cur, None, None, st.g_cur, 0.0, 0.0, 0)
    s = eng.sums.cpu().numpy()
    st.Ud, st.Um = float(s[1]), float(s[2])
    st.U = st.Ud + reg.alpha * st.Um
    st.d_cur.copy_(eng.d)
    st.synced = xcur


def sharded_proposal(chain, xcur, dt, L, alpha, trace=None):
    """hmc.py:85-177 over a row-sharded model; returns the reference's 6-tuple (dsyn holds this
    rank's rows)."""
    from ._engine import reg_params

    m = chain.model
    if chain.regularization in ("Smoothness", "TV") and int(np.prod(m.mshape)) != m.M:
        raise ValueError("Smoothness/TV are defined on the full (nz, ny, nx) grid and cannot "
                         "be used with a topography-carved model")
    reg = reg_params(chain.regularization, chain.constraint, m.mshape, alpha, chain.beta,
                     chain.log_factor)
    st = chain.cache.get("shard_state")
    if st is None:
        st = chain.cache["shard_state"] = _ShardState(chain, alpha)
    if st.synced is not xcur or chain.cache.get("shard_alpha") != alpha:
        _set_state(st, chain, reg, xcur)
        chain.cache["shard_alpha"] = alpha
    eng = st.eng
    n = len(xcur)
    pcur = np.random.randn(n) * chain.Sigma
    u = np.random.rand()
    st.p.copy_(eng.vec(pcur))
    tx = tu = None
    if trace is not None:
        tx, tu = np.zeros((L + 1, n)), np.zeros(L + 1)
        tx[0], tu[0] = np.asarray(xcur), st.U
    # p -= dt/2 grad(x_cur); x += dt p; clamp  -- the cached gradient is injected through eng.g
    eng.g.copy_(st.g_cur)
    half = reg_params("Damping", chain.constraint, m.mshape, 0.0, chain.beta, chain.log_factor)
    # grad = 2*g + 0*dR: feed g_cur/2 so the fused kernel reproduces the cached gradient exactly
    eng.g.mul_(0.5)
    eng.update(half, st.x_cur, st.mw_cur, st.apr, m.wmsq_dev, st.low, st.high, st.p, st.xa, st.mwa,
               None, 0.5 * dt, dt, 1)
    K0 = float(eng.sums[4])
    xin, xout, mwin, mwout = st.xa, st.xb, st.mwa, st.mwb
    for i in range(1, L + 1):
        last = i == L
        _grad_eval(st, chain, reg, xin, mwin, xout, mwout, st.gnew if last else None,
                   0.5 * dt if last else dt, dt, 0 if last else 1)
        if trace is not None:
            s = eng.sums.cpu().numpy()
            tx[i], tu[i] = xin[:n].cpu().numpy(), s[1] + alpha * s[2]
        if not last:
            xin, xout = xout, xin
            mwin, mwout = (mwout, mwin) if st.logc else (xin, xout)
    s = eng.sums.cpu().numpy()
    Ud, Um, Knew = float(s[1]), float(s[2]), float(s[3])
    Unew = Ud + alpha * Um
    Hcur, Hnew = K0 + st.U, Knew + Unew
    accept = bool(Hnew < Hcur or u < np.exp(-(Hnew - Hcur)))  # hmc.py:167
    if trace is not None:
        trace.update(x=tx, U=tu, Hcur=Hcur, Hnew=Hnew, L=L, accept=accept)
    chain.proposals.append((int(L), accept))
    if accept:
        st.x_cur.copy_(xin)
        if st.logc:
            st.mw_cur.copy_(mwin)
        st.g_cur.copy_(st.gnew)
        st.d_cur.copy_(eng.d)
        st.U, st.Ud, st.Um = Unew, Ud, Um
        xcur = st.x_cur[:n].cpu().numpy()
        st.synced = xcur
    return xcur, st.U, st.d_cur.cpu().numpy(), accept, st.Ud, st.Um
