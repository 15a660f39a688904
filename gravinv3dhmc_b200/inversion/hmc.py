"""Hamiltonian Monte Carlo sampling with a device-resident leapfrog loop.

Mirror of the reference's inversion/hmc.py: `HMCSample(...)` (:358-403) and class `HamitonianMC`
(:29-343) with the same arguments, the same RNG call order (legacy global numpy stream seeded with
`seed + myrank`: per proposal `randint(Lmin, Lmax+1)`, `randn(n)`, `rand()`; :260,297,95,165), the
same accept-counting loop (:295-334), the same output files (`<save_folder><rank>/misfit.dat`,
`model.dat`, `%.8f`, space separated) and the same progress line (:336-342).

The trajectory itself (momentum half step, L position/momentum updates with clamp-and-flip, the
two streaming passes over Aw per step, regulariser gradient, Metropolis test) runs in
libgravinv_b200.so (`gi_hmc_propose`); the host only draws the random numbers and writes files.
With a row-sharded model (`GravMagModule(..., shard=(rank, world))`) the same loop is driven
through the building-block kernels with two NCCL all-reduces per gradient evaluation.

`rng="philox"` (extension) draws momentum and the uniform on the device instead; the chain is
then statistically equivalent but not bit-compatible with numpy's MT19937 stream.
"""
from __future__ import annotations

import ctypes as C
import sys

import numpy as np
from scipy.sparse import coo_matrix

from .. import _lib
from ._engine import reg_params


class HamitonianMC:
    def __init__(self, UserDefinedModel):
        self.invert_Mass = None
        self.model = UserDefinedModel
        self.dobs = np.zeros(2)
        self.boundaries = np.zeros((2, 2))
        self.dt = None
        self.Lrange = [10, 50]
        self.seed = None
        self.myrank = None
        self.save_folder = None
        self.cache = {}
        self.rng = "numpy"
        self.quiet = False
        self.plotsamples = False
        self.im = [0, 0]
        self._h = None          # gi_hmc handle (single GPU)
        self._synced = None     # host array whose contents the device state mirrors
        self.proposals = []     # (L, accept) log of this chain
        self.output = "text"    # "text" (the reference's files) | "binary" | "none"  (inversion/sink.py)
        self.sink = None        # SampleSink: on-device posterior mean / std of the accepted models
        self._philox_counter = 0

    # ---- reference helper API ---------------------------------------------------------------
    def _kinetic(self, p):
        """hmc.py:44-50 (identity inverse mass)"""
        return np.dot(self.invert_Mass @ p, p) * 0.5

    def _misfit_and_grad(self, x, alpha):
        """hmc.py:71-78"""
        return self.model.misfit_and_grad(x, self.aprior_model, self.low, self.high,
                                          self.constraint, self.log_factor, alpha,
                                          regulization=self.regularization, beta=self.beta)

    def _kernelw(self):
        return self.model.kernelw()

    # ---- device handle ----------------------------------------------------------------------
    def _reg(self, alpha):
        # JointModule also knows "MS1" / "MStry" (potential.py:1701-1736): MS kernels, see _ensure_handle
        name = getattr(self.model, "extra_regs", {}).get(self.regularization, self.regularization)
        return reg_params(name, self.constraint, self.model.mshape, alpha, self.beta, self.log_factor)

    def _sharded(self):
        return getattr(self.model, "world", 1) > 1

    def _ensure_handle(self, alpha):
        m = self.model
        if self.regularization in ("Smoothness", "TV") and getattr(m, "nocenter", False):
            raise AttributeError("'JointModule' object has no attribute 'fd3d'")  # potential.py:1753,1765
        if self.regularization in ("Smoothness", "TV") and int(np.prod(m.mshape)) != m.M:
            raise ValueError("Smoothness/TV are defined on the full (nz, ny, nx) grid and cannot "
                             "be used with a topography-carved model")
        if self._h is not None:
            if self._alpha != alpha:
                reg = self._reg(alpha)
                _lib.check(_lib.lib().gi_hmc_set_reg(self._h, C.byref(reg)), "gi_hmc_set_reg")
                self._alpha = alpha
                self._synced = None
            return
        _lib.require_cuda()
        L = _lib.lib()
        cfg = _lib.HmcConfig(m.n_total, m.M, m.ld, 1 if m.fixed else 0,
                             1 if getattr(m, "nocenter", False) else 0, self._reg(alpha))
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        wmsq = np.ones(m.M) if self.regularization == "MStry" else m.WmSquare.diagonal()
        self._host = dict(dobs=f(getattr(m, "dobs_sampler", m.dobs)), low=f(self.low), high=f(self.high),
                          apr=f(self.aprior_model), wmsq=f(wmsq))
        fix = f(m.grav_fix) if m.fixed else None
        if fix is not None:
            self._host["fix"] = fix
        h = C.c_void_p()
        _lib.check(L.gi_hmc_create(C.byref(cfg), _lib.ptr(m.Aw_pad), _lib.ptr(self._host["dobs"]),
                                   _lib.ptr(fix), _lib.ptr(self._host["low"]),
                                   _lib.ptr(self._host["high"]), _lib.ptr(self._host["apr"]),
                                   _lib.ptr(self._host["wmsq"]), _lib.stream_ptr(), C.byref(h)),
                   "gi_hmc_create")
        self._h = h
        self._alpha = alpha
        if self.sink is not None:
            _lib.check(L.gi_hmc_attach_stats(h, self.sink.h, int(getattr(self, "sink_slot", 0)),
                                             _lib.ptr(m.wminv_dev)), "gi_hmc_attach_stats")
        if m.wavelet in ("1D", "3D"):  # potential.py:693-696: forward through the compressed kernel
            nz, ny, nx = (int(v) for v in m.mshape)
            cp = m.Awcp
            _lib.check(L.gi_hmc_set_wavelet(h, 1 if m.wavelet == "1D" else 3, nz, ny, nx,
                                            _lib.ptr(cp.indptr), _lib.ptr(cp.indices),
                                            _lib.ptr(cp.data), cp.shape[1]), "gi_hmc_set_wavelet")

    def close(self):
        if self._h is not None:
            _lib.lib().gi_hmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _sync_state(self, xcur):
        if self._synced is not xcur:
            x = np.ascontiguousarray(xcur, dtype=np.float64)
            _lib.check(_lib.lib().gi_hmc_set_state(self._h, _lib.ptr(x)), "gi_hmc_set_state")
            self._synced = xcur

    # ---- one proposal -------------------------------------------------------------------------
    def _leapfrog(self, xcur, dt, L, alpha, fignum=0, trace=None):
        """One HMC proposal from `xcur` (hmc.py:85-177).  Returns
        (xcur, U, dsyn, AcceptFlag, U_data, U_model) like the reference; `trace` (dict, extension)
        receives the per-leapfrog positions/potentials and Hcur/Hnew."""
        if self._sharded():
            return self._leapfrog_sharded(xcur, dt, L, alpha, trace)
        self._ensure_handle(alpha)
        self._sync_state(xcur)
        lib = _lib.lib()
        n = len(xcur)
        res = _lib.HmcResult()
        tx = tu = None
        if trace is not None:
            tx = np.zeros((L + 1, n))
            tu = np.zeros(L + 1)
        if self.rng == "philox":
            if trace is not None:
                raise ValueError("tracing needs injected draws (rng='numpy')")
            _lib.check(lib.gi_hmc_propose_philox(self._h, int(self.seed), self._philox_counter,
                                                 float(self.Sigma), int(L), float(dt),
                                                 C.byref(res)), "gi_hmc_propose_philox")
            self._philox_counter += 1
        else:
            pcur = np.random.randn(n) * self.Sigma       # hmc.py:95
            u = np.random.rand()                         # hmc.py:165 (nothing else draws in between)
            _lib.check(lib.gi_hmc_propose(self._h, _lib.ptr(pcur), int(L), float(dt), float(u),
                                          C.byref(res), _lib.ptr(tx), _lib.ptr(tu)),
                       "gi_hmc_propose")
        if trace is not None:
            trace.update(x=tx, U=tu, Hcur=res.Hcur, Hnew=res.Hnew, L=L, accept=bool(res.accept))
        accept = bool(res.accept)
        self.proposals.append((int(L), accept))
        if accept:
            xnew = np.empty(n)
            self._dsyn = np.empty(self.model.n_total)
            _lib.check(lib.gi_hmc_get_state(self._h, _lib.ptr(xnew), _lib.ptr(self._dsyn), None),
                       "gi_hmc_get_state")
            xcur = xnew
            self._synced = xcur
        elif getattr(self, "_dsyn", None) is None:
            self._dsyn = np.empty(self.model.n_total)
            _lib.check(lib.gi_hmc_get_state(self._h, None, _lib.ptr(self._dsyn), None),
                       "gi_hmc_get_state")
        return xcur, res.U, self._dsyn, accept, res.U_data, res.U_model

    def _leapfrog_sharded(self, xcur, dt, L, alpha, trace=None):
        """Row-sharded trajectory: identical replicated M-vector state on every rank, the kernel
        rows split by observation; two all-reduces per gradient evaluation."""
        from .sharded import sharded_proposal

        return sharded_proposal(self, xcur, dt, L, alpha, trace)

    # ---- file output (hmc.py:241-249) -------------------------------------------------------------
    def _save_models_add(self, x):
        with open(self.save_folder + "/" + "model" + ".dat", "a") as f:
            np.savetxt(f, x, fmt="%.8f", delimiter=" ")

    def _save_misfit_add(self, misfit):
        with open(self.save_folder + "/" + "misfit" + ".dat", "a") as f:
            np.savetxt(f, misfit, fmt="%.8f", delimiter=" ")

    # ---- the chain ------------------------------------------------------------------------------
    def sample(self, nsamples, ndraws, **kwargs):
        """hmc.py:252-343"""
        from .sink import SampleWriter

        writes = (not self._sharded()) or self.model.rank == 0
        writer = SampleWriter(self.save_folder, self.output if writes else "none", self.model.M)
        if self.sink is not None:
            if self._sharded():
                raise NotImplementedError("the device sink needs the single-GPU handle")
            if self._h is not None:  # created before the sink was set
                _lib.check(_lib.lib().gi_hmc_attach_stats(
                    self._h, self.sink.h, int(getattr(self, "sink_slot", 0)),
                    _lib.ptr(self.model.wminv_dev)), "gi_hmc_attach_stats")
            if not self.sink.user_window:  # the recorded samples: hmc.py:318
                self.sink.window(ndraws, nsamples)
        np.random.seed(self.seed)
        _, WmInv, _ = self._kernelw()
        wminv = WmInv.diagonal()
        mw = self.initial_model
        if self.constraint == "logarithmic":
            x = (1 / self.log_factor) * np.log((mw - self.low) / (self.high - mw))
            self._print("Using logarithmic boundary constraint.")
        elif self.constraint == "mandatory":
            x = mw
            self._print("Using mandatory boundary constraint.")
        else:
            raise ValueError("Please choose right boundary constraint(mandatory, logarithmic)!")
        data_size = self.dobs.shape[0]
        model_size = self.initial_model.shape[0]
        misfit = np.zeros((1, 7))
        m_cache = np.zeros((1, len(x)))
        ncount = 0
        i = 0
        alpha = self.RegulFactor
        max_proposals = kwargs.get("max_proposals")
        while i < ndraws + nsamples:
            if max_proposals is not None and ncount >= max_proposals:
                break
            L = np.random.randint(self.Lrange[0], self.Lrange[1] + 1)
            x, U, _, AcceptFlag, U_data, U_model = self._leapfrog(x, self.dt, L, alpha, i)
            U_data_normed = U_data / data_size
            U_model_normed = U_model / model_size
            U_normed = U_data_normed + alpha * U_model_normed
            if AcceptFlag:
                if i >= ndraws and writer.mode != "none":
                    misfit[0, :] = [U, U_data, U_model, U_normed, U_data_normed, U_model_normed, alpha]
                    if self.constraint == "logarithmic":
                        mw = (self.low + self.high * np.e ** (self.log_factor * x)) / \
                             (1 + np.e ** (self.log_factor * x))
                    else:
                        mw = x
                    m_cache[0, :] = wminv * mw      # m = WmInv @ mw
                    writer.append(misfit, m_cache)
                i += 1
            ncount += 1
            if i > -1:
                msg = "chain {}: {:.2%}, misfit(total, data, alpha, model)=({:.7f},{:.7f},{:.2f},{:.7f}) " \
                      "-- accept ratio {:.2%}\n".format(self.myrank, i / (ndraws + nsamples), U_normed,
                                                       U_data_normed, alpha, U_model_normed, i / ncount)
                self._print(msg)
        self.x_final = x
        return x

    def _print(self, msg):
        if not self.quiet:
            print(msg)
            sys.stdout.flush()


def setup_chain(model, delta, Lrange, initial_model, aprior_model, boundaries, constraint,
                log_factor, dobs, adaptiveRegul, RegulRate, RegulFactor, regularization, beta, seed,
                Sigma, nbest=100, myrank=0, save_folder="mychain", plotsamples=False, im=[0, 0],
                rng="numpy", quiet=False):
    """The chain object of hmc.py:358-401 (everything HMCSample does before `chain.sample`)."""
    if constraint not in _lib.CONSTRAINTS:
        raise ValueError("Please choose right boundary constraint(mandatory, logarithmic)!")
    if regularization not in _lib.REG_KINDS and regularization not in getattr(model, "extra_regs", {}):
        raise ValueError("Please choose regularization from 'MS','Damping', 'Smoothness', 'TV'.")
    chain = HamitonianMC(model)
    chain.myrank = myrank
    chain.save_folder = save_folder + str(myrank)
    chain.seed = seed + myrank
    chain.nbest = nbest
    boundaries = np.asarray(boundaries, dtype=np.float64)
    nt = boundaries.shape[0]
    chain.boundaries = boundaries
    chain.constraint = constraint
    chain.log_factor = log_factor
    chain.Lrange = Lrange
    chain.dt = delta
    chain.Sigma = Sigma
    chain.adaptiveRegul = adaptiveRegul
    chain.RegulRate = RegulRate
    chain.RegulFactor = RegulFactor
    chain.regularization = regularization
    chain.beta = beta
    row = np.arange(0, nt)
    chain.invert_Mass = coo_matrix((np.ones(nt), (row, row))).tocsr()
    _, _, Wm = chain._kernelw()
    chain.low = Wm @ boundaries[:, 0]
    chain.high = Wm @ boundaries[:, 1]
    chain.im = im
    chain.initial_model = Wm @ np.asarray(initial_model, dtype=np.float64)
    chain.aprior_model = Wm @ np.asarray(aprior_model, dtype=np.float64)
    chain.dobs = np.asarray(dobs, dtype=np.float64)
    chain.plotsamples = plotsamples
    chain.rng = rng
    chain.quiet = quiet
    return chain


def HMCSample(model, nsamples, ndraws, delta, Lrange, initial_model, aprior_model, boundaries,
              constraint, log_factor, dobs, adaptiveRegul, RegulRate, RegulFactor, regularization,
              beta, seed, Sigma, nbest=100, myrank=0, save_folder="mychain", plotsamples=False,
              im=[0, 0], rng="numpy", quiet=False, max_proposals=None):
    """HMC sampling function -- hmc.py:358-403.  `adaptiveRegul`, `RegulRate` and `nbest` are
    accepted and (as in the reference's sampler) unused.  Returns the chain object (extension; the
    reference returns None)."""
    chain = setup_chain(model, delta, Lrange, initial_model, aprior_model, boundaries, constraint,
                        log_factor, dobs, adaptiveRegul, RegulRate, RegulFactor, regularization,
                        beta, seed, Sigma, nbest, myrank, save_folder, plotsamples, im, rng, quiet)
    chain.sample(nsamples, ndraws, max_proposals=max_proposals)
    return chain
