"""Sample sinks of the HMC samplers.

The reference appends every accepted model to ``<save_folder><rank>/model.dat`` as one text row of M
``%.8f`` numbers (inversion/hmc.py:241-249; ~10 MB per sample at 1M voxels) and its post-processing
(example/uniformgrid/plot_uniform.py:44-135) reads the last ``last`` rows back, forms ``np.mean`` /
``np.std`` per voxel, forwards both through ``prism.gz`` and writes ``inversion_model.dat`` /
``inversion_anomaly.dat``.  Three sinks are offered here, selected per chain object with
``chain.output`` (``hmc.HamitonianMC``) / ``HMCBatch.output``:

``"text"``    the reference's files, byte for byte the same format (default);
``"binary"``  ``model.f64`` / ``misfit.f64``: raw little-endian float64 rows + ``samples.json``;
``"none"``    nothing is written (and, for the streaming batch sampler, nothing is copied back);

and independently ``SampleSink``: running mean / variance of the accepted models kept ON THE DEVICE
(`gi_stats_*`, Welford, gated by the Metropolis flag inside the sampler's own stream), which
reproduces plot_uniform.py's statistics without ever moving a sample off the GPU.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from .. import _lib

MISFIT_COLUMNS = ["U", "U_data", "U_model", "U_normed", "U_data_normed", "U_model_normed", "alpha"]


class SampleSink:
    """Per-chain running mean / std of m = WmInv @ mw over the accepted samples.

    `skip` / `take`: ignore the first `skip` accepted samples of every chain, then accumulate `take`
    (None: all) -- ``skip = ndraws + nsamples - last, take = last`` is plot_uniform.py:43-44."""

    def __init__(self, model, nslots=1, skip=0, take=None):
        self.torch = _lib.require_cuda()
        self.L = _lib.lib()
        self.model = model
        self.nslots = int(nslots)
        self.h = C.c_void_p()
        _lib.check(self.L.gi_stats_create(model.M, model.ld, self.nslots, C.byref(self.h)),
                   "gi_stats_create")
        self.user_window = skip != 0 or take is not None
        self.window(skip, take)

    def window(self, skip=0, take=None):
        _lib.check(self.L.gi_stats_window(self.h, int(skip), -1 if take is None else int(take)),
                   "gi_stats_window")

    def reset(self):
        _lib.check(self.L.gi_stats_reset(self.h), "gi_stats_reset")

    def add(self, mw, slot=0):
        """offer one accepted weighted model (numpy [M] or a padded device vector)"""
        t = self.torch
        if not hasattr(mw, "data_ptr"):
            v = t.zeros(self.model.ld, dtype=t.float64, device=self.model.Aw_pad.device)
            v[: self.model.M] = t.as_tensor(np.asarray(mw, dtype=np.float64), device=v.device)
            mw = v
        _lib.check(self.L.gi_stats_add(self.h, int(slot), _lib.ptr(mw), _lib.ptr(self.model.wminv_dev),
                                       _lib.stream_ptr()), "gi_stats_add")
        _lib.sync()

    def result(self, slot=None):
        """(mean[M], std[M], count) of one chain, or pooled over all chains (slot=None);
        np.mean / np.std(ddof=0) of plot_uniform.py:103-104"""
        M = self.model.M
        mean, std = np.zeros(M), np.zeros(M)
        cnt, seen = C.c_int64(), C.c_int64()
        _lib.check(self.L.gi_stats_result(self.h, -1 if slot is None else int(slot), _lib.ptr(mean),
                                          _lib.ptr(std), C.byref(cnt), C.byref(seen), _lib.stream_ptr()),
                   "gi_stats_result")
        self.seen = int(seen.value)
        return mean, std, int(cnt.value)

    def forward(self, slot=None):
        """(dpre_mean, dpre_std): forward data of the mean and std models, plot_uniform.py:112-114
        (`prism.gz` of a mesh carrying them = A m = Aw (Wm m)), this rank's observation rows"""
        t, m = self.torch, self.model
        f64 = dict(dtype=t.float64, device=m.Aw_pad.device)
        mw_mean, mw_std = t.zeros(m.ld, **f64), t.zeros(m.ld, **f64)
        cnt, seen = C.c_int64(), C.c_int64()
        _lib.check(self.L.gi_stats_result_dev(self.h, -1 if slot is None else int(slot),
                                              _lib.ptr(m.wm_dev), _lib.ptr(mw_mean), _lib.ptr(mw_std),
                                              C.byref(cnt), C.byref(seen), _lib.stream_ptr()),
                   "gi_stats_result_dev")
        eng = m.engine()
        out = []
        for v in (mw_mean, mw_std):
            _lib.check(self.L.gi_gemv_fwd(eng.plan, _lib.ptr(eng.Aw), _lib.ptr(v), _lib.ptr(eng.d),
                                          _lib.stream_ptr()), "gi_gemv_fwd")
            out.append(eng.d.cpu().numpy().copy())
        return out[0], out[1]

    def save(self, folder, slot=None):
        """inversion_model.dat (x, y, z, mean, std) and inversion_anomaly.dat (xobs, yobs, heights,
        dpre_mean, dpre_std, dobs - dpre_mean) exactly as plot_uniform.py:117-131 writes them
        (uncarved grids: one row per cell)."""
        m = self.model
        mean, std, _ = self.result(slot)
        dmean, dstd = self.forward(slot)
        os.makedirs(folder, exist_ok=True)
        if int(np.prod(m.mshape)) == m.M:
            zs, ys, xs = np.meshgrid(np.asarray(m.mzs)[:-1], np.asarray(m.mys)[:-1],
                                     np.asarray(m.mxs)[:-1], indexing="ij")
            np.savetxt(os.path.join(folder, "inversion_model.dat"),
                       np.c_[xs.ravel(), ys.ravel(), zs.ravel(), mean, std], fmt="%.8f", delimiter=" ")
        else:
            np.savetxt(os.path.join(folder, "inversion_model.dat"), np.c_[mean, std], fmt="%.8f",
                       delimiter=" ")
        lo, hi = m.rows
        np.savetxt(os.path.join(folder, "inversion_anomaly.dat"),
                   np.c_[np.asarray(m.lonobs)[lo:hi], np.asarray(m.latobs)[lo:hi],
                         np.asarray(m.heightobs)[lo:hi], dmean, dstd, m.dobs[lo:hi] - dmean],
                   fmt="%.8f", delimiter=" ")
        return mean, std, dmean, dstd

    def launches(self):
        return int(self.L.gi_stats_launch_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.gi_stats_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- file sinks -----------------------------------------------------------------------------------
class SampleWriter:
    """Appends (misfit row, model row) of one chain in the chosen format."""

    def __init__(self, folder, mode="text", M=None):
        if mode not in ("text", "binary", "none"):
            raise ValueError("output must be 'text', 'binary' or 'none'")
        self.folder, self.mode, self.M = folder, mode, M
        if mode == "none":
            return
        if not os.path.exists(folder):
            os.mkdir(folder)
        for name in ("model.dat", "model.f64", "misfit.f64", "samples.json"):
            # hmc.py:257-258 removes a stale model.dat (misfit.dat is appended to, as in the reference)
            if os.path.exists(os.path.join(folder, name)):
                os.remove(os.path.join(folder, name))
        if mode == "binary":
            with open(os.path.join(folder, "samples.json"), "w") as f:
                json.dump({"format": "gravinv3dhmc_b200 samples v1", "dtype": "<f8", "M": int(M),
                           "model": "model.f64", "misfit": "misfit.f64",
                           "misfit_columns": MISFIT_COLUMNS}, f)

    def append(self, misfit_row, model_row):
        if self.mode == "none":
            return
        if self.mode == "text":  # hmc.py:241-249
            with open(self.folder + "/" + "misfit" + ".dat", "a") as f:
                np.savetxt(f, np.asarray(misfit_row, dtype=np.float64).reshape(1, -1), fmt="%.8f",
                           delimiter=" ")
            with open(self.folder + "/" + "model" + ".dat", "a") as f:
                np.savetxt(f, np.asarray(model_row, dtype=np.float64).reshape(1, -1), fmt="%.8f",
                           delimiter=" ")
            return
        with open(os.path.join(self.folder, "misfit.f64"), "ab") as f:
            f.write(np.asarray(misfit_row, dtype="<f8").tobytes())
        with open(os.path.join(self.folder, "model.f64"), "ab") as f:
            f.write(np.asarray(model_row, dtype="<f8").tobytes())


def read_samples(folder, last=None):
    """(misfit [n, 7], models [n, M]) of a chain folder in either format; `last` keeps the final rows
    (plot_uniform.py:49-54 skips the first nsamples - last lines of model.dat)."""
    meta = os.path.join(folder, "samples.json")
    if os.path.exists(meta):
        info = json.load(open(meta))
        M = int(info["M"])
        models = np.fromfile(os.path.join(folder, info["model"]), dtype=info["dtype"]).reshape(-1, M)
        misfit = np.fromfile(os.path.join(folder, info["misfit"]), dtype=info["dtype"]).reshape(-1, 7)
    else:
        misfit = np.atleast_2d(np.loadtxt(os.path.join(folder, "misfit.dat")))
        models = np.atleast_2d(np.loadtxt(os.path.join(folder, "model.dat")))
    if last is not None:
        misfit, models = misfit[-last:], models[-last:]
    return misfit, models


def posterior_from_samples(models):
    """plot_uniform.py:103-104"""
    return np.mean(models, axis=0), np.std(models, axis=0)
