"""Potential-energy model of the gravity inversion: sensitivity matrix, sensitivity weighting,
data misfit and the four regularisers, all evaluated on the GPU.

Mirror of the reference's inversion/potential.py class `GravMagModule` (:34-845): same constructor
arguments (`dobs, mrange, mspacing, obsurface, fixed, grav_fix, mratio, mseg, mdivisionsection,
weightfactor, coordinate, njobs, field, mangle, wavelet, mtopo=`), same attributes (`Aw, Wm, WmInv,
WmSquare, mshape, mxs, mys, mzs, mask, mesh`), same methods (`kernelw, data_all, model_*_all,
misfit_and_grad, fd3d`), same exceptions.

What differs by design:
  * `Aw` is a CUDA tensor ([N, M] view of a zero-padded [N, ld] buffer, `Aw_pad`) assembled and
    weighted in place on the device -- the unweighted `A` is never held separately;
  * `njobs` is accepted and ignored; `field="magnetic"` is supported for `coordinate="cartesian"`
    (total-field anomaly kernel along `mangle`, potential.py:125-149); the reference's spherical
    magnetic branch is an empty stub (potential.py:109-111) and raises its ValueError here;
  * `shard=(rank, world)` + `group=` (extension) keeps only this rank's observation rows
    (contiguous chunks like gravmag/prism.py:986-996) and sums the column norms with one
    all-reduce; everything else is replicated.
"""
from __future__ import annotations

import time

import numpy as np
from scipy.sparse import coo_matrix

from .. import _lib, mesher
from ..gravmag import prism, tesseroid
from ..gravmag._common import split_rows
from ._engine import BlockEngine, reg_params


class GravMagModule:
    def __init__(self, dobs, mrange, mspacing, obsurface, fixed=False, grav_fix=[], mratio=1,
                 mseg=False, mdivisionsection=[], weightfactor=0.5, coordinate="cartesian", njobs=1,
                 field="gravity", mangle=(90, 0), wavelet=False, **kwargs):
        shard = kwargs.pop("shard", None)
        self.group = kwargs.pop("group", None)
        self.verbose = kwargs.pop("verbose", True)
        timing = kwargs.pop("timing", False)  # extension: CUDA-event times of the setup kernels
        self.timing = {}
        self.dobs = np.asarray(dobs, dtype=np.float64)
        self.fixed = fixed
        self.grav_fix = grav_fix
        self.mrange = mrange
        self.mspacing = mspacing
        self.mratio = mratio
        self.weightfactor = weightfactor
        self.mseg = mseg
        self.mdivisionsection = mdivisionsection
        self.lonobs, self.latobs, self.heightobs = obsurface[0], obsurface[1], obsurface[2]
        self.inc, self.dec = mangle[0], mangle[1]
        self.njobs = njobs
        self.topocarve = False
        self.wavelet = wavelet
        self.coordinate = coordinate

        magnetic = field == "magnetic" and coordinate == "cartesian"
        if (field != "gravity" and not magnetic) or coordinate not in ("spherical", "cartesian"):
            # potential.py:151 (the reference's spherical magnetic branch is an empty stub)
            raise ValueError("Please choose coordinate from(cartesian, spherical) and field "
                             "from(gravity, magnetic)!")
        self._say("Calculating {} field in {} coordinate.".format(field, coordinate))
        spherical = coordinate == "spherical"
        if magnetic:  # potential.py:125-149: plain PrismMesh, total-field anomaly kernel
            mesh = mesher.PrismMesh(mrange, mspacing, mratio)
        elif spherical:
            mesh = (mesher.TesseroidMeshSegment(mrange, mspacing, mdivisionsection) if mseg
                    else mesher.TesseroidMesh(mrange, mspacing, mratio))
        else:
            mesh = (mesher.PrismMeshSegment(mrange, mspacing, mdivisionsection) if mseg
                    else mesher.PrismMesh(mrange, mspacing, mratio))
        for key, value in kwargs.items():  # potential.py:94-98 / 116-120: any extra kwarg = topography
            self.topocarve = True
            self.mask = mesh.carvetopo(value[0], value[1], value[2])
        if magnetic:
            from ..utils import ang2vec

            mesh.addprop("magnetization", ang2vec(np.zeros(mesh.size), self.inc, self.dec))
        else:
            mesh.addprop("density", np.zeros(mesh.size))
        self.mesh = mesh

        n_total = len(self.lonobs)
        self.n_total = n_total
        self.rank, self.world = (0, 1) if shard is None else (int(shard[0]), int(shard[1]))
        self.rows = split_rows(n_total, self.world)[self.rank]

        self._say("Start of calculate kernel")
        start = time.time()
        ev = self._events(timing)
        table = mesh.bounds_table()
        if spherical:
            ncols = table.shape[0]
            table, _ = tesseroid._check_table(table)
            Apad, M = tesseroid.assemble(self.lonobs, self.latobs, self.heightobs, table,
                                         rows=self.rows, ncols=ncols)
        elif magnetic:
            from ..utils import dircos

            Apad, M = prism.assemble_field("tf", self.lonobs, self.latobs, self.heightobs, table,
                                           vec=dircos(self.inc, self.dec), rows=self.rows)
        else:
            # structured meshes share their corners: one evaluation per mesh node (bit-identical)
            out = prism.assemble_grid(self.lonobs, self.latobs, self.heightobs, mesh, rows=self.rows)
            Apad, M = out if out is not None else prism.assemble(
                self.lonobs, self.latobs, self.heightobs, table, rows=self.rows)
        self._lap(ev, "assemble_ms")
        self._say("kernel.shape ({}, {})".format(n_total, M))
        self._say("End of calculate kernel:%.6f s" % (time.time() - start))
        self.M, self.ld = M, int(Apad.shape[1])
        self.mshape = mesh.shape
        self.mxs, self.mys, self.mzs = mesh.get_xs(), mesh.get_ys(), mesh.get_zs()

        self._say("Start to weight kernel")
        start = time.time()
        self.Aw_pad = Apad
        ev = self._events(timing)
        self.sensitivityWeighting()
        self._lap(ev, "weight_ms")
        self._say("End of weighting kernel: %.6f s" % (time.time() - start))

        self._engine = None
        self._mg_cache = {}
        if wavelet in ("1D", "3D"):
            from ..gravmag import compressor1D as cp1D, compressor3D as cp3D

            self._say("Using {} wavelet to compress kernel.".format(wavelet))
            self.Awcp = (cp1D.kernelcompressor(self.Aw) if wavelet == "1D"
                         else cp3D.kernelcompressor(self.Aw, self.mshape))

    def _say(self, msg):
        if self.verbose:
            print(msg)

    @staticmethod
    def _events(enabled):
        if not enabled:
            return None
        torch = _lib.require_cuda()
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def _lap(self, start, key):
        if start is None:
            return
        torch = _lib.require_cuda()
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        e.synchronize()
        self.timing[key] = start.elapsed_time(e)

    def set_dobs(self, dobs):
        """replace the observed data (extension; e.g. synthetic data generated from this kernel)"""
        self.dobs = np.asarray(dobs, dtype=np.float64)
        self._engine = None

    def forward_local(self, mw):
        """device vector Aw[rows] @ mw for this rank's observation rows (extension)"""
        eng = self.engine()
        mw_d = eng.vec(mw)
        _lib.check(eng.L.gi_gemv_fwd(eng.plan, _lib.ptr(eng.Aw), _lib.ptr(mw_d), _lib.ptr(eng.d),
                                     _lib.stream_ptr()), "gi_gemv_fwd")
        return eng.d.clone()

    # ------------------------------------------------------------------ weighting
    def sensitivityWeighting(self):
        """Wm = diag((sum_j A_ji^2)^weightfactor), Aw = A Wm^-1 (potential.py:232-264), in place on
        the device.  Column sums of squares are partial per row shard -> one all-reduce."""
        torch = _lib.require_cuda()
        L = _lib.lib()
        A = self.Aw_pad
        n, ld = (int(v) for v in A.shape)
        s = _lib.stream_ptr()
        f64 = dict(dtype=torch.float64, device=A.device)
        sumsq = torch.zeros(ld, **f64)
        _lib.check(L.gi_colsumsq(_lib.ptr(A), n, self.M, ld, _lib.ptr(sumsq), 0, s), "gi_colsumsq")
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(sumsq, op=dist.ReduceOp.SUM, group=self.group)
        wm, wminv, wmsq = (torch.zeros(ld, **f64) for _ in range(3))
        _lib.check(L.gi_weights_from_sumsq(_lib.ptr(sumsq), self.M, float(self.weightfactor),
                                           _lib.ptr(wm), _lib.ptr(wminv), _lib.ptr(wmsq), s),
                   "gi_weights_from_sumsq")
        if self.M and float(wm[self.M - 1]) == 0.0:
            # potential.py:247-251 leaves the scalar 0 in ADiagInv when the LAST entry is zero and
            # coo_matrix then fails; fail the same way instead of dividing by zero silently
            raise ValueError("sensitivity weighting: the last column of the kernel is all zero")
        _lib.check(L.gi_scale_columns(_lib.ptr(A), n, self.M, ld, _lib.ptr(wminv), s),
                   "gi_scale_columns")
        _lib.sync()
        self.wm_dev, self.wminv_dev, self.wmsq_dev = wm, wminv, wmsq
        self.Aw = A[:, : self.M]
        row = np.arange(0, self.M)
        diag = lambda v: coo_matrix((v[: self.M].cpu().numpy(), (row, row)),
                                    shape=(self.M, self.M)).tocsr()
        self.Wm, self.WmInv, self.WmSquare = diag(wm), diag(wminv), diag(wmsq)

    # ------------------------------------------------------------------ fd3d
    @staticmethod
    def fd3d(shape):
        """Forward-difference matrix of potential.py:266-361 (rows: per layer x-differences then
        y-differences, then all z-differences; +1 at the cell, -1 at the next).  The CUDA stencil in
        gi_update applies D^T D / D^T(t/sqrt(t^2+beta)) without forming it; this host builder is
        kept for API compatibility and for tests."""
        nz, ny, nx = shape
        per_layer = (nx - 1) * ny + (ny - 1) * nx
        nderivs = per_layer * nz + nx * ny * (nz - 1)
        idx = np.arange(nz * ny * nx).reshape(nz, ny, nx)
        rows, c0, c1 = [np.zeros(0, dtype=np.int64)], [np.zeros(0, dtype=np.int64)], \
            [np.zeros(0, dtype=np.int64)]
        for k in range(nz):
            a = idx[k, :, :-1].ravel()
            rows.append(per_layer * k + np.arange(a.size)); c0.append(a); c1.append(a + 1)
            b = idx[k, :-1, :].ravel()
            rows.append(per_layer * k + a.size + np.arange(b.size)); c0.append(b); c1.append(b + nx)
        front = per_layer * nz
        for k in range(nz - 1):
            a = idx[k].ravel()
            rows.append(front + nx * ny * k + np.arange(a.size)); c0.append(a); c1.append(a + nx * ny)
        rows, c0, c1 = (np.concatenate(v) for v in (rows, c0, c1))
        I = np.concatenate([rows, rows])
        J = np.concatenate([c0, c1])
        V = np.concatenate([np.ones(rows.size), -np.ones(rows.size)])
        return coo_matrix((V, (I, J)), (nderivs, nx * ny * nz)).tocsr()

    # ------------------------------------------------------------------ sampler duck type
    def kernelw(self):
        """(Aw, WmInv, Wm) -- potential.py:584-589.  Aw is a CUDA tensor."""
        return self.Aw, self.WmInv, self.Wm

    def engine(self):
        if self._engine is None:
            lo, hi = self.rows
            fix = None
            if self.fixed:
                fix = np.asarray(self.grav_fix, dtype=np.float64)[lo:hi]
            self._engine = BlockEngine(self.Aw_pad, self.M, self.dobs[lo:hi], float(np.mean(self.dobs)),
                                       self.n_total, fix, self.group)
        return self._engine

    def _forward(self, mw_dev):
        """device forward data of the wavelet-compressed kernel (None without wavelet)"""
        if self.wavelet == "1D":
            from ..gravmag import compressor1D as cp1D
            return cp1D.modelcompressor(mw_dev[: self.M], self.Awcp)
        if self.wavelet == "3D":
            from ..gravmag import compressor3D as cp3D
            return cp3D.modelcompressor(mw_dev[: self.M], self.Awcp, self.mshape)
        return None

    def _evaluate(self, mw, mwapr, alpha, regularization, beta, constraint="mandatory",
                  log_factor=0.0, wmsq=None):
        """(U, grad, dpre, Ud, Um) at the weighted model `mw` (numpy)."""
        eng = self.engine()
        reg = reg_params(regularization, "mandatory", self.mshape, alpha, beta, log_factor)
        if regularization in ("Smoothness", "TV") and int(np.prod(self.mshape)) != self.M:
            raise ValueError("Smoothness/TV are defined on the full (nz, ny, nx) grid and cannot "
                             "be used with a topography-carved model")
        mw_d, apr_d = eng.vec(mw), eng.vec(mwapr)
        eng.data_pass(mw_d, self._forward(mw_d))
        pm, grad = eng.vec(), eng.vec()
        eng.update(reg, mw_d, mw_d, apr_d, self.wmsq_dev if wmsq is None else wmsq, None, None, pm, None,
                   None, grad, 0.0, 0.0, 0)
        sums = eng.sums.cpu().numpy()
        Ud, Um = float(sums[1]), float(sums[2])
        return (Ud + alpha * Um, grad[: self.M].cpu().numpy(), eng.d.cpu().numpy(), Ud, Um)

    def misfit_and_grad(self, x, mwapr, low, high, constraint, log_fator, alpha,
                        regulization="Damping", beta=0.01):
        """misfit, grad, dpre, data_value, model_value -- potential.py:812-845 (the gradient is
        taken with respect to mw and, as in the reference, NOT chain-ruled to x)."""
        x = np.asarray(x, dtype=np.float64)
        if constraint == "logarithmic":
            mw = (low + high * np.e ** (log_fator * x)) / (1 + np.e ** (log_fator * x))
        elif constraint == "mandatory":
            mw = x
        else:
            raise ValueError("Please choose right boundary constraint(mandatory, logarithmic)!")
        if regulization not in _lib.REG_KINDS:
            raise ValueError("Please choose regularization from 'MS','Damping', 'Smoothness', 'TV'.")
        return self._evaluate(mw, mwapr, alpha, regulization, beta)

    def data_all(self, mw):
        """dpre, data_value, data_gradient -- potential.py:688-717"""
        eng = self.engine()
        mw_d = eng.vec(mw)
        eng.data_pass(mw_d, self._forward(mw_d))
        sums = eng.sums.cpu().numpy()
        return eng.d.cpu().numpy(), float(sums[1]), 2.0 * eng.g[: self.M].cpu().numpy()

    def _model_all(self, regularization, mw, mwapr, beta):
        eng = self.engine()
        reg = reg_params(regularization, "mandatory", self.mshape, 1.0, beta, 0.0)
        mw_d, apr_d = eng.vec(mw), eng.vec(mwapr)
        eng.g.zero_()
        pm, grad = eng.vec(), eng.vec()
        eng.update(reg, mw_d, mw_d, apr_d, self.wmsq_dev, None, None, pm, None, None, grad, 0.0,
                   0.0, 0)
        return float(eng.sums[2]), grad[: self.M].cpu().numpy()

    def model_MS_all(self, mw, mwapr, beta):
        """potential.py:719-736"""
        return self._model_all("MS", mw, mwapr, beta)

    def model_Damping_all(self, mw, mwapr):
        """potential.py:775-784"""
        return self._model_all("Damping", mw, mwapr, 0.0)

    def model_Smoothness_all(self, mw, mwapr):
        """potential.py:786-796"""
        return self._model_all("Smoothness", mw, mwapr, 0.0)

    def model_TV_all(self, mw, mwapr, beta):
        """potential.py:798-810"""
        return self._model_all("TV", mw, mwapr, beta)

    # value-only variants used by the reference's (unused) adaptive-alpha hooks, potential.py:591-686
    def data(self, x, low, high, constraint, log_fator):
        x = np.asarray(x, dtype=np.float64)
        if constraint == "logarithmic":
            x = (low + high * np.e ** (log_fator * x)) / (1 + np.e ** (log_fator * x))
        elif constraint != "mandatory":
            raise ValueError("Please choose right boundary constraint(mandatory, logarithmic)!")
        return self.data_all(x)[1]


class JointModule(GravMagModule):
    """Joint gravity (gz) + magnetic (total-field anomaly) model on ONE prism mesh: mirror of the
    reference's `JointModule` (inversion/potential.py:847-1812) -- same constructor arguments
    (`dobs_gz, dobs_tf, mrange, mspacing, obsurface, mratio, coordinate, njobs, mangle, wavelet,
    mtopo=`), same attributes (`kernel_gz, kernel_tf, A, Aw, dobs, dobsw, Wm, WmInv, WmSquare, Wb,
    meshrho, meshmag, mshape, mxs, mys, mzs`), same sampler duck type (`kernelw`, `misfit_and_grad`).

    The kernel is the block matrix [[K_gz, 0], [0, K_tf]] of (2N x 2M) (potential.py:938-941; the
    reference holds it dense, so does this class -- on the device); the model vector is
    [density (M) | magnetisation intensity (M)].  `weightKDM` (potential.py:1003-1067): Wm = column
    norms of A, Wb = 1 on the gz rows and std(K_gz) / std(K_tf) on the tf rows, Aw = Wb A Wm^-1,
    dobsw = Wb [dobs_gz | dobs_tf].  The data term has NO mean removal (`data_all`,
    potential.py:1665-1680: |Aw mw - dobsw|^2), unlike `GravMagModule`.

    Reference behaviour kept: `coordinate="spherical"` fails with UnboundLocalError (`kernel_tf` is never
    assigned there, potential.py:885-899); Smoothness / TV fail with AttributeError (`self.fd3d`
    does not exist on JointModule, potential.py:1753,1765); `regulization` also accepts "MS1" (the
    matrix form of MS: same value and gradient) and "MStry" (MS without the Wm weighting,
    potential.py:1701-1736)."""

    nocenter = True   # sampler handles: r = d - dobs (gi_hmc_config.nocenter)
    extra_regs = {"MS1": "MS", "MStry": "MS"}

    def __init__(self, dobs_gz, dobs_tf, mrange, mspacing, obsurface, mratio=1, coordinate="cartesian",
                 njobs=1, mangle=(90, 0), wavelet=False, **kwargs):
        from ..utils import ang2vec, dircos

        if kwargs.pop("shard", None) is not None:
            raise NotImplementedError("JointModule is a single-GPU model (the block kernel is not row-sharded)")
        self.group = kwargs.pop("group", None)
        self.verbose = kwargs.pop("verbose", True)
        self.timing = {}
        self.dobs_gz = np.asarray(dobs_gz, dtype=np.float64)
        self.dobs_tf = np.asarray(dobs_tf, dtype=np.float64)
        self.mrange, self.mspacing, self.mratio = mrange, mspacing, mratio
        self.lonobs, self.latobs, self.heightobs = obsurface[0], obsurface[1], obsurface[2]
        self.inc, self.dec = mangle[0], mangle[1]
        self.njobs = njobs
        self.topocarve = False
        self.wavelet = wavelet
        self.fixed, self.grav_fix = False, []
        self.coordinate = coordinate
        self.weightfactor = 0.5
        if coordinate == "spherical":
            self._say("Joint inversion in {} coordinate.".format(coordinate))
            # potential.py:885-899 assembles kernel_gz only; `self.kernel_tf = kernel_tf` then fails
            raise UnboundLocalError("cannot access local variable 'kernel_tf' where it is not associated "
                                    "with a value")
        if coordinate != "cartesian":
            raise ValueError("Please choose coordinate from(cartesian, spherical)!")  # potential.py:923
        self._say("Joint inversion in {} coordinate.".format(coordinate))
        mesh = mesher.PrismMesh(mrange, mspacing, mratio)
        for key, value in kwargs.items():  # potential.py:902-906: any extra kwarg = topography
            self.topocarve = True
            self.mask = mesh.carvetopo(value[0], value[1], value[2])
        self.mesh = mesh
        self.meshrho, self.meshmag = mesh.copy(), mesh.copy()
        self.meshrho.addprop("density", np.zeros(mesh.size))
        self.meshmag.addprop("magnetization", ang2vec(np.zeros(mesh.size), self.inc, self.dec))
        table = mesh.bounds_table()
        n = len(self.lonobs)
        out = prism.assemble_grid(self.lonobs, self.latobs, self.heightobs, mesh)
        Kgz, Mc = out if out is not None else prism.assemble(self.lonobs, self.latobs, self.heightobs, table)
        Ktf, _ = prism.assemble_field("tf", self.lonobs, self.latobs, self.heightobs, table,
                                      vec=dircos(self.inc, self.dec))
        torch = _lib.require_cuda()
        self.Mcells = Mc
        self.M, self.n_total = 2 * Mc, 2 * n
        self.ld = _lib.padded_ld(self.M)
        self.rank, self.world, self.rows = 0, 1, (0, self.n_total)
        A = torch.zeros((self.n_total, self.ld), dtype=torch.float64, device=Kgz.device)
        A[:n, :Mc] = Kgz[:, :Mc]            # potential.py:938-941 np.block
        A[n:, Mc:2 * Mc] = Ktf[:, :Mc]
        self.kernel_gz, self.kernel_tf = Kgz[:, :Mc], Ktf[:, :Mc]
        self.mshape = mesh.shape
        self.mxs, self.mys, self.mzs = mesh.get_xs(), mesh.get_ys(), mesh.get_zs()
        self.Aw_pad = A
        self.weightKDM()
        self._engine = None
        self._mg_cache = {}
        if wavelet == "1D":
            from ..gravmag import compressor1D as cp1D

            self._say("Using {} wavelet to compress kernel.".format(wavelet))
            self.Awcp = cp1D.kernelcompressor(self.Aw)
        elif wavelet == "3D":
            # potential.py:951: cp3D reshapes every 2M-long row to (nz, ny, nx)
            raise ValueError("cannot reshape array of size {} into shape {}".format(self.M, tuple(self.mshape)))

    @property
    def A(self):
        """the un-weighted block kernel (potential.py:942), rebuilt on demand: A = Wb^-1 Aw Wm"""
        n, Mc = self.n_total // 2, self.Mcells
        A = self.Aw_pad[:, : self.M] * self.wm_dev[: self.M]
        A[n:] /= self._wb_ratio
        return A

    def weightKDM(self):
        """potential.py:1003-1067, in place on the device"""
        torch = _lib.require_cuda()
        L, s = _lib.lib(), _lib.stream_ptr()
        A = self.Aw_pad
        n2, ld = (int(v) for v in A.shape)
        n = n2 // 2
        f64 = dict(dtype=torch.float64, device=A.device)
        sumsq = torch.zeros(ld, **f64)
        _lib.check(L.gi_colsumsq(_lib.ptr(A), n2, self.M, ld, _lib.ptr(sumsq), 0, s), "gi_colsumsq")
        wm, wminv, wmsq = (torch.zeros(ld, **f64) for _ in range(3))
        _lib.check(L.gi_weights_from_sumsq(_lib.ptr(sumsq), self.M, 0.5, _lib.ptr(wm), _lib.ptr(wminv),
                                           _lib.ptr(wmsq), s), "gi_weights_from_sumsq")
        # Wb (potential.py:1041-1052, "method3"): the population standard deviations of the two kernels
        std_gz = float(self.kernel_gz.std(unbiased=False))
        std_tf = float(self.kernel_tf.std(unbiased=False))
        self._wb_ratio = std_gz / std_tf
        wb = np.append(np.ones_like(self.dobs_gz), np.ones_like(self.dobs_tf) * (std_gz / std_tf))
        A[n:] *= self._wb_ratio            # Aw = (Wb A) WmInv: rows first, then columns
        _lib.check(L.gi_scale_columns(_lib.ptr(A), n2, self.M, ld, _lib.ptr(wminv), s), "gi_scale_columns")
        _lib.sync()
        self.wm_dev, self.wminv_dev, self.wmsq_dev = wm, wminv, wmsq
        self.Aw = A[:, : self.M]
        row = np.arange(0, self.M)
        diag = lambda v, m: coo_matrix((v, (np.arange(m), np.arange(m))), shape=(m, m)).tocsr()
        self.Wm = diag(wm[: self.M].cpu().numpy(), self.M)
        self.WmInv = diag(wminv[: self.M].cpu().numpy(), self.M)
        self.WmSquare = diag(wmsq[: self.M].cpu().numpy(), self.M)
        self.Wb = diag(wb, n2)
        self.dobs = np.append(self.dobs_gz, self.dobs_tf)
        self.dobsw = self.Wb @ self.dobs
        self.dobs_sampler = self.dobsw     # what the data term compares with (potential.py:1677)

    def forward(self, model):
        """potential.py:1069-1074: un-weighted kernel times un-weighted model"""
        torch = _lib.require_cuda()
        m = torch.as_tensor(np.asarray(model, dtype=np.float64), device=self.Aw_pad.device)
        return (self.A @ m).cpu().numpy()

    def engine(self):
        if self._engine is None:
            # no mean removal: dobs_mean = 0 and n_total = 0 switch the centring off (gi_residual)
            self._engine = BlockEngine(self.Aw_pad, self.M, self.dobsw, 0.0, 0, None, None)
        return self._engine

    def misfit_and_grad(self, x, mwapr, low, high, constraint, log_fator, alpha, regulization="Damping",
                        beta=0.01):
        """potential.py:1775-1812"""
        x = np.asarray(x, dtype=np.float64)
        if constraint == "logarithmic":
            mw = (low + high * np.e ** (log_fator * x)) / (1 + np.e ** (log_fator * x))
        elif constraint == "mandatory":
            mw = x
        else:
            raise ValueError("Please choose right boundary constraint(mandatory, logarithmic)!")
        if regulization in ("Smoothness", "TV"):
            raise AttributeError("'JointModule' object has no attribute 'fd3d'")  # potential.py:1753,1765
        if regulization not in ("MS", "MS1", "MStry", "Damping"):
            raise ValueError("Please choose regularization from 'MS','Damping', 'Smoothness', 'TV'.")
        wmsq = self.wmsq_dev
        if regulization == "MStry":
            wmsq = self.wmsq_dev.clone()
            wmsq[: self.M] = 1.0
        return self._evaluate(mw, mwapr, alpha, self.extra_regs.get(regulization, regulization), beta,
                              wmsq=wmsq)
