"""TEST INFRASTRUCTURE ONLY -- CPU oracle (numpy + small C library) for the GravInv3DHMC
inversion hot path.

This is a *restatement* of the reference's algorithm used solely as the checker for the
CUDA implementation.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package
(``gravinv3dhmc_b200``) never does, and has no CPU fallback.

Pinning status
--------------
* prism gz, tesseroid gz, mesh bounds/masks, sensitivity weighting, regularisers,
  ``misfit_and_grad`` and the leapfrog/Metropolis loop are pinned against the UNMODIFIED
  reference executed in the build container (``oracle/ref_harness.py``); the outputs are
  committed under ``tests/golden/`` by ``oracle/make_golden.py`` and re-checked by
  ``tests/test_oracle_pinning.py`` (which also compares live when ``/root/reference`` is
  mounted).
* wavelet compressors (``compressor1D/3D``): **parity unpinned**.  The arithmetic lives in
  the third-party PyWavelets package (``pywt``; the reference pins no version and it is not
  installed here).  ``dwt_*`` below restates pywt's published db4 /
  ``periodization`` / ``coeffs_to_array`` conventions; only self-consistency
  (orthonormality, threshold-0 equivalence with the dense product) can be checked.

Each function cites the reference file:line it follows (paths relative to the reference root).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# constants.py:29,33,34,44
SI2MGAL = 100000.0
G_SPHERICAL = 0.00000000006673
G = 0.00000006673
MEAN_EARTH_RADIUS = 6378137.0
RATIO_G = 1.6  # gravmag/tesseroid.py:77


# --------------------------------------------------------------------------------------
# C library
# --------------------------------------------------------------------------------------
def build(force: bool = False) -> str:
    so = os.path.join(HERE, "_build", "liboracle.so")
    srcs = [os.path.join(HERE, "csrc", f) for f in ("oracle_prism.c", "oracle_tess.c", "oracle_fields.c")]
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared"] + srcs
                              + ["-o", so, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        dp = ctypes.POINTER(ctypes.c_double)
        i64 = ctypes.c_int64
        L.oracle_prism_gz.argtypes = [dp, dp, dp, i64, dp, i64, ctypes.c_double, dp, i64, dp, dp]
        L.oracle_prism_gz.restype = None
        L.oracle_tess_gz.argtypes = [dp, dp, dp, dp, i64, dp, i64, ctypes.c_double,
                                     ctypes.c_double, ctypes.c_double, dp, i64,
                                     ctypes.POINTER(ctypes.c_int), ctypes.POINTER(i64)]
        L.oracle_tess_gz.restype = ctypes.c_int
        L.oracle_tess_leaves.argtypes = [dp, dp, dp, dp, i64, dp, i64, ctypes.c_double,
                                         ctypes.c_void_p, i64]
        L.oracle_tess_leaves.restype = ctypes.c_int
        L.oracle_tess_field.argtypes = [ctypes.c_int, dp, dp, dp, dp, i64, dp, i64, ctypes.c_double,
                                        ctypes.c_double, ctypes.c_double, dp, i64, dp, dp,
                                        ctypes.POINTER(ctypes.c_int)]
        L.oracle_tess_field.restype = ctypes.c_int
        L.oracle_prism_field.argtypes = [ctypes.c_int, dp, dp, dp, i64, dp, i64, ctypes.c_double, dp,
                                         dp, i64, dp, ctypes.c_int, dp]
        L.oracle_prism_field.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def prism_gz(xp, yp, zp, bounds, dens=None, threads: int = 1):
    """(result, kernel2d) like gravmag/prism.py:911-918 for an explicit bounds table.

    ``bounds`` is [M,6] = x1,x2,y1,y2,z1,z2 of the ACTIVE prisms in mesh order
    (prism.py:299-312).  ``threads`` > 1 splits observation rows like prism.py:986-996.
    """
    xp, yp, zp, bounds = _c(xp), _c(yp), _c(zp), _c(bounds).reshape(-1, 6)
    if xp.shape != yp.shape or xp.shape != zp.shape:
        raise ValueError("Input arrays xp, yp, and zp must have same length!")  # prism.py:295
    N, M = xp.shape[0], bounds.shape[0]
    K = np.zeros((N, M))
    res = np.zeros(N)
    d = None if dens is None else _c(dens)
    scale = G * SI2MGAL  # prism.py:314-315
    L = _lib()

    def run(lo, hi):
        if hi <= lo:
            return
        L.oracle_prism_gz(_dp(xp[lo:hi]), _dp(yp[lo:hi]), _dp(zp[lo:hi]), hi - lo, _dp(bounds), M,
                          scale, _dp(K[lo:hi]), M, None if d is None else _dp(d),
                          None if d is None else _dp(res[lo:hi]))

    _run_rows(run, N, threads)
    return res, K


# constants.py:26,37,41,50
SI2EOTVOS = 1000000000.0
CM = 10. ** (-7)
T2NT = 10. ** (6)
G0 = 9.80
PRISM_FIELDS = {"potential": 0, "geoid": 0, "gx": 1, "gy": 2, "gz": 3, "gxx": 4, "gxy": 5, "gxz": 6,
                "gyy": 7, "gyz": 8, "gzz": 9, "tf": 10, "bx": 11, "by": 12, "bz": 13}


def prism_field_scale(field):
    """the factor gravmag/prism.py applies after the corner sums (:150, 178, 231, 367, 729, 777)"""
    if field == "potential":
        return G
    if field == "geoid":
        return G / G0
    if field in ("gx", "gy", "gz"):
        return G * SI2MGAL
    if field in ("tf", "bx", "by", "bz"):
        return CM * T2NT
    return G * SI2EOTVOS


def dircos(inc, dec):
    """utils.py:448-474"""
    d2r = np.pi / 180.
    return [np.cos(d2r * inc) * np.cos(d2r * dec), np.cos(d2r * inc) * np.sin(d2r * dec),
            np.sin(d2r * inc)]


def prism_field(field, xp, yp, zp, bounds, dens=None, inc=None, dec=None, mag=None, threads: int = 1):
    """(result, kernel2d) of gravmag/prism.py's `potential, geoid, gx .. gzz, tf` (:875-982) and the
    result of `bx, by, bz` (:735-870, kernel2d = None) for an explicit bounds table.  ``dens``: per
    prism densities; ``mag``: per-prism magnetisation vectors [M,3] (tf, bx, by, bz)."""
    xp, yp, zp, bounds = _c(xp), _c(yp), _c(zp), _c(bounds).reshape(-1, 6)
    if xp.shape != yp.shape or xp.shape != zp.shape:
        raise ValueError("Input arrays xp, yp, and zp must have same length!")
    N, M = xp.shape[0], bounds.shape[0]
    code = PRISM_FIELDS[field]
    want_kernel = code <= 10
    K = np.zeros((N, M)) if want_kernel else None
    res = np.zeros(N)
    vec = _c(dircos(inc, dec)) if field == "tf" else _c([0.0, 0.0, 0.0])
    if code >= 10:
        w = None if mag is None else _c(mag).reshape(M, 3)
        nw = 3
    else:
        w = None if dens is None else _c(dens)
        nw = 1
    L = _lib()

    def run(lo, hi):
        if hi <= lo:
            return
        L.oracle_prism_field(code, _dp(xp[lo:hi]), _dp(yp[lo:hi]), _dp(zp[lo:hi]), hi - lo, _dp(bounds),
                             M, prism_field_scale(field), _dp(vec), None if K is None else _dp(K[lo:hi]),
                             M, None if w is None else _dp(w), nw, None if w is None else _dp(res[lo:hi]))

    _run_rows(run, N, threads)
    return res, K


def convert_coords(lon, lat, height):
    """gravmag/tesseroid.py:109-123"""
    lon = np.radians(lon)
    lat = np.radians(lat)
    return lon, np.sin(lat), np.cos(lat), MEAN_EARTH_RADIUS + height


def tess_gz(lon, lat, height, bounds, ratio=RATIO_G, threads: int = 1, stats=None):
    """kernel2d like gravmag/tesseroid.py:421-431 for an explicit bounds table [M,6] =
    w,e,s,n,top,bottom (degenerate cells already dropped as in tesseroid.py:126-153).
    Raises OverflowError like _tesseroid_numba.py:53-54; returns (kernel2d, error_code)."""
    lon, lat, height = _c(lon), _c(lat), _c(height)
    assert lon.shape == lat.shape == height.shape, "Input coordinate arrays must have same shape"
    assert ratio > 0
    bounds = _c(bounds).reshape(-1, 6)
    lonr, sinlat, coslat, radius = (_c(a) for a in convert_coords(lon, lat, height))
    N, M = lon.shape[0], bounds.shape[0]
    K = np.zeros((N, M))
    L = _lib()
    errs, ovfs = [], []

    def run(lo, hi):
        if hi <= lo:
            return
        ovf = ctypes.c_int(0)
        st = (ctypes.c_int64 * 2)(0, 0)
        e = L.oracle_tess_gz(_dp(lonr[lo:hi]), _dp(sinlat[lo:hi]), _dp(coslat[lo:hi]),
                             _dp(radius[lo:hi]), hi - lo, _dp(bounds), M, ratio, SI2MGAL, G,
                             _dp(K[lo:hi]), M, ctypes.byref(ovf), st)
        errs.append(e)
        ovfs.append(ovf.value)
        if stats is not None:
            stats.append((st[0], st[1]))

    _run_rows(run, N, threads)
    if any(ovfs):
        raise OverflowError
    return K, int(sum(errs))


TESS_FIELDS = {"potential": 0, "geoid": 0, "gx": 1, "gy": 2, "gz": 3, "gxx": 4, "gxy": 5, "gxz": 6,
               "gyy": 7, "gyz": 8, "gzz": 9}
RATIO_V, RATIO_GG = 1, 8  # gravmag/tesseroid.py:76,78
G_SPHERICAL_S = 0.00000000006673  # constants.py:33 "Gs"


def tess_field_scales(field, forward=False):
    """(ratio, scale1, scale2) of gravmag/tesseroid.py:324-510 -- `kernel2d*SI2MGAL*G` is two
    multiplications, `kernel2d *= G` one; gy is scaled with Gs (tesseroid.py:416-417) -- and of
    gravmag/tesseroidforward.py:236-788 (`result *= SI2MGAL*G`: one multiplication by the product,
    gy with G)."""
    if field in ("potential", "geoid"):
        return RATIO_V, (G if field == "potential" else G / 9.80), 1.0
    unit, ratio = (SI2MGAL, RATIO_G) if field in ("gx", "gy", "gz") else (SI2EOTVOS, RATIO_GG)
    if forward:
        return ratio, unit * G, 1.0
    return ratio, unit, (G_SPHERICAL_S if field == "gy" else G)


def tess_field(field, lon, lat, height, bounds, dens=None, ratio=None, forward=False, threads: int = 1):
    """(result, kernel2d, error_code) of gravmag/tesseroid.py's `potential, geoid, gx .. gzz`
    (:324-510), or with forward=True the result of gravmag/tesseroidforward.py's functions
    (kernel2d = None), for an explicit bounds table."""
    lon, lat, height = _c(lon), _c(lat), _c(height)
    assert lon.shape == lat.shape == height.shape, "Input coordinate arrays must have same shape"
    r0, s1, s2 = tess_field_scales(field, forward)
    ratio = r0 if ratio is None else ratio
    assert ratio > 0
    bounds = _c(bounds).reshape(-1, 6)
    lonr, sinlat, coslat, radius = (_c(a) for a in convert_coords(lon, lat, height))
    N, M = lon.shape[0], bounds.shape[0]
    K = None if forward else np.zeros((N, M))
    res = np.zeros(N)
    d = None if dens is None else _c(dens)
    L = _lib()
    errs, ovfs = [], []

    def run(lo, hi):
        if hi <= lo:
            return
        ovf = ctypes.c_int(0)
        e = L.oracle_tess_field(TESS_FIELDS[field], _dp(lonr[lo:hi]), _dp(sinlat[lo:hi]),
                                _dp(coslat[lo:hi]), _dp(radius[lo:hi]), hi - lo, _dp(bounds), M, ratio,
                                s1, s2, None if K is None else _dp(K[lo:hi]), M,
                                None if d is None else _dp(d), None if d is None else _dp(res[lo:hi]),
                                ctypes.byref(ovf))
        errs.append(e)
        ovfs.append(ovf.value)

    _run_rows(run, N, threads)
    if any(ovfs):
        raise OverflowError
    return res, K, int(sum(errs))


def tess_leaves(lon, lat, height, bounds, ratio=RATIO_G, threads: int = 1):
    """per-pair leaf counts of the adaptive subdivision (_tesseroid_numba.py:32-71): the index
    bookkeeping of the tesseroid kernel, int32 [N, M] (-1 = stack overflow)."""
    lon, lat, height = _c(lon), _c(lat), _c(height)
    bounds = _c(bounds).reshape(-1, 6)
    lonr, sinlat, coslat, radius = (_c(a) for a in convert_coords(lon, lat, height))
    N, M = lon.shape[0], bounds.shape[0]
    out = np.zeros((N, M), dtype=np.int32)
    L = _lib()

    def run(lo, hi):
        if hi > lo:
            L.oracle_tess_leaves(_dp(lonr[lo:hi]), _dp(sinlat[lo:hi]), _dp(coslat[lo:hi]),
                                 _dp(radius[lo:hi]), hi - lo, _dp(bounds), M, ratio,
                                 out[lo:hi].ctypes.data_as(ctypes.c_void_p), M)

    _run_rows(run, N, threads)
    return out


_QLIB = None


def build_quad(force: bool = False) -> str:
    """liboracle_quad.so: the binary128 (libquadmath) leaf evaluation of oracle_tess_quad.c"""
    so = os.path.join(HERE, "_build", "liboracle_quad.so")
    src = os.path.join(HERE, "csrc", "oracle_tess_quad.c")
    if force or not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", so,
                               "-lquadmath", "-lm"])
    return so


def tess_gz_pairs_quad(lon, lat, height, bounds, obs_idx, cell_idx, ratio=RATIO_G, threads: int = 1):
    """(values, leaves) for the listed (observation, cell) pairs: the reference's subdivision (FP64
    decisions, identical leaves) with every leaf's GLQ sum and the accumulation in binary128, rounded
    to FP64 at the end and scaled like kernel2d (tesseroid.py:430) -- the "exact" value of the same
    quadrature, against which FP64 implementations are compared on the near field."""
    global _QLIB
    if _QLIB is None:
        _QLIB = ctypes.CDLL(build_quad())
        dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)
        _QLIB.oracle_tess_gz_pairs_quad.argtypes = [dp, dp, dp, dp, ip, dp, ip, ctypes.c_int64,
                                                    ctypes.c_double, ctypes.c_double, ctypes.c_double, dp,
                                                    ctypes.c_void_p]
        _QLIB.oracle_tess_gz_pairs_quad.restype = None
    bounds = _c(bounds).reshape(-1, 6)
    lonr, sinlat, coslat, radius = (_c(a) for a in convert_coords(_c(lon), _c(lat), _c(height)))
    obs_idx = np.ascontiguousarray(obs_idx, dtype=np.int64)
    cell_idx = np.ascontiguousarray(cell_idx, dtype=np.int64)
    n = obs_idx.size
    out, leaves = np.zeros(n), np.zeros(n, dtype=np.int32)
    ip = ctypes.POINTER(ctypes.c_int64)

    def run(lo, hi):
        if hi > lo:
            _QLIB.oracle_tess_gz_pairs_quad(_dp(lonr), _dp(sinlat), _dp(coslat), _dp(radius),
                                            obs_idx[lo:hi].ctypes.data_as(ip), _dp(bounds),
                                            cell_idx[lo:hi].ctypes.data_as(ip), hi - lo, ratio, SI2MGAL, G,
                                            _dp(out[lo:hi]), leaves[lo:hi].ctypes.data_as(ctypes.c_void_p))

    _run_rows(run, n, threads)
    return out, leaves


def _run_rows(fn, N, threads):
    if threads <= 1:
        fn(0, N)
        return
    from concurrent.futures import ThreadPoolExecutor

    n = max(1, N // threads)
    cuts = [(i * n, (i + 1) * n) for i in range(threads - 1)]
    cuts.append((cuts[-1][1] if cuts else 0, N))
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda c: fn(*c), cuts))


# --------------------------------------------------------------------------------------
# mesh bookkeeping (index -> bounds, masks)   mesher/mesh.py
# --------------------------------------------------------------------------------------
class OracleMesh:
    """Restatement of PrismMesh / TesseroidMesh / *Segment index arithmetic, one cell at a
    time (slow, literal).  mesher/mesh.py:166-223, 229-270, 601-645, 651-686."""

    def __init__(self, bounds, spacing, ratio=1, divisionsection=None, zdown=True):
        x1, x2, y1, y2, z1, z2 = bounds
        self.segmented = divisionsection is not None
        self.zdown = zdown
        self.mask = []
        if not self.segmented:
            dz, dy, dx = spacing
            self.dims = (dx, dy, dz)
            self.ratio = ratio
            nx = int(np.ceil((x2 - x1) / dx))
            ny = int(np.ceil((y2 - y1) / dy))
            if ratio == 1:
                nz = int(np.ceil((z2 - z1) / dz))
                bounds_big = x1, x1 + nx * dx, y1, y1 + ny * dy, z1, z1 + nz * dz
            else:  # mesh.py:181-198
                z_SubNum = 1
                while True:
                    z_SubDepth = z1 + dz * (1 - ratio ** z_SubNum) / (1 - ratio)
                    if z_SubDepth < z2 and (z2 - z_SubDepth) > dz:
                        z_SubNum += 1
                    else:
                        break
                nz = int(z_SubNum)
                bounds_big = x1, x1 + nx * dx, y1, y1 + ny * dy, z1, z2
        else:  # mesh.py:601-633
            dzlist, dy, dx = spacing
            self.dims = (dx, dy, dzlist)
            self.segment = len(dzlist)
            self.divisionsection = divisionsection
            nx = int(np.ceil((x2 - x1) / dx))
            ny = int(np.ceil((y2 - y1) / dy))
            nz = 0
            nzlist = np.zeros(self.segment)
            nzsumlist = np.zeros(self.segment)
            for i in range(self.segment):
                nzlist[i] = int(np.ceil((divisionsection[i + 1] - divisionsection[i]) / dzlist[i]))
                nz = nz + nzlist[i]
                nzsumlist[i] = nz
            self.nzlist, self.nzsumlist = nzlist, nzsumlist
            bounds_big = (x1, x1 + nx * dx, y1, y1 + ny * dy, z1,
                          divisionsection[-2] + nzlist[-1] * dzlist[-1])
        self.bounds = bounds_big
        self.shape = tuple(int(i) for i in (nz, ny, nx))
        self.size = int(nx * ny * nz)

    def cell(self, index):
        """bounds of cell ``index`` or None if masked (mesh.py:229-270 / 651-686)."""
        if index in self.mask:
            return None
        nz, ny, nx = self.shape
        k = index // (nx * ny)
        j = (index - k * (nx * ny)) // nx
        i = (index - k * (nx * ny) - j * nx)
        x1 = self.bounds[0] + self.dims[0] * i
        x2 = x1 + self.dims[0]
        y1 = self.bounds[2] + self.dims[1] * j
        y2 = y1 + self.dims[1]
        if not self.segmented:
            if self.ratio == 1:
                z1 = self.bounds[4] + self.dims[2] * k
                z2 = z1 + self.dims[2] if k < nz - 1 else self.bounds[5]
            else:
                z2 = self.bounds[4] + self.dims[2] * (1 - self.ratio ** (k + 1)) / (1 - self.ratio)
                z1 = z2 - self.dims[2] * self.ratio ** k
                if k == nz - 1:
                    z2 = self.bounds[5]
        else:
            for iseg in range(self.segment):
                if k < self.nzsumlist[iseg]:
                    kloc = iseg
                    break
            if kloc == 0:
                z1 = self.bounds[4] + self.dims[2][kloc] * k
                z2 = z1 + self.dims[2][kloc]
            else:
                z1 = self.divisionsection[kloc] + self.dims[2][kloc] * (k - self.nzsumlist[kloc - 1])
                z2 = z1 + self.dims[2][kloc]
        return tuple(float(v) for v in (x1, x2, y1, y2, z1, z2))

    def active_bounds(self):
        """[M,6] table of the non-masked cells in mesh order, and their flat indices."""
        mask = set(self.mask)
        idx = [i for i in range(self.size) if i not in mask]
        save, self.mask = self.mask, []
        tab = np.array([self.cell(i) for i in idx], dtype=np.float64).reshape(-1, 6)
        self.mask = save
        return tab, np.asarray(idx, dtype=np.int64)

    # mesh.py:396-445 / 799-841
    def get_xs(self):
        x1, x2 = self.bounds[0], self.bounds[1]
        dx = self.dims[0]
        xs = np.arange(x1, x2 + dx, dx)
        return xs[:-1] if xs.size > self.shape[2] + 1 else xs

    def get_ys(self):
        y1, y2 = self.bounds[2], self.bounds[3]
        dy = self.dims[1]
        ys = np.arange(y1, y2 + dy, dy)
        return ys[:-1] if ys.size > self.shape[1] + 1 else ys

    def get_zs(self):
        z1, z2 = self.bounds[4], self.bounds[5]
        nz = self.shape[0]
        if self.segmented:
            zs = []
            for iseg in range(self.segment):
                zs.extend(list(np.arange(self.divisionsection[iseg], self.divisionsection[iseg + 1],
                                         self.dims[2][iseg])))
            zs.append(z2)
            zs = np.array(zs)
        elif self.ratio == 1:
            zs = np.arange(z1, z2 + self.dims[2], self.dims[2])
        else:
            zs = np.zeros(nz + 1)
            for k in range(nz):
                bottom = self.bounds[4] + self.dims[2] * (1 - self.ratio ** (k + 1)) / (1 - self.ratio)
                zs[k] = bottom - self.dims[2] * self.ratio ** k
            zs[nz] = z2
        return zs[:-1] if zs.size > nz + 1 else zs

    def carvetopo(self, x, y, height, below=False):
        """mesher/mesh.py:301-394 (centres + cubic) and :717-797 (tops + nearest).
        Unlike the reference it does not write carve_topo_interp.txt into the CWD."""
        import scipy.interpolate

        nz, ny, nx = self.shape
        x1, x2, y1, y2, z1, z2 = self.bounds
        dx, dy, dz = self.dims
        xc = np.arange(x1, x2, dx) + 0.5 * dx
        if len(xc) > nx:
            xc = xc[:-1]
        yc = np.arange(y1, y2, dy) + 0.5 * dy
        if len(yc) > ny:
            yc = yc[:-1]
        if self.segmented:
            zc = []
            for iseg in range(self.segment):
                zc.extend(list(np.arange(self.divisionsection[iseg], self.divisionsection[iseg + 1],
                                         dz[iseg])))
            zc = np.array(zc)
            method = "nearest"
        else:
            if self.ratio == 1:
                zc = np.arange(z1, z2, dz) + 0.5 * dz
            else:
                zc = np.zeros(nz)
                for k in range(0, nz - 1):
                    bottom = self.bounds[4] + self.dims[2] * (1 - self.ratio ** (k + 1)) / (1 - self.ratio)
                    zc[k] = bottom - 0.5 * self.dims[2] * self.ratio ** k
                zc[nz - 1] = bottom + 0.5 * (z2 - bottom)
            method = "cubic"
        if len(zc) > nz:
            zc = zc[:-1]
        XC, YC = np.meshgrid(xc, yc)
        topo = scipy.interpolate.griddata((x, y), height, (XC, YC), method=method).ravel()
        if self.zdown:
            topo = -1 * topo
        topo_mask = topo.mask if np.ma.isMA(topo) else [False] * len(topo)
        c = 0
        for cellz in zc:
            for h, masked in zip(topo, topo_mask):
                if below:
                    if masked or (cellz > h and self.zdown) or (cellz < h and not self.zdown):
                        self.mask.append(c)
                else:
                    if masked or (cellz < h and self.zdown) or (cellz > h and not self.zdown):
                        self.mask.append(c)
                c += 1
        return self.mask


def check_tesseroids(bounds):
    """gravmag/tesseroid.py:126-153: assert validity, drop (with a warning) the degenerate cells.
    Returns a boolean keep-mask."""
    b = np.asarray(bounds, dtype=np.float64).reshape(-1, 6)
    w, e, s, n, top, bottom = b.T
    assert np.all((w <= e) & (s <= n) & (top >= bottom)), "Invalid tesseroid dimensions"
    return ~((e - w <= 1e-6) | (n - s <= 1e-6) | (top - bottom <= 1e-3))


def rho2carve(rho, mask):
    """utils.py:714-727"""
    m = set(int(i) for i in mask)
    return np.array([rho[i] for i in range(rho.shape[0]) if i not in m])


def carve2rho(rhocarve, rho, mask):
    """utils.py:729-749 (mutates and returns a copy, like the reference)."""
    m = set(int(i) for i in mask)
    j = 0
    for i in range(rho.shape[0]):
        if i not in m:
            rho[i] = rhocarve[j]
            j += 1
    return rho.copy()


# --------------------------------------------------------------------------------------
# potential energy   inversion/potential.py
# --------------------------------------------------------------------------------------
def sensitivity_weighting(A, weightfactor=0.5):
    """inversion/potential.py:232-264 -> (Aw, wm, wminv, wmsq) with the diagonals as vectors.

    Sequential sum over observations j for each column (potential.py:241-244) == np.add.reduce
    along axis 0 in row order for a C-ordered array is pairwise, so do the literal loop over
    rows (vectorised over columns; same order of additions per column)."""
    A = np.asarray(A, dtype=np.float64)
    ADiagSquare = np.zeros(A.shape[1])
    for j in range(A.shape[0]):
        ADiagSquare += A[j, :] ** 2
    ADiag = np.power(ADiagSquare, weightfactor)
    # potential.py:247-251: the loop leaves ADiagInv = 1/ADiag unless the LAST entry is 0,
    # in which case it is the scalar 0 (coo_matrix would then fail; not reachable in practice).
    if abs(ADiag[-1]) == 0:
        raise ZeroDivisionError("reference would build WmInv from scalar 0")
    with np.errstate(divide="ignore"):
        ADiagInv = 1.0 / ADiag
    ADiagSquare = ADiag * ADiag
    Aw = A * ADiagInv[None, :]  # A @ diag(ADiagInv): one multiply per entry
    return Aw, ADiag, ADiagInv, ADiagSquare


def fd3d(shape):
    """inversion/potential.py:266-361, vectorised; identical CSR (checked against the
    reference's loop builder in tests/test_oracle_pinning.py)."""
    nz, ny, nx = shape
    per_layer = (nx - 1) * ny + (ny - 1) * nx
    nderivs = per_layer * nz + nx * ny * (nz - 1)
    idx = np.arange(nz * ny * nx).reshape(nz, ny, nx)
    rows, c0, c1 = [], [], []
    for k in range(nz):
        a = idx[k, :, :-1].ravel()
        rows.append(per_layer * k + np.arange(a.size))
        c0.append(a)
        c1.append(a + 1)
        b = idx[k, :-1, :].ravel()
        rows.append(per_layer * k + a.size + np.arange(b.size))
        c0.append(b)
        c1.append(b + nx)
    front = per_layer * nz
    for k in range(nz - 1):
        a = idx[k].ravel()
        rows.append(front + nx * ny * k + np.arange(a.size))
        c0.append(a)
        c1.append(a + nx * ny)
    rows = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
    c0 = np.concatenate(c0) if c0 else np.zeros(0, dtype=np.int64)
    c1 = np.concatenate(c1) if c1 else np.zeros(0, dtype=np.int64)
    I = np.concatenate([rows, rows])
    J = np.concatenate([c0, c1])
    V = np.concatenate([np.ones(rows.size), -np.ones(rows.size)])
    return sp.coo_matrix((V, (I, J)), (nderivs, nx * ny * nz)).tocsr()


class OracleModel:
    """The sampler-facing duck type of GravMagModule (potential.py:584-589, 688-845) built from
    an explicit weighted kernel.  ``wavelet`` in {False,'1D','3D'} uses the restated compressors."""

    def __init__(self, Aw, wm, dobs, mshape, fixed=False, grav_fix=None, wavelet=False):
        self.Aw = np.asarray(Aw, dtype=np.float64)
        self.wm = np.asarray(wm, dtype=np.float64)
        self.wminv = 1.0 / self.wm
        self.wmsq = self.wm * self.wm
        self.dobs = np.asarray(dobs, dtype=np.float64)
        self.mshape = tuple(mshape)
        self.fixed = fixed
        self.grav_fix = grav_fix
        self.wavelet = wavelet
        self._R3d = None
        if wavelet == "1D":
            self.Awcp = kernelcompressor_1d(self.Aw)
        elif wavelet == "3D":
            self.Awcp = kernelcompressor_3d(self.Aw, self.mshape)

    def R3d(self):
        # the reference rebuilds this on every call (potential.py:791,803); value-identical
        if self._R3d is None:
            self._R3d = fd3d(self.mshape)
        return self._R3d

    def data_all(self, mw):
        """potential.py:688-717"""
        if self.wavelet == "1D":
            dpre = modelcompressor_1d(mw, self.Awcp)
        elif self.wavelet == "3D":
            dpre = modelcompressor_3d(mw, self.Awcp, self.mshape)
        else:
            dpre = np.dot(self.Aw, mw)
        dinv = dpre + self.grav_fix if self.fixed else dpre
        r = (dinv - np.mean(dinv)) - (self.dobs - np.mean(self.dobs))
        data_value = np.linalg.norm(r) ** 2
        data_gradient = 2 * np.dot(self.Aw.T, r)
        return dpre, data_value, data_gradient

    def model_MS_all(self, mw, mwapr, beta):
        """potential.py:719-736"""
        mwSquare = (mw - mwapr) ** 2
        model_value = np.sum((self.wmsq * mwSquare) / (mwSquare + beta))
        model_gradient = (2 * beta * self.wmsq * (mw - mwapr)) / (mwSquare + beta) ** 2
        return model_value, model_gradient

    def model_Damping_all(self, mw, mwapr):
        """potential.py:775-784"""
        return np.dot((mw - mwapr).T, (mw - mwapr)), 2 * (mw - mwapr)

    def model_Smoothness_all(self, mw, mwapr):
        """potential.py:786-796"""
        R = self.R3d()
        t = R @ (mw - mwapr)
        return np.dot(t.T, t), 2 * R.T @ R @ (mw - mwapr)

    def model_TV_all(self, mw, mwapr, beta):
        """potential.py:798-810"""
        R = self.R3d()
        t1 = R @ (mw - mwapr)
        t2 = np.sqrt(t1 ** 2 + beta)
        return np.sum(t2), R.T @ (t1 / t2)

    def misfit_and_grad(self, x, mwapr, low, high, constraint, log_fator, alpha,
                        regulization="Damping", beta=0.01):
        """potential.py:812-845"""
        if constraint == "logarithmic":
            mw = (low + high * np.e ** (log_fator * x)) / (1 + np.e ** (log_fator * x))
        elif constraint == "mandatory":
            mw = x
        else:
            raise ValueError("Please choose right boundary constraint(mandatory, logarithmic)!")
        dpre, data_value, data_gradient = self.data_all(mw)
        if regulization == "MS":
            mv, mg = self.model_MS_all(mw, mwapr, beta)
        elif regulization == "Damping":
            mv, mg = self.model_Damping_all(mw, mwapr)
        elif regulization == "Smoothness":
            mv, mg = self.model_Smoothness_all(mw, mwapr)
        elif regulization == "TV":
            mv, mg = self.model_TV_all(mw, mwapr, beta)
        else:
            raise ValueError("Please choose regularization from 'MS','Damping', 'Smoothness', 'TV'.")
        return data_value + alpha * mv, data_gradient + alpha * mg, dpre, data_value, mv


# --------------------------------------------------------------------------------------
# sampler   inversion/hmc.py
# --------------------------------------------------------------------------------------
class DrawStream:
    """Replays the reference's legacy global-RNG call order (hmc.py:260,297,95,165):
    seed -> per proposal randint(Lmin, Lmax+1), randn(n)*Sigma, rand()."""

    def __init__(self, seed):
        self.rs = np.random.RandomState(seed)

    def next_L(self, Lrange):
        return int(self.rs.randint(Lrange[0], Lrange[1] + 1))

    def next_p(self, n):
        return self.rs.randn(n)

    def next_u(self):
        return float(self.rs.rand())


def leapfrog(model, xcur, dt, L, alpha, p0, u, mwapr, low, high, constraint="mandatory",
             log_factor=1000, regularization="Damping", beta=0.01, trace=None):
    """inversion/hmc.py:85-177 with the momentum draw ``p0`` (already multiplied by Sigma) and
    the uniform ``u`` injected.  ``trace`` (list) receives (x, U) after every gradient call.
    Returns (x, U, dsyn, accept, U_data, U_model, Hcur, Hnew)."""
    mg = lambda x: model.misfit_and_grad(x, mwapr, low, high, constraint, log_factor, alpha,
                                         regulization=regularization, beta=beta)
    pnew = p0 * 1.0
    xnew = xcur * 1.0
    K = np.dot(pnew, pnew) * 0.5  # hmc.py:44-50 with the identity inverse mass
    U, grad, dsyn, U_data, U_model = mg(xnew)
    if trace is not None:
        trace.append((xnew.copy(), U))
    Hcur = K + U
    dsyn_new, Unew, Unew_data, Unew_model = dsyn.copy(), U, U_data, U_model
    pnew -= dt * grad * 0.5
    for i in range(L):
        xnew += dt * pnew
        if constraint == "mandatory":  # hmc.py:121-144 (the while loop runs at most once)
            idx1 = xnew > high
            idx2 = xnew < low
            xnew[idx1] = high[idx1]
            pnew[idx1] = -pnew[idx1]
            xnew[idx2] = low[idx2]
            pnew[idx2] = -pnew[idx2]
        Unew, grad, dsyn_new, Unew_data, Unew_model = mg(xnew)
        if trace is not None:
            trace.append((xnew.copy(), Unew))
        if i < L - 1:
            pnew -= dt * grad
        else:
            pnew -= dt * grad * 0.5
    pnew = -pnew
    Knew = np.dot(pnew, pnew) * 0.5
    Hnew = Knew + Unew
    accept = False
    if Hnew < Hcur or u < np.exp(-(Hnew - Hcur)):
        xcur, U, dsyn, accept, U_data, U_model = xnew, Unew, dsyn_new, True, Unew_data, Unew_model
    return xcur, U, dsyn, accept, U_data, U_model, Hcur, Hnew


def hmc_sample(model, nsamples, ndraws, delta, Lrange, initial_model, aprior_model, boundaries,
               constraint, log_factor, RegulFactor, regularization, beta, seed, Sigma, myrank=0,
               max_proposals=None, trace=None):
    """inversion/hmc.py:358-403 + 252-343 without the file output: returns dict with the
    misfit rows (7 columns, hmc.py:310-316), the accepted models m = WmInv @ mw and the
    per-proposal (L, accept) log."""
    wm = model.wm
    low = wm * boundaries[:, 0]
    high = wm * boundaries[:, 1]
    mw0 = wm * initial_model
    mwapr = wm * aprior_model
    draws = DrawStream(seed + myrank)
    if constraint == "logarithmic":
        x = (1 / log_factor) * np.log((mw0 - low) / (high - mw0))
    elif constraint == "mandatory":
        x = mw0
    else:
        raise ValueError("Please choose right boundary constraint(mandatory, logarithmic)!")
    data_size, model_size = model.dobs.shape[0], mw0.shape[0]
    alpha = RegulFactor
    rows, models, log = [], [], []
    i = ncount = 0
    while i < ndraws + nsamples:
        if max_proposals is not None and ncount >= max_proposals:
            break
        L = draws.next_L(Lrange)
        p0 = draws.next_p(model_size) * Sigma
        tr = [] if trace is not None else None
        # the reference draws u AFTER the trajectory (hmc.py:165); RandomState order is preserved
        # because nothing else consumes the stream in between.
        u = draws.next_u()
        x, U, _, acc, Ud, Um, Hc, Hn = leapfrog(model, x, delta, L, alpha, p0, u, mwapr, low, high,
                                                constraint, log_factor, regularization, beta, tr)
        if trace is not None:
            trace.append(dict(L=L, accept=acc, steps=tr, Hcur=Hc, Hnew=Hn))
        log.append((L, bool(acc)))
        if acc:
            if i >= ndraws:
                Udn, Umn = Ud / data_size, Um / model_size
                rows.append([U, Ud, Um, Udn + alpha * Umn, Udn, Umn, alpha])
                if constraint == "logarithmic":
                    mw = (low + high * np.e ** (log_factor * x)) / (1 + np.e ** (log_factor * x))
                else:
                    mw = x
                models.append(model.wminv * mw)
            i += 1
        ncount += 1
    return dict(misfit=np.array(rows).reshape(-1, 7), models=np.array(models), log=log, x=x)


# --------------------------------------------------------------------------------------
# wavelet compressors   gravmag/compressor1D.py, compressor3D.py   (PARITY UNPINNED: pywt absent)
# --------------------------------------------------------------------------------------
# PyWavelets 'db4' decomposition low-pass filter (dec_lo); dec_hi is its quadrature mirror.
DB4_DEC_LO = np.array([-0.010597401784997278, 0.032883011666982945, 0.030841381835986965,
                       -0.18703481171888114, -0.02798376941698385, 0.6308807679295904,
                       0.7148465705525415, 0.23037781330885523])
DB4_DEC_HI = np.array([-0.23037781330885523, 0.7148465705525415, -0.6308807679295904,
                       -0.02798376941698385, 0.18703481171888114, 0.030841381835986965,
                       -0.032883011666982945, -0.010597401784997278])
WAVELET_LEVEL = 2      # compressor1D.py:24, compressor3D.py:24
WAVELET_THRESH = 0.001  # compressor1D.py:25, compressor3D.py:25


def dwt_per(x, axis=-1):
    """Single-level db4 DWT, mode='periodization' (pywt.dwt convention): odd lengths are
    extended by repeating the last sample, output length ceil(n/2),
    cA[o] = sum_j dec_lo[j] * xe[(2*o + F/2 - j) mod ne]  with F = 8."""
    x = np.moveaxis(np.asarray(x, dtype=np.float64), axis, -1)
    n = x.shape[-1]
    if n % 2:
        x = np.concatenate([x, x[..., -1:]], axis=-1)
    ne = x.shape[-1]
    no = ne // 2
    F = DB4_DEC_LO.size
    o = np.arange(no)
    cA = np.zeros(x.shape[:-1] + (no,))
    cD = np.zeros_like(cA)
    for j in range(F):
        src = (2 * o + F // 2 - j) % ne
        cA += DB4_DEC_LO[j] * x[..., src]
        cD += DB4_DEC_HI[j] * x[..., src]
    return np.moveaxis(cA, -1, axis), np.moveaxis(cD, -1, axis)


def wavedec_1d(x, level=WAVELET_LEVEL):
    """pywt.wavedec(..., 'db4', mode='periodization', level) -> [cA_n, cD_n, ..., cD_1]"""
    coeffs = []
    a = np.asarray(x, dtype=np.float64)
    for _ in range(level):
        a, d = dwt_per(a)
        coeffs.append(d)
    coeffs.append(a)
    return coeffs[::-1]


def coeffs_to_array_1d(coeffs):
    """pywt.coeffs_to_array for wavedec output: plain concatenation [cA, cD_n, ..., cD_1]."""
    return np.concatenate(coeffs)


def wavedecn_3d(x, level=WAVELET_LEVEL):
    """pywt.wavedecn: per level transform axes 0,1,2 in order; detail dict keyed
    'aad','ada','add','daa','dad','dda','ddd' (letter order = axis order)."""
    a = np.asarray(x, dtype=np.float64)
    out = []
    for _ in range(level):
        parts = {"": a}
        for ax in range(3):
            nxt = {}
            for key, v in parts.items():
                ca, cd = dwt_per(v, axis=ax)
                nxt[key + "a"] = ca
                nxt[key + "d"] = cd
            parts = nxt
        a = parts.pop("aaa")
        out.append(parts)
    return [a] + out[::-1]


def coeffs_to_array_3d(coeffs):
    """pywt.coeffs_to_array for wavedecn output: Mallat packing; the array grows by the detail
    shape at each level; 'd' on an axis selects the upper block starting at the current
    approximation size along that axis; gaps (non-nesting shapes) stay zero."""
    a0 = coeffs[0]
    shapes = [a0.shape]
    total = list(a0.shape)
    for d in coeffs[1:]:
        dshape = d["ddd"].shape
        total = [t + s for t, s in zip(total, dshape)]
    arr = np.zeros(total)
    arr[tuple(slice(0, s) for s in a0.shape)] = a0
    pos = list(a0.shape)
    for d in coeffs[1:]:
        dshape = d["ddd"].shape
        for key, v in d.items():
            sl = []
            for ax, ch in enumerate(key):
                if ch == "a":
                    sl.append(slice(0, v.shape[ax]))
                else:
                    sl.append(slice(pos[ax], pos[ax] + v.shape[ax]))
            arr[tuple(sl)] = v
        pos = [p + s for p, s in zip(pos, dshape)]
    return arr


def kernelcompressor_1d(Aw, thr=WAVELET_THRESH):
    """gravmag/compressor1D.py:17-42"""
    rows = []
    for irow in range(Aw.shape[0]):
        c = coeffs_to_array_1d(wavedec_1d(Aw[irow, :].copy()))
        c[np.abs(c) < thr] = 0
        rows.append(c)
    return sp.csr_matrix(np.array(rows).reshape(Aw.shape[0], -1))


def modelcompressor_1d(m, Awcp):
    """gravmag/compressor1D.py:45-60"""
    return Awcp @ coeffs_to_array_1d(wavedec_1d(m))


def kernelcompressor_3d(Aw, mshape, thr=WAVELET_THRESH):
    """gravmag/compressor3D.py:17-44"""
    CZ, CY, CX = mshape
    rows = []
    for irow in range(Aw.shape[0]):
        c = coeffs_to_array_3d(wavedecn_3d(Aw[irow, :].copy().reshape((CZ, CY, CX))))
        c[np.abs(c) < thr] = 0
        rows.append(c.reshape(1, -1))
    return sp.csr_matrix(np.array(rows).reshape(Aw.shape[0], -1))


def modelcompressor_3d(m, Awcp, mshape):
    """gravmag/compressor3D.py:47-68"""
    CZ, CY, CX = mshape
    c = coeffs_to_array_3d(wavedecn_3d(np.asarray(m).reshape((CZ, CY, CX))))
    return np.squeeze(Awcp @ c.reshape((-1, 1)))


# --------------------------------------------------------------------------------------
# regularised conjugate gradient + bootstrap   inversion/reginv.py  
# --------------------------------------------------------------------------------------
class OracleCG:
    """``ConjugateGradient`` of inversion/reginv.py:22-491 from an explicit weighted kernel.
    ``A`` is the unweighted kernel (reginv.py:489 forwards the result through it)."""

    def __init__(self, A, dobs, mshape, wavelet=False):
        # reginv.py:120-149 newkernel (the fixed np.sqrt instead of weightfactor)
        self.wavelet = wavelet
        self.A = np.asarray(A, dtype=np.float64)
        self.Aw, self.wm, self.wminv, self.wmsq = sensitivity_weighting(self.A, 0.5)
        self.wm = np.sqrt(self.wm * self.wm)  # reginv.py:129 (sqrt of the sum of squares)
        self.dobs = np.asarray(dobs, dtype=np.float64)
        self.mshape = tuple(mshape)
        self.dsize, self.msize = self.A.shape
        self._R = None
        if wavelet == "1D":  # reginv.py:107-117 (PyWavelets restated: parity unpinned)
            self.Awcp = kernelcompressor_1d(self.Aw)
        elif wavelet == "3D":
            self.Awcp = kernelcompressor_3d(self.Aw, self.mshape)

    def R3d(self):
        if self._R is None:
            self._R = fd3d(self.mshape)  # reginv.py:151-246 is the same builder as potential.py
        return self._R

    def _dpre(self, mw):  # reginv.py:250-255
        if self.wavelet == "1D":
            return modelcompressor_1d(mw, self.Awcp)
        if self.wavelet == "3D":
            return modelcompressor_3d(mw, self.Awcp, self.mshape)
        return np.dot(self.Aw, mw)

    def data(self, mw):  # reginv.py:248-257
        return np.linalg.norm(self._dpre(mw) - self.dobs) ** 2

    def data_gfun(self, mw):  # reginv.py:259-269
        return 2 * np.dot(self.Aw.T, (self._dpre(mw) - self.dobs))

    def model(self, reg, mw, mwapr, beta):
        if reg == "MS":  # reginv.py:271-281
            sq = (mw - mwapr) ** 2
            return np.sum((self.wmsq * sq) / (sq + beta))
        if reg == "Damping":  # reginv.py:295-302
            return np.dot((mw - mwapr).T, (mw - mwapr))
        if reg == "Smoothness":  # reginv.py:313-321
            t = self.R3d() @ (mw - mwapr)
            return np.dot(t.T, t)
        if reg == "TV":  # reginv.py:333-343
            t = self.R3d() @ (mw - mwapr)
            return np.sum(np.sqrt(t ** 2 + beta))
        raise ValueError("Please choose regularization from 'MS','Damping', 'Smoothness', 'TV'.")

    def model_gfun(self, reg, mw, mwapr, beta):
        if reg == "MS":  # reginv.py:283-293 -- the denominator uses mw*mw, NOT (mw - mwapr)^2
            return ((2 * beta * self.wmsq) * (mw - mwapr)) / (mw * mw + beta) ** 2
        if reg == "Damping":  # reginv.py:304-311
            return 2 * (mw - mwapr)
        if reg == "Smoothness":  # reginv.py:323-331
            R = self.R3d()
            return 2 * R.T @ R @ (mw - mwapr)
        if reg == "TV":  # reginv.py:345-355
            R = self.R3d()
            t1 = R @ (mw - mwapr)
            return R.T @ (t1 / np.sqrt(t1 ** 2 + beta))
        raise ValueError("Please choose regularization from 'MS','Damping', 'Smoothness', 'TV'.")

    def CG(self, initialModel, apriorModel, boundary, regularization="MS", beta=0.01, q=0.9,
           maxk=100):
        """reginv.py:357-491 -> model_inv, data_inv, data_misfit, model_misfit, regul_factor"""
        reg = regularization
        mw = self.wm * initialModel
        mwapr = self.wm * apriorModel
        rhomin, rhomax = boundary[0], boundary[1]
        data_misfit, model_misfit, regul_factor = [], [], []
        for k in range(0, maxk):
            if k == 0:
                alpha = 0
            elif k == 1:
                alpha = self.data(mw_new) / self.model(reg, mw_new, mwapr, beta)
            elif self.data(mw) - self.data(mw_new) < 0.01 * self.data(mw):
                alpha = q * alpha
            regul_factor.append(alpha)
            if k == 0:
                data_misfit.append(self.data(mw) / self.dsize)
                I = self.data_gfun(mw) + alpha * self.model_gfun(reg, mw, mwapr, beta)
                model_misfit.append(self.model(reg, mw, mwapr, beta) / self.msize)
                Iw = I
            else:
                I_old, Iw_old, mw = I, Iw, mw_new
                I = self.data_gfun(mw) + alpha * self.model_gfun(reg, mw, mwapr, beta)
                mu = np.linalg.norm(I) ** 2 / np.linalg.norm(I_old) ** 2
                Iw = I + mu * Iw_old
            kstep = np.dot(Iw.T, I) / (np.linalg.norm(self.Aw @ Iw) ** 2
                                       + alpha * np.linalg.norm(Iw) ** 2)
            mw_new = mw - kstep * Iw
            mtemp = self.wminv * mw_new
            mtemp[mtemp < rhomin] = rhomin
            mtemp[mtemp > rhomax] = rhomax
            mw_new = self.wm * mtemp
            if k > 0:
                data_misfit.append(self.data(mw_new) / self.dsize)
                model_misfit.append(self.model(reg, mw_new, mwapr, beta) / self.msize)
                if self.data(mw_new) / self.dsize < 0.001:
                    break
        model_inv = self.wminv * mw_new
        return model_inv, self.A @ model_inv, data_misfit, model_misfit, regul_factor


class OracleBootStrap(OracleCG):
    """``BootStrap`` of inversion/reginv.py:494-748 (MS regulariser only, no prior model)."""

    def __init__(self, A, dobs, mshape, boundary, samples=100, beta=0.01, maxk=100):
        super().__init__(A, dobs, mshape)
        self.boundary, self.samples, self.beta, self.maxk = boundary, samples, beta, maxk

    def model_MS(self, mw):  # reginv.py:599-606
        sq = mw * mw
        return np.sum((self.wmsq * sq) / (sq + self.beta ** 2))

    def model_gfun_MS(self, mw):  # reginv.py:620-629
        r2 = mw * mw + self.beta ** 2
        return ((2 * self.wmsq) * (mw * self.beta ** 2)) / (r2 * r2)

    @staticmethod
    def _data(mw, Aw, dobs):  # reginv.py:588-597
        return np.linalg.norm(np.dot(Aw, mw) - dobs) ** 2

    def CG(self, Aw, dobs, initialModel):
        """reginv.py:631-713 -> model_inv, data_misfit, model_misfit, regul_factor"""
        mw = self.wm * initialModel
        rhomin, rhomax = self.boundary[0], self.boundary[1]
        q = 0.9
        data_misfit, model_misfit, regul_factor = [], [], []
        for k in range(0, self.maxk):
            if k == 0:
                alpha = 0
            elif k == 1:
                alpha = self._data(mw_new, Aw, dobs) / self.model_MS(mw_new)
            elif self._data(mw, Aw, dobs) - self._data(mw_new, Aw, dobs) < 0.01 * self._data(mw, Aw, dobs):
                alpha = q * alpha
            regul_factor.append(alpha)
            if k == 0:
                I = 2 * np.dot(Aw.T, np.dot(Aw, mw) - dobs) + alpha * self.model_gfun_MS(mw)
                Iw = I
            else:
                I_old, Iw_old, mw = I, Iw, mw_new
                I = 2 * np.dot(Aw.T, np.dot(Aw, mw) - dobs) + alpha * self.model_gfun_MS(mw)
                mu = np.linalg.norm(I) ** 2 / np.linalg.norm(I_old) ** 2
                Iw = I + mu * Iw_old
            kstep = np.dot(Iw.T, I) / (np.linalg.norm(Aw @ Iw) ** 2 + alpha * np.linalg.norm(Iw) ** 2)
            mw_new = mw - kstep * Iw
            mtemp = self.wminv * mw_new
            mtemp[mtemp < rhomin] = rhomin
            mtemp[mtemp > rhomax] = rhomax
            mw_new = self.wm * mtemp
            if k > 0:
                if self._data(mw_new, Aw, dobs) < 0.1:
                    break
                data_misfit.append(self._data(mw_new, Aw, dobs) / self.dsize)
                model_misfit.append(self.model_MS(mw_new) / self.msize)
        return self.wminv * mw_new, data_misfit, model_misfit, regul_factor

    def BSCG(self, initialModel):
        """reginv.py:715-748: replicate s resamples the observation rows with the legacy global RNG
        seeded by s (``np.random.seed(s); np.random.choice``) and runs CG on the gathered rows."""
        model_inv_all = np.zeros((self.samples, self.msize))
        data_misfit_all = np.zeros((self.samples, self.maxk - 1))
        model_misfit_all = np.zeros((self.samples, self.maxk - 1))
        regul_factor_all = np.zeros((self.samples, self.maxk))
        for sample in range(self.samples):
            rs = np.random.RandomState(sample)
            idx = rs.choice(np.arange(0, self.dsize), size=self.dsize, replace=True, p=None)
            m, dm, mm, rf = self.CG(self.Aw[idx, :], self.dobs[idx], initialModel)
            model_inv_all[sample, :] = m
            data_misfit_all[sample, :] = dm  # ValueError when a replicate stopped early, as reginv.py:745
            model_misfit_all[sample, :] = mm
            regul_factor_all[sample, :] = rf
        return model_inv_all, data_misfit_all, model_misfit_all, regul_factor_all
