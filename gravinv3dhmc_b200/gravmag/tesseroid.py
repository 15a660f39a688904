"""Gravity effect gz of tesseroids (spherical prisms) and its sensitivity matrix on the GPU.

Mirror of the reference's gravmag/tesseroid.py `gz` (:421-431) -> `_dispatcher` (:156-186) ->
`_forward_model` (:189-232) -> numba `engine(kernelz)` (gravmag/_tesseroid_numba.py:25-72):
2x2x2 Gauss-Legendre quadrature with adaptive subdivision (distance/size ratio 1.6, LIFO stack of
100 cells).  Same arguments and `(result, kernel2d)` return; raises `AssertionError` for invalid
input, `OverflowError` when the subdivision stack overflows and emits the same `RuntimeWarning`s.

The other fields (`potential, geoid, gx, gy, gxx, gxy, gxz, gyy, gyz, gzz`; tesseroid.py:324-510) go
through `gi_tess_field_assemble` (csrc/tess_fields.cu): the same adaptive engine with the GLQ kernels
of _tesseroid_numba.py:160-328, the reference's distance-size ratios (RATIO_V = 1, RATIO_G = 1.6,
RATIO_GG = 8) and its scale factors (including gy's `Gs`, tesseroid.py:416-417).
"""
from __future__ import annotations

import warnings

import numpy as np

from .. import _lib
from ..constants import G, G_SI, MEAN_EARTH_RADIUS, SI2EOTVOS, SI2MGAL, g0
from ._common import matvec_padded, model_table, to_device

RATIO_V = 1       # gravmag/tesseroid.py:76
RATIO_G = 1.6     # gravmag/tesseroid.py:77
RATIO_GG = 8      # gravmag/tesseroid.py:78
STACK_SIZE = 100  # gravmag/tesseroid.py:79


def _convert_coords(lon, lat, height):
    """degrees -> radians, sin/cos of latitude, radius (gravmag/tesseroid.py:109-123)."""
    lon = np.radians(lon)
    lat = np.radians(lat)
    return lon, np.sin(lat), np.cos(lat), MEAN_EARTH_RADIUS + height


def _check_table(table):
    """tesseroid.py:126-153: validate, drop degenerate cells (with the reference's warning)."""
    if table.shape[0] == 0:
        return table, 0
    w, e, s, n, top, bottom = table.T
    bad = ~((w <= e) & (s <= n) & (top >= bottom))
    if bad.any():
        raise AssertionError("Invalid tesseroid dimensions {}".format(list(table[np.argmax(bad)])))
    tiny = (e - w <= 1e-6) | (n - s <= 1e-6) | (top - bottom <= 1e-3)
    ndrop = int(tiny.sum())
    if ndrop:
        warnings.warn("Encountered tesseroid with dimensions smaller than the numerical threshold "
                      "(1e-6 degrees or 1e-3 m). Ignoring this tesseroid.", RuntimeWarning)
        table = table[~tiny]
    return table, ndrop


FIELD_CODES = {"potential": 0, "geoid": 0, "gx": 1, "gy": 2, "gz": 3, "gxx": 4, "gxy": 5, "gxz": 6,
               "gyy": 7, "gyz": 8, "gzz": 9}


def field_scales(field, forward=False):
    """(default ratio, scale1, scale2): kernel = (raw * scale1) * scale2 -- tesseroid.py:375-507
    (`kernel2d*SI2MGAL*G` is two multiplications, `kernel2d *= G` one, gy uses Gs) and, with
    forward=True, tesseroidforward.py:287-787 (`result *= SI2MGAL*G`: one factor, gy with G)."""
    if field == "potential":
        return RATIO_V, G, 1.0
    if field == "geoid":
        return RATIO_V, G / g0, 1.0
    unit, ratio = (SI2MGAL, RATIO_G) if field in ("gx", "gy", "gz") else (SI2EOTVOS, RATIO_GG)
    if forward:
        return ratio, unit * G, 1.0
    return ratio, unit, (G_SI if field == "gy" else G)


def assemble(lon, lat, height, table, ratio=RATIO_G, rows=None, device=None, ncols=None, field=None,
             forward=False):
    """Device sensitivity matrix [nrows, ld] for an explicit, already validated [M,6] table
    (w, e, s, n, top, bottom).  `ncols` >= M reserves trailing zero columns.  `field` (None = the
    dedicated gz kernel) selects one of the other fields."""
    torch = _lib.require_cuda()
    lon, lat, height = (np.ascontiguousarray(a, dtype=np.float64) for a in (lon, lat, height))
    assert lon.shape == lat.shape == height.shape, "Input coordinate arrays must have same shape"
    assert ratio > 0, "Invalid ratio {}. Must be > 0.".format(ratio)
    lo, hi = (0, lon.shape[0]) if rows is None else rows
    table = np.ascontiguousarray(table, dtype=np.float64).reshape(-1, 6)
    M = table.shape[0]
    ncols = M if ncols is None else ncols
    ld = _lib.padded_ld(ncols)
    n = hi - lo
    dev = device or torch.device("cuda", torch.cuda.current_device())
    Gd = torch.empty((n, ld), dtype=torch.float64, device=dev)
    if n == 0 or ld == 0:
        return Gd, ncols
    lonr, sinlat, coslat, radius = _convert_coords(lon[lo:hi], lat[lo:hi], height[lo:hi])
    a_d = [to_device(a, torch, dev) for a in (lonr, sinlat, coslat, radius)]
    tab_d = to_device(table if M else np.zeros((1, 6)), torch, dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    L = _lib.lib()
    if field is None:
        _lib.check(L.gi_tess_gz_assemble(_lib.ptr(a_d[0]), _lib.ptr(a_d[1]), _lib.ptr(a_d[2]),
                                         _lib.ptr(a_d[3]), n, _lib.ptr(tab_d), M, float(ratio),
                                         SI2MGAL, G, _lib.ptr(Gd), ld, _lib.ptr(status),
                                         _lib.stream_ptr()), "gi_tess_gz_assemble")
    else:
        _, s1, s2 = field_scales(field, forward)
        _lib.check(L.gi_tess_field_assemble(FIELD_CODES[field], _lib.ptr(a_d[0]), _lib.ptr(a_d[1]),
                                            _lib.ptr(a_d[2]), _lib.ptr(a_d[3]), n, _lib.ptr(tab_d), M,
                                            float(ratio), s1, s2, _lib.ptr(Gd), ld, _lib.ptr(status),
                                            _lib.stream_ptr()), "gi_tess_field_assemble")
    err, overflow = (int(v) for v in status.cpu())
    if overflow:
        raise OverflowError("tesseroid subdivision stack overflow (STACK_SIZE = 100)")
    if err != 0:  # tesseroid.py:228-229
        warnings.warn("Stopped dividing a tesseroid because it's dimensions would be below the "
                      "minimum numerical threshold (1e-6 degrees or 1e-3 m). Will compute without "
                      "division. Cannot guarantee the accuracy of the solution.", RuntimeWarning)
    return Gd, ncols


def leaf_counts(lon, lat, height, table, ratio=RATIO_G):
    """int32 [n_obs, M] number of leaf cells the adaptive subdivision evaluates per (observation,
    tesseroid) pair (-1 = stack overflow) -- the decision bookkeeping of the engine
    (_tesseroid_numba.py:32-71, 135-157), for diagnostics and parity tests."""
    torch = _lib.require_cuda()
    lon, lat, height = (np.ascontiguousarray(a, dtype=np.float64) for a in (lon, lat, height))
    table = np.ascontiguousarray(table, dtype=np.float64).reshape(-1, 6)
    M, n = table.shape[0], lon.shape[0]
    ld = _lib.padded_ld(M)
    dev = torch.device("cuda", torch.cuda.current_device())
    out = torch.empty((n, ld), dtype=torch.float64, device=dev)
    a_d = [to_device(a, torch, dev) for a in _convert_coords(lon, lat, height)]
    tab_d = to_device(table, torch, dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().gi_tess_gz_leafcount(_lib.ptr(a_d[0]), _lib.ptr(a_d[1]), _lib.ptr(a_d[2]),
                                               _lib.ptr(a_d[3]), n, _lib.ptr(tab_d), M, float(ratio),
                                               _lib.ptr(out), ld, _lib.ptr(status), _lib.stream_ptr()),
               "gi_tess_gz_leafcount")
    return out[:, :M].cpu().numpy().astype(np.int32)


def _field(field, lon, lat, height, model, dens, ratio, njobs, pool, device_out, forward=False):
    """tesseroid.py:156-232 (_dispatcher / _forward_model) for one field"""
    assert njobs > 0, "Invalid number of jobs {}. Must be > 0.".format(njobs)
    if njobs == 1:
        assert pool is None, "njobs should be number of processes in the pool"
    table, rho = model_table(model, dens, "tesseroid")
    ncols = table.shape[0]
    if ncols:
        keep = ~((table[:, 1] - table[:, 0] <= 1e-6) | (table[:, 3] - table[:, 2] <= 1e-6)
                 | (table[:, 4] - table[:, 5] <= 1e-3))
        table, _ = _check_table(table)
        if rho is not None:
            rho = rho[keep]
    torch = _lib.require_cuda()
    M = table.shape[0]
    have_rho = M and rho is not None and np.any(rho != 0)
    if forward:
        # tesseroidforward.py: only the result is wanted -> observation chunks bound the footprint
        n = len(np.atleast_1d(lon))
        res = np.zeros(n)
        chunk = max(1, int((512 << 20) // (8 * max(_lib.padded_ld(ncols), 1))))
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            Gd, _ = assemble(lon, lat, height, table, ratio=ratio, ncols=ncols, rows=(lo, hi),
                             field=field, forward=True)
            if have_rho:
                res[lo:hi] = matvec_padded(Gd, M, to_device(rho, torch, Gd.device), torch).cpu().numpy()
        return res
    Gd, _ = assemble(lon, lat, height, table, ratio=ratio, ncols=ncols, field=None if field == "gz" else field)
    if have_rho:
        res = matvec_padded(Gd, M, to_device(rho, torch, Gd.device), torch)
    else:
        res = torch.zeros(Gd.shape[0], dtype=torch.float64, device=Gd.device)
    if device_out:
        return res, Gd
    return res.cpu().numpy(), Gd[:, :ncols].cpu().numpy()


def gz(lon, lat, height, model, dens=None, ratio=RATIO_G, njobs=1, pool=None, device_out=False):
    """Calculate gz (mGal, density in g/cm^3) of a tesseroid model and the kernel matrix.

    `kernel2d` has one column per non-masked tesseroid; degenerate tesseroids are skipped, which
    (as in the reference, tesseroid.py:98-105 vs :218-231) leaves that many trailing zero columns."""
    return _field("gz", lon, lat, height, model, dens, ratio, njobs, pool, device_out)


def _make(field, ratio0, lines):
    def fn(lon, lat, height, model, dens=None, ratio=ratio0, njobs=1, pool=None, device_out=False):
        return _field(field, lon, lat, height, model, dens, ratio, njobs, pool, device_out)

    fn.__name__ = field
    fn.__doc__ = ("`(result, kernel2d)` of the {} field of a tesseroid model (gravmag/tesseroid.py:{}; "
                  "`njobs`/`pool` accepted and ignored).".format(field, lines))
    return fn


potential = _make("potential", RATIO_V, "324-377")
geoid = _make("geoid", RATIO_V, "380-391")
gx = _make("gx", RATIO_G, "394-404")
gy = _make("gy", RATIO_G, "407-418")
gxx = _make("gxx", RATIO_GG, "433-443")
gxy = _make("gxy", RATIO_GG, "446-456")
gxz = _make("gxz", RATIO_GG, "459-469")
gyy = _make("gyy", RATIO_GG, "472-482")
gyz = _make("gyz", RATIO_GG, "485-495")
gzz = _make("gzz", RATIO_GG, "498-508")
