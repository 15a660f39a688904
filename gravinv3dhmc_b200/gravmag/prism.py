"""Gravity effect gz of right rectangular prisms and its sensitivity matrix, assembled on the GPU.

Mirror of the reference's gravmag/prism.py `gz` (:911-918) -> `_dispatcher_gravity` (:998-1038)
-> `_gz` (:291-316) -> Cython `_prism.gz` (gravmag/_prism.pyx:263-290): same arguments, same
`(result, kernel2d)` return, same `ValueError` for mismatched coordinate arrays.  `njobs`/`pool`
are accepted and ignored (the reference's worker pool splits rows; here one CUDA kernel covers
every (observation, prism) pair -- `gi_prism_gz_assemble`).

The other prism fields of the reference (`potential, geoid, gx, gy, gxx, gxy, gxz, gyy, gyz, gzz,
tf, bx, by, bz`; prism.py:875-982, 735-870) go through `gi_prism_field_assemble` (csrc/fields.cu),
one templated kernel per field with the same layout; `gz` keeps its dedicated kernels (including the
structured-grid one).
"""
from __future__ import annotations

import numpy as np

from .. import _lib
from ..constants import CM, G, SI2EOTVOS, SI2MGAL, T2NT, g0
from ..utils import dircos
from ._common import matvec_padded, model_table, to_device


def assemble(xp, yp, zp, table, rows=None, device=None):
    """Device sensitivity matrix for an explicit [M,6] bounds table.

    Returns a torch.float64 CUDA tensor of shape [nrows, ld] (ld = M rounded up to 32, padding
    columns zero).  `rows=(lo, hi)` assembles only that observation range (row sharding)."""
    torch = _lib.require_cuda()
    xp, yp, zp = (np.ascontiguousarray(a, dtype=np.float64) for a in (xp, yp, zp))
    if xp.shape != yp.shape or xp.shape != zp.shape:
        raise ValueError("Input arrays xp, yp, and zp must have same length!")  # prism.py:295-296
    lo, hi = (0, xp.shape[0]) if rows is None else rows
    table = np.ascontiguousarray(table, dtype=np.float64).reshape(-1, 6)
    M = table.shape[0]
    ld = _lib.padded_ld(M)
    n = hi - lo
    dev = device or torch.device("cuda", torch.cuda.current_device())
    Gd = torch.empty((n, ld), dtype=torch.float64, device=dev)
    if n == 0 or ld == 0:
        return Gd, M
    x_d, y_d, z_d = (to_device(a[lo:hi], torch, dev) for a in (xp, yp, zp))
    tab_d = to_device(table if M else np.zeros((1, 6)), torch, dev)
    L = _lib.lib()
    _lib.check(L.gi_prism_gz_assemble(_lib.ptr(x_d), _lib.ptr(y_d), _lib.ptr(z_d), n,
                                      _lib.ptr(tab_d), M, G * SI2MGAL, _lib.ptr(Gd), ld,
                                      _lib.stream_ptr()), "gi_prism_gz_assemble")
    _lib.sync()
    return Gd, M


def assemble_grid(xp, yp, zp, mesh, rows=None, device=None):
    """Device sensitivity matrix of a structured prism mesh whose cells share their edges bit for
    bit (`mesh.node_axes()` is not None): every mesh corner is evaluated once per observation and
    shared by the cells around it (`gi_prism_gz_assemble_grid`).  Same bits as `assemble` on
    `mesh.bounds_table()`; returns None when the mesh does not qualify."""
    axes = mesh.node_axes() if hasattr(mesh, "node_axes") else None
    if axes is None or mesh.celltype.__name__ != "Prism":
        return None
    torch = _lib.require_cuda()
    xp, yp, zp = (np.ascontiguousarray(a, dtype=np.float64) for a in (xp, yp, zp))
    if xp.shape != yp.shape or xp.shape != zp.shape:
        raise ValueError("Input arrays xp, yp, and zp must have same length!")  # prism.py:295-296
    lo, hi = (0, xp.shape[0]) if rows is None else rows
    nz, ny, nx = mesh.shape
    cmap = mesh.column_map()
    M = int(mesh.size - len(mesh.mask))
    ld = _lib.padded_ld(M)
    n = hi - lo
    dev = device or torch.device("cuda", torch.cuda.current_device())
    # with a column map the kernel only writes active columns: start from zeros
    Gd = (torch.zeros if cmap is not None else torch.empty)((n, ld), dtype=torch.float64, device=dev)
    if n == 0 or ld == 0:
        return Gd, M
    x_d, y_d, z_d = (to_device(a[lo:hi], torch, dev) for a in (xp, yp, zp))
    xn, yn, zn = (to_device(a, torch, dev) for a in axes)
    cm_d = None if cmap is None else torch.as_tensor(cmap).to(dev)
    _lib.check(_lib.lib().gi_prism_gz_assemble_grid(
        _lib.ptr(x_d), _lib.ptr(y_d), _lib.ptr(z_d), n, _lib.ptr(xn), _lib.ptr(yn), _lib.ptr(zn),
        int(nx), int(ny), int(nz), _lib.ptr(cm_d), M, G * SI2MGAL, _lib.ptr(Gd), ld,
        _lib.stream_ptr()), "gi_prism_gz_assemble_grid")
    _lib.sync()
    return Gd, M


def gz(xp, yp, zp, prisms, dens=None, njobs=1, pool=None, device_out=False):
    """Calculate the g_z gravity component (mGal, density in g/cm^3) and the kernel matrix.

    Returns `(result, kernel2d)`: `result[l] = sum_c density_c * kernel2d[l, c]`,
    `kernel2d` of shape (n_obs, n_active_prisms).  With `device_out=True` both are CUDA tensors
    and `kernel2d` keeps its zero-padded leading dimension (a [n_obs, ld] tensor)."""
    table, rho = model_table(prisms, dens, "prism")
    out = assemble_grid(xp, yp, zp, prisms) if (table.shape[0] and hasattr(prisms, "node_axes")) else None
    Gd, M = out if out is not None else assemble(xp, yp, zp, table)
    torch = _lib.require_cuda()
    if M and rho is not None and np.any(rho != 0):
        res = matvec_padded(Gd, M, to_device(rho, torch, Gd.device), torch)
    else:
        res = torch.zeros(Gd.shape[0], dtype=torch.float64, device=Gd.device)
    if device_out:
        return res, Gd
    return res.cpu().numpy(), Gd[:, :M].cpu().numpy()


# ---------------------------------------------------------------------------------------------
# the other fields
# ---------------------------------------------------------------------------------------------
FIELD_CODES = {"potential": 0, "geoid": 0, "gx": 1, "gy": 2, "gz": 3, "gxx": 4, "gxy": 5, "gxz": 6,
               "gyy": 7, "gyz": 8, "gzz": 9, "tf": 10, "vx": 11, "vy": 12, "vz": 13}


def field_scale(field):
    """factor applied after the 8-corner sum (prism.py:150, 178, 231, 367, 729, 777)"""
    if field == "potential":
        return G
    if field == "geoid":
        return G / g0
    if field in ("gx", "gy", "gz"):
        return G * SI2MGAL
    if field in ("tf", "vx", "vy", "vz"):
        return CM * T2NT
    return G * SI2EOTVOS


def assemble_field(field, xp, yp, zp, table, vec=None, rows=None, device=None):
    """[nrows, ld] device sensitivity matrix of `field` for an explicit [M,6] bounds table."""
    torch = _lib.require_cuda()
    xp, yp, zp = (np.ascontiguousarray(a, dtype=np.float64) for a in (xp, yp, zp))
    if xp.shape != yp.shape or xp.shape != zp.shape:
        raise ValueError("Input arrays xp, yp, and zp must have same length!")  # prism.py:132-133
    lo, hi = (0, xp.shape[0]) if rows is None else rows
    table = np.ascontiguousarray(table, dtype=np.float64).reshape(-1, 6)
    M = table.shape[0]
    ld = _lib.padded_ld(M)
    n = hi - lo
    dev = device or torch.device("cuda", torch.cuda.current_device())
    Gd = torch.empty((n, ld), dtype=torch.float64, device=dev)
    if n == 0 or ld == 0:
        return Gd, M
    x_d, y_d, z_d = (to_device(a[lo:hi], torch, dev) for a in (xp, yp, zp))
    tab_d = to_device(table if M else np.zeros((1, 6)), torch, dev)
    v = None if vec is None else np.ascontiguousarray(vec, dtype=np.float64)
    _lib.check(_lib.lib().gi_prism_field_assemble(FIELD_CODES[field], _lib.ptr(x_d), _lib.ptr(y_d),
                                                  _lib.ptr(z_d), n, _lib.ptr(tab_d), M,
                                                  field_scale(field), _lib.ptr(v), _lib.ptr(Gd), ld,
                                                  _lib.stream_ptr()), "gi_prism_field_assemble")
    _lib.sync()
    return Gd, M


def _gravity_field(field, xp, yp, zp, prisms, dens, device_out):
    table, rho = model_table(prisms, dens, "prism")
    Gd, M = assemble_field(field, xp, yp, zp, table)
    torch = _lib.require_cuda()
    if M and rho is not None and np.any(rho != 0):
        res = matvec_padded(Gd, M, to_device(rho, torch, Gd.device), torch)
    else:
        res = torch.zeros(Gd.shape[0], dtype=torch.float64, device=Gd.device)
    if device_out:
        return res, Gd
    return res.cpu().numpy(), Gd[:, :M].cpu().numpy()


def _make_gravity(field, line):
    def fn(xp, yp, zp, prisms, dens=None, njobs=1, pool=None, device_out=False):
        return _gravity_field(field, xp, yp, zp, prisms, dens, device_out)

    fn.__name__ = field
    fn.__doc__ = ("`(result, kernel2d)` of the {} field of right rectangular prisms (gravmag/prism.py:{}; "
                  "`njobs`/`pool` accepted and ignored).".format(field, line))
    return fn


potential = _make_gravity("potential", "875-882")
geoid = _make_gravity("geoid", "884-891")
gx = _make_gravity("gx", "893-900")
gy = _make_gravity("gy", "902-909")
gxx = _make_gravity("gxx", "920-927")
gxy = _make_gravity("gxy", "929-936")
gxz = _make_gravity("gxz", "938-945")
gyy = _make_gravity("gyy", "947-954")
gyz = _make_gravity("gyz", "956-963")
gzz = _make_gravity("gzz", "965-972")


def _magnetization_table(prisms, pmag, fvec):
    """(table[M,6], m[M,3]) of the prisms the reference would compute (prism.py:711-721, 758-766)"""
    if hasattr(prisms, "bounds_table") and hasattr(prisms, "active_indices"):
        if "magnetization" not in prisms.props and pmag is None:
            return np.zeros((0, 6)), np.zeros((0, 3))
        tab = prisms.bounds_table()
        mags = None if pmag is not None else np.asarray(prisms.props["magnetization"])[prisms.active_indices()]
    else:
        rows, mags = [], []
        for cell in prisms:
            if cell is None or ("magnetization" not in cell.props and pmag is None):
                continue
            rows.append(cell.get_bounds())
            if pmag is None:
                mags.append(cell.props["magnetization"])
        tab = np.asarray(rows, dtype=np.float64).reshape(-1, 6)
        mags = None if pmag is not None else np.asarray(mags)
    M = tab.shape[0]
    if pmag is not None:
        if isinstance(pmag, (float, int)):
            if fvec is None:
                raise TypeError("a scalar magnetisation needs the field direction")
            m = np.tile(pmag * np.asarray(fvec), (M, 1))
        else:
            m = np.tile(np.asarray(pmag, dtype=np.float64), (M, 1))
    elif mags.ndim == 1:  # scalar intensities along the regional field (prism.py:716-717)
        m = mags[:, None] * np.asarray(fvec)[None, :]
    else:
        m = mags.astype(np.float64)
    return tab, m.reshape(M, 3)


def _vector_forward(comps, xp, yp, zp, table, vec, m):
    """sum over prisms of (V vec-rows) . m, assembled in observation chunks that bound the footprint"""
    torch = _lib.require_cuda()
    n, M = len(xp), table.shape[0]
    out = np.zeros(n)
    if M == 0 or n == 0:
        return out
    chunk = max(1, int((256 << 20) // (8 * _lib.padded_ld(M))))
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        for field, col in comps:
            Gd, _ = assemble_field(field, xp, yp, zp, table, vec=vec[col] if isinstance(vec, dict) else vec,
                                   rows=(lo, hi))
            w = to_device(m[:, col], torch, Gd.device)
            out[lo:hi] += matvec_padded(Gd, M, w, torch).cpu().numpy()
    return out


def tf(xp, yp, zp, prisms, inc, dec, pmag=None, njobs=1, pool=None, device_out=False):
    """Total-field magnetic anomaly (nT) and its kernel f.(V f) (gravmag/prism.py:975-982, 665-732).
    The result honours per-prism magnetisation vectors: f.(V m) = (V f).m, three component passes."""
    fvec = np.asarray(dircos(inc, dec), dtype=np.float64)
    table, m = _magnetization_table(prisms, pmag, fvec)
    xp_, yp_, zp_ = (np.ascontiguousarray(a, dtype=np.float64) for a in (xp, yp, zp))
    Gd, M = assemble_field("tf", xp_, yp_, zp_, table, vec=fvec)
    res = _vector_forward((("vx", 0), ("vy", 1), ("vz", 2)), xp_, yp_, zp_, table, fvec, m)
    if device_out:
        torch = _lib.require_cuda()
        return torch.as_tensor(res, device=Gd.device), Gd
    return res, Gd[:, :M].cpu().numpy()


def _b_component(field, xp, yp, zp, prisms, pmag):
    table, m = _magnetization_table(prisms, pmag, None)
    xp_, yp_, zp_ = (np.ascontiguousarray(a, dtype=np.float64) for a in (xp, yp, zp))
    if xp_.shape != yp_.shape or xp_.shape != zp_.shape:
        raise ValueError("Input arrays xp, yp, and zp must have same shape!")  # prism.py:755-756
    # row `field` of V applied to the unit vectors: V is symmetric, so component c of the row is the
    # kernel with vec = e_c
    e = {0: [1.0, 0.0, 0.0], 1: [0.0, 1.0, 0.0], 2: [0.0, 0.0, 1.0]}
    return _vector_forward(((field, 0), (field, 1), (field, 2)), xp_, yp_, zp_, table, e, m)


def bx(xp, yp, zp, prisms, pmag=None):
    """x component of the magnetic induction (nT), gravmag/prism.py:735-778"""
    return _b_component("vx", xp, yp, zp, prisms, pmag)


def by(xp, yp, zp, prisms, pmag=None):
    """y component of the magnetic induction (nT), gravmag/prism.py:781-824"""
    return _b_component("vy", xp, yp, zp, prisms, pmag)


def bz(xp, yp, zp, prisms, pmag=None):
    """z component of the magnetic induction (nT), gravmag/prism.py:827-870"""
    return _b_component("vz", xp, yp, zp, prisms, pmag)
