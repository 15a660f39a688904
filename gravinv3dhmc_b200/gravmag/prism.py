"""Gravity effect gz of right rectangular prisms and its sensitivity matrix, assembled on the GPU.

Mirror of the reference's gravmag/prism.py `gz` (:911-918) -> `_dispatcher_gravity` (:998-1038)
-> `_gz` (:291-316) -> Cython `_prism.gz` (gravmag/_prism.pyx:263-290): same arguments, same
`(result, kernel2d)` return, same `ValueError` for mismatched coordinate arrays.  `njobs`/`pool`
are accepted and ignored (the reference's worker pool splits rows; here one CUDA kernel covers
every (observation, prism) pair -- `gi_prism_gz_assemble`).

Only `gz` is provided: the other 13 prism fields of the reference are outside the inversion path.
"""
from __future__ import annotations

import numpy as np

from .. import _lib
from ..constants import G, SI2MGAL
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
