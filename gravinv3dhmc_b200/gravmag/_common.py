"""Shared host-side plumbing of the kernel builders: model -> bounds table, row chunking,
device buffers.  torch is used for device memory only."""
from __future__ import annotations

import numpy as np

from .. import _lib


def split_rows(size: int, nparts: int):
    """Contiguous row chunks exactly like the reference's `_split_arrays`
    (gravmag/prism.py:986-996, gravmag/tesseroid.py:311-321): `size // nparts` rows per part,
    the remainder goes to the last part."""
    if nparts < 1:
        raise AssertionError("Invalid number of jobs {}. Must be > 0.".format(nparts))
    n = size // nparts
    strides = [(i * n, (i + 1) * n) for i in range(nparts - 1)]
    strides.append((strides[-1][-1] if strides else 0, size))
    return strides


def model_table(model, dens, kind):
    """(table[M,6], density[M] or None) of the cells the reference would compute, in order.

    A cell is skipped when it is None (masked) or when it has no 'density' property and `dens` is
    None (prism.py:299-300, tesseroid.py:131-134).  `model` may be one of our meshes (fast path)
    or any iterable of Prism / Tesseroid / None objects."""
    if hasattr(model, "bounds_table") and hasattr(model, "active_indices"):
        has_density = "density" in model.props
        if not has_density and dens is None:
            return np.zeros((0, 6)), None
        tab = model.bounds_table()
        if dens is not None:
            rho = np.full(tab.shape[0], float(dens))
        else:
            rho = np.asarray(model.props["density"], dtype=np.float64)[model.active_indices()]
        return tab, rho
    rows, rho = [], []
    for cell in model:
        if cell is None or ("density" not in cell.props and dens is None):
            continue
        rows.append(cell.get_bounds())
        rho.append(float(dens) if dens is not None else float(cell.props["density"]))
    return np.asarray(rows, dtype=np.float64).reshape(-1, 6), np.asarray(rho, dtype=np.float64)


def to_device(a, torch, device=None):
    t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))
    return t.to(device or "cuda", non_blocking=False)


def matvec_padded(Gpad, M, vec, torch):
    """result = G[:, :M] @ vec through the library's deterministic forward kernel."""
    L = _lib.lib()
    N, ld = Gpad.shape
    x = torch.zeros(ld, dtype=torch.float64, device=Gpad.device)
    x[:M] = vec
    d = torch.zeros(N, dtype=torch.float64, device=Gpad.device)
    if N == 0 or M == 0:
        return d
    import ctypes as C

    plan = C.c_void_p()
    _lib.check(L.gi_plan_create(N, M, ld, 1, C.byref(plan)), "gi_plan_create")
    try:
        _lib.check(L.gi_gemv_fwd(plan, _lib.ptr(Gpad), _lib.ptr(x), _lib.ptr(d), _lib.stream_ptr()),
                   "gi_gemv_fwd")
        _lib.sync()
    finally:
        L.gi_plan_destroy(plan)
    return d


def rmatvec_padded(Gpad, M, vec, torch):
    """result = G[:, :M].T @ vec through the library's deterministic adjoint kernel."""
    L = _lib.lib()
    N, ld = Gpad.shape
    r = torch.as_tensor(vec, dtype=torch.float64, device=Gpad.device).contiguous()
    g = torch.zeros(ld, dtype=torch.float64, device=Gpad.device)
    if N == 0 or M == 0:
        return g[:M]
    import ctypes as C

    plan = C.c_void_p()
    _lib.check(L.gi_plan_create(N, M, ld, 1, C.byref(plan)), "gi_plan_create")
    try:
        _lib.check(L.gi_gemv_adj(plan, _lib.ptr(Gpad), _lib.ptr(r), _lib.ptr(g), _lib.stream_ptr()),
                   "gi_gemv_adj")
        _lib.sync()
    finally:
        L.gi_plan_destroy(plan)
    return g[:M]
