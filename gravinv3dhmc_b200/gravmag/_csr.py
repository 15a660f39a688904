"""Device CSR matrix produced by the wavelet kernel compressors (the reference returns a
scipy.sparse.csr_matrix, gravmag/compressor1D.py:40, compressor3D.py:42)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib

WAVELET_THRESHOLD = 0.001  # compressor1D.py:25, compressor3D.py:25
_ROWS_PER_BATCH_BYTES = 1 << 30


class DeviceCSR:
    """CSR on the GPU: `indptr` int64 [nrows+1], `indices` int32, `data` float64.  `A @ v`
    runs `gi_csr_spmv`; `toscipy()` gives the scipy matrix the reference would hold."""

    def __init__(self, shape, indptr, indices, data):
        self.shape = tuple(int(v) for v in shape)
        self.indptr, self.indices, self.data = indptr, indices, data

    @property
    def nnz(self):
        return int(self.data.numel())

    def toscipy(self):
        import scipy.sparse as sp

        return sp.csr_matrix((self.data.cpu().numpy(), self.indices.cpu().numpy(),
                              self.indptr.cpu().numpy()), shape=self.shape)

    def matvec(self, v_dev, out=None):
        torch = _lib.require_cuda()
        if out is None:
            out = torch.empty(self.shape[0], dtype=torch.float64, device=self.data.device)
        _lib.check(_lib.lib().gi_csr_spmv(_lib.ptr(self.indptr), _lib.ptr(self.indices),
                                          _lib.ptr(self.data), self.shape[0], _lib.ptr(v_dev),
                                          _lib.ptr(out), _lib.stream_ptr()), "gi_csr_spmv")
        return out

    def __matmul__(self, v):
        torch = _lib.require_cuda()
        if hasattr(v, "data_ptr"):
            return self.matvec(v.reshape(-1).contiguous())
        vd = torch.as_tensor(np.ascontiguousarray(v, dtype=np.float64).reshape(-1),
                             device=self.data.device)
        return self.matvec(vd).cpu().numpy()


def rows_to_csr(Aw, transform, ncoef, thr):
    """Transform every row of the device matrix `Aw` ([N, M] view, row stride Aw.stride(0)) with
    `transform(src_ptr, batch, row_stride, out_tensor)` in row batches, zero |c| < thr and pack the
    survivors into a DeviceCSR (column order ascending, like scipy's csr_matrix of a dense array)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    N = int(Aw.shape[0])
    stride = int(Aw.stride(0))
    dev = Aw.device
    s = _lib.stream_ptr()
    batch = int(max(1, min(65535, _ROWS_PER_BATCH_BYTES // (8 * max(ncoef, 1)), max(N, 1))))
    dense = torch.empty((batch, ncoef), dtype=torch.float64, device=dev)
    counts = torch.zeros(N, dtype=torch.int64, device=dev)
    # pass 1: counts per row
    for r0 in range(0, N, batch):
        nb = min(batch, N - r0)
        transform(Aw.data_ptr() + 8 * r0 * stride, nb, stride, dense)
        _lib.check(L.gi_csr_count(_lib.ptr(dense), nb, ncoef, ncoef, float(thr),
                                  C.c_void_p(counts.data_ptr() + 8 * r0), s), "gi_csr_count")
    indptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(counts, 0)
    nnz = int(indptr[-1])
    indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
    data = torch.empty(max(nnz, 1), dtype=torch.float64, device=dev)[:nnz]
    # pass 2: fill (indptr is absolute, so every batch writes straight into the final arrays)
    for r0 in range(0, N, batch):
        nb = min(batch, N - r0)
        transform(Aw.data_ptr() + 8 * r0 * stride, nb, stride, dense)
        _lib.check(L.gi_csr_fill(_lib.ptr(dense), nb, ncoef, ncoef, float(thr),
                                 C.c_void_p(indptr.data_ptr() + 8 * r0), _lib.ptr(indices),
                                 _lib.ptr(data), s), "gi_csr_fill")
    _lib.sync()
    return DeviceCSR((N, ncoef), indptr, indices, data)


def as_device_vector(m, device):
    torch = _lib.require_cuda()
    if hasattr(m, "data_ptr"):
        return m.reshape(-1).contiguous(), True
    return torch.as_tensor(np.ascontiguousarray(m, dtype=np.float64).reshape(-1), device=device), False
