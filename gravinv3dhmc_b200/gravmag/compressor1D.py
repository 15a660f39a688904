"""1-D wavelet compression of the kernel matrix -- mirror of the reference's
gravmag/compressor1D.py: `kernelcompressor(Kernel_Grv)` (:17-42) and
`modelcompressor(DensityModel, Gkernelsp)` (:45-60).

Every kernel row is transformed with a level-2 db4 DWT in 'periodization' mode (`gi_dwt_db4_l2_1d`),
coefficients with |c| < 0.001 are zeroed and the rows are packed into a CSR matrix on the device;
the forward product is `Awcp @ DWT(model)` (`gi_csr_spmv`).  The arithmetic of the reference lives
in PyWavelets (absent here, no version pinned): conventions restated, parity unpinned.
"""
from __future__ import annotations

import ctypes as C

from .. import _lib
from ._csr import WAVELET_THRESHOLD, DeviceCSR, as_device_vector, rows_to_csr  # noqa: F401


def ncoef(n):
    out = C.c_int64()
    _lib.check(_lib.lib().gi_dwt_db4_l2_1d(None, int(n), None, C.byref(out), None), "gi_dwt_db4_l2_1d")
    return int(out.value)


def kernelcompressor(Kernel_Grv, thr=WAVELET_THRESHOLD):
    """CSR (device) of the thresholded wavelet coefficients of every row of `Kernel_Grv`
    (a CUDA tensor [N, M]; `GravMagModule.Aw`)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    if not hasattr(Kernel_Grv, "data_ptr"):
        import numpy as np

        Kernel_Grv = torch.as_tensor(np.ascontiguousarray(Kernel_Grv, dtype=np.float64)).cuda()
    M = int(Kernel_Grv.shape[1])
    nc = ncoef(M)

    def transform(src, nb, stride, dense):
        _lib.check(L.gi_dwt_db4_l2_1d_batch(C.c_void_p(src), nb, stride, M, _lib.ptr(dense), nc, None,
                                            _lib.stream_ptr()), "gi_dwt_db4_l2_1d_batch")

    return rows_to_csr(Kernel_Grv, transform, nc, thr)


def modelcompressor(DensityModel, Gkernelsp):
    """data = Gkernelsp @ DWT(DensityModel); numpy in -> numpy out, CUDA tensor in -> CUDA tensor."""
    torch = _lib.require_cuda()
    dev = Gkernelsp.data.device
    m, on_dev = as_device_vector(DensityModel, dev)
    nc = Gkernelsp.shape[1]
    coef = torch.empty(nc, dtype=torch.float64, device=dev)
    _lib.check(_lib.lib().gi_dwt_db4_l2_1d(_lib.ptr(m), int(m.numel()), _lib.ptr(coef), None,
                                           _lib.stream_ptr()), "gi_dwt_db4_l2_1d")
    d = Gkernelsp.matvec(coef)
    return d if on_dev else d.cpu().numpy()
