"""3-D wavelet compression of the kernel matrix -- mirror of the reference's
gravmag/compressor3D.py: `kernelcompressor(Kernel_Grv, mshape)` (:17-44) and
`modelcompressor(DensityModel, Gkernelsp, mshape)` (:47-68).

Every kernel row is reshaped to (CZ, CY, CX), transformed with a level-2 db4 `wavedecn` in
'periodization' mode and packed like `pywt.coeffs_to_array` (`gi_dwt_db4_l2_3d`); coefficients with
|c| < 0.001 are zeroed and the rows become a device CSR matrix; forward = `Awcp @ DWT3(model)`.
PyWavelets is absent here and unpinned by the reference: conventions restated, parity unpinned.
"""
from __future__ import annotations

import ctypes as C

from .. import _lib
from ._csr import WAVELET_THRESHOLD, DeviceCSR, as_device_vector, rows_to_csr  # noqa: F401


def coeff_shape(mshape):
    shp = (C.c_int32 * 3)()
    cz, cy, cx = (int(v) for v in mshape)
    _lib.check(_lib.lib().gi_dwt_db4_l2_3d(None, cz, cy, cx, None, C.byref(shp), None),
               "gi_dwt_db4_l2_3d")
    return tuple(int(v) for v in shp)


def kernelcompressor(Kernel_Grv, mshape, thr=WAVELET_THRESHOLD):
    torch = _lib.require_cuda()
    L = _lib.lib()
    if not hasattr(Kernel_Grv, "data_ptr"):
        import numpy as np

        Kernel_Grv = torch.as_tensor(np.ascontiguousarray(Kernel_Grv, dtype=np.float64)).cuda()
    cz, cy, cx = (int(v) for v in mshape)
    if cz * cy * cx != int(Kernel_Grv.shape[1]):
        raise ValueError("cannot reshape array of size {} into shape {}".format(
            int(Kernel_Grv.shape[1]), (cz, cy, cx)))  # numpy's reshape error in compressor3D.py:33
    fs = coeff_shape(mshape)
    nc = fs[0] * fs[1] * fs[2]

    def transform(src, nb, stride, dense):
        _lib.check(L.gi_dwt_db4_l2_3d_batch(C.c_void_p(src), nb, stride, cz, cy, cx, _lib.ptr(dense),
                                            nc, None, _lib.stream_ptr()), "gi_dwt_db4_l2_3d_batch")

    return rows_to_csr(Kernel_Grv, transform, nc, thr)


def modelcompressor(DensityModel, Gkernelsp, mshape):
    torch = _lib.require_cuda()
    dev = Gkernelsp.data.device
    m, on_dev = as_device_vector(DensityModel, dev)
    cz, cy, cx = (int(v) for v in mshape)
    coef = torch.empty(Gkernelsp.shape[1], dtype=torch.float64, device=dev)
    _lib.check(_lib.lib().gi_dwt_db4_l2_3d(_lib.ptr(m), cz, cy, cx, _lib.ptr(coef), None,
                                           _lib.stream_ptr()), "gi_dwt_db4_l2_3d")
    d = Gkernelsp.matvec(coef)
    return d if on_dev else d.cpu().numpy()
