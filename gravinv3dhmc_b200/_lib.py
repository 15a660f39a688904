"""ctypes binding of libgravinv_b200.so (the C ABI declared in include/gravinv_b200.h).

There is no CPU fallback: `lib()` raises if the shared library is missing, and every compute call
raises `GravInvError` if CUDA is unavailable.  torch is used only to own device memory and
streams (`Tensor.data_ptr()`, `torch.cuda.current_stream()`).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# GRAVINV_B200_LIB points at an alternative build of the same ABI (kernel tuning experiments)
SO_PATH = os.environ.get("GRAVINV_B200_LIB") or os.path.join(HERE, "_build", "libgravinv_b200.so")

GI_OK, GI_ERR_INVALID, GI_ERR_CUDA, GI_ERR_OVERFLOW, GI_ERR_NOMEM, GI_ERR_BUSY = 0, -1, -2, -3, -4, -5
REG_KINDS = {"Damping": 0, "MS": 1, "Smoothness": 2, "TV": 3}
CONSTRAINTS = {"mandatory": 0, "logarithmic": 1}


class GravInvError(RuntimeError):
    pass


class RegParams(C.Structure):
    _fields_ = [("reg_kind", C.c_int32), ("constraint", C.c_int32), ("nz", C.c_int32),
                ("ny", C.c_int32), ("nx", C.c_int32), ("reserved", C.c_int32),
                ("alpha", C.c_double), ("beta", C.c_double), ("log_factor", C.c_double)]


class HmcConfig(C.Structure):
    _fields_ = [("N", C.c_int64), ("M", C.c_int64), ("ld", C.c_int64), ("fixed", C.c_int32),
                ("nocenter", C.c_int32), ("reg", RegParams)]


class StreamRecord(C.Structure):
    _fields_ = [("chain", C.c_int32), ("accept", C.c_int32), ("L", C.c_int32), ("reserved", C.c_int32),
                ("seq", C.c_int64), ("U", C.c_double), ("U_data", C.c_double), ("U_model", C.c_double),
                ("Hcur", C.c_double), ("Hnew", C.c_double)]


class HmcResult(C.Structure):
    _fields_ = [("accept", C.c_int32), ("L", C.c_int32), ("U", C.c_double), ("U_data", C.c_double),
                ("U_model", C.c_double), ("Hcur", C.c_double), ("Hnew", C.c_double),
                ("Unew", C.c_double), ("Unew_data", C.c_double), ("Unew_model", C.c_double)]


class CgConfig(C.Structure):
    _fields_ = [("N", C.c_int64), ("M", C.c_int64), ("ld", C.c_int64), ("ncols", C.c_int32),
                ("variant", C.c_int32), ("reg", RegParams), ("q", C.c_double), ("stop_tol", C.c_double),
                ("rhomin", C.c_double), ("rhomax", C.c_double)]


CG_REGINV, CG_BOOTSTRAP = 0, 1
STREAM_QUEUE_DEPTH = 4  # GI_STREAM_QUEUE_DEPTH

_P = C.c_void_p
_I64 = C.c_int64
_D = C.c_double

# name -> (restype, argtypes); must list every symbol include/gravinv_b200.h declares
SIGNATURES = {
    "gi_abi_version": (C.c_int, []),
    "gi_last_error": (C.c_char_p, []),
    "gi_device_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(_I64)]),
    "gi_prism_gz_assemble": (C.c_int, [_P, _P, _P, _I64, _P, _I64, _D, _P, _I64, _P]),
    "gi_prism_gz_assemble_grid": (C.c_int, [_P, _P, _P, _I64, _P, _P, _P, C.c_int32, C.c_int32,
                                            C.c_int32, _P, _I64, _D, _P, _I64, _P]),
    "gi_prism_field_assemble": (C.c_int, [C.c_int32, _P, _P, _P, _I64, _P, _I64, _D, _P, _P, _I64, _P]),
    "gi_tess_field_assemble": (C.c_int, [C.c_int32, _P, _P, _P, _P, _I64, _P, _I64, _D, _D, _D, _P, _I64,
                                          _P, _P]),
    "gi_tess_gz_assemble": (C.c_int, [_P, _P, _P, _P, _I64, _P, _I64, _D, _D, _D, _P, _I64, _P, _P]),
    "gi_tess_gz_leafcount": (C.c_int, [_P, _P, _P, _P, _I64, _P, _I64, _D, _P, _I64, _P, _P]),
    "gi_colsumsq": (C.c_int, [_P, _I64, _I64, _I64, _P, C.c_int, _P]),
    "gi_weights_from_sumsq": (C.c_int, [_P, _I64, _D, _P, _P, _P, _P]),
    "gi_scale_columns": (C.c_int, [_P, _I64, _I64, _I64, _P, _P]),
    "gi_plan_create": (C.c_int, [_I64, _I64, _I64, C.c_int32, C.POINTER(_P)]),
    "gi_plan_destroy": (C.c_int, [_P]),
    "gi_plan_info": (C.c_int, [_P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "gi_gemv_fwd": (C.c_int, [_P, _P, _P, _P, _P]),
    "gi_data_sum": (C.c_int, [_P, _P, _P, _P, _P]),
    "gi_residual": (C.c_int, [_P, _P, _P, _P, _I64, _P, _P, _P]),
    "gi_gemv_adj": (C.c_int, [_P, _P, _P, _P, _P]),
    "gi_update": (C.c_int, [_P, C.POINTER(RegParams), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                            _D, _D, C.c_int, _P, _P]),
    "gi_fused_create": (C.c_int, [_I64, _I64, _I64, _P, _P, C.POINTER(_P)]),
    "gi_fused_destroy": (C.c_int, [_P]),
    "gi_fused_pass": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P, _P, _P]),
    "gi_hmc_create": (C.c_int, [C.POINTER(HmcConfig), _P, _P, _P, _P, _P, _P, _P, _P,
                                C.POINTER(_P)]),
    "gi_hmc_destroy": (C.c_int, [_P]),
    "gi_hmc_set_reg": (C.c_int, [_P, C.POINTER(RegParams)]),
    "gi_hmc_set_state": (C.c_int, [_P, _P]),
    "gi_hmc_get_state": (C.c_int, [_P, _P, _P, _P]),
    "gi_hmc_get_misfit": (C.c_int, [_P, C.POINTER(_D), C.POINTER(_D), C.POINTER(_D), _P]),
    "gi_hmc_propose": (C.c_int, [_P, _P, C.c_int32, _D, _D, C.POINTER(HmcResult), _P, _P]),
    "gi_hmc_propose_philox": (C.c_int, [_P, C.c_uint64, C.c_uint64, _D, C.c_int32, _D,
                                        C.POINTER(HmcResult)]),
    "gi_hmc_leapfrog_steps": (C.c_int, [_P, _P, C.c_int32, _D]),
    "gi_hmc_launch_count": (_I64, [_P]),
    "gi_hmc_eval_path": (C.c_int32, [_P]),
    "gi_hmc_stream": (_P, [_P]),
    "gi_plan_batch_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(_I64)]),
    "gi_gemm_fwd": (C.c_int, [_P, _P, _P, _P, _P]),
    "gi_data_sum_batched": (C.c_int, [_P, _P, _P, _P, _P]),
    "gi_residual_batched": (C.c_int, [_P, _P, _P, _P, _I64, _P, _P, _P]),
    "gi_gemm_adj": (C.c_int, [_P, _P, _P, _P, _P]),
    "gi_update_batched": (C.c_int, [_P, C.POINTER(RegParams), _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                    _P, _P, _P, _D, _P, C.c_int32, C.c_int32, _P, _P]),
    "gi_hmcb_create": (C.c_int, [C.POINTER(HmcConfig), C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P,
                                 C.POINTER(_P)]),
    "gi_hmcb_destroy": (C.c_int, [_P]),
    "gi_hmcb_set_reg": (C.c_int, [_P, C.POINTER(RegParams)]),
    "gi_hmcb_set_state": (C.c_int, [_P, _P]),
    "gi_hmcb_get_state": (C.c_int, [_P, _P, _P, _P]),
    "gi_hmcb_get_misfit": (C.c_int, [_P, _P, _P, _P, _P]),
    "gi_hmcb_propose": (C.c_int, [_P, _P, _P, _D, _P, _P, _P, _P]),
    "gi_hmcb_propose_philox": (C.c_int, [_P, C.c_uint64, C.c_uint64, _D, _P, _D, _P]),
    "gi_hmcb_set_shard": (C.c_int, [_P, _I64, _P, _P, C.c_int32, _P, _P, _P]),
    "gi_peer_create": (C.c_int, [C.c_int32, C.c_int32, _I64, C.POINTER(_P)]),
    "gi_peer_export": (C.c_int, [_P, _P]),
    "gi_peer_connect": (C.c_int, [_P, _P]),
    "gi_peer_destroy": (C.c_int, [_P]),
    "gi_peer_bytes_sent": (_I64, [_P]),
    "gi_peer_allreduce_small": (C.c_int, [_P, _P, C.c_int32, _P]),
    "gi_peer_columns": (C.c_int, [_I64, C.c_int32, _P]),
    "gi_hmcb_peer_bytes": (_I64, [_P, C.c_int32]),
    "gi_hmcb_set_peer": (C.c_int, [_P, _P, _I64, _P]),
    "gi_hmcb_owned_columns": (C.c_int, [_P, C.POINTER(_I64), C.POINTER(_I64)]),
    "gi_hmcb_stream_begin": (C.c_int, [_P, _D]),
    "gi_hmcb_stream_feed": (C.c_int, [_P, C.c_int32, C.c_int32, _D, _P]),
    "gi_hmcb_stream_feed_dev": (C.c_int, [_P, C.c_int32, C.c_int32, _D, _P]),
    "gi_hmcb_stream_runway": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "gi_hmcb_stream_close_chain": (C.c_int, [_P, C.c_int32]),
    "gi_hmcb_stream_advance": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int32), _P]),
    "gi_hmcb_stream_advance_begin": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "gi_hmcb_stream_advance_end": (C.c_int, [_P, _P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "gi_hmcb_stream_queue_space": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32)]),
    "gi_hmcb_stream_cancel": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32)]),
    "gi_hmcb_leapfrog_steps": (C.c_int, [_P, _P, C.c_int32, _D]),
    "gi_hmcb_launch_count": (_I64, [_P]),
    "gi_hmcb_padded_chains": (C.c_int32, [_P]),
    "gi_legacy_randn_scaled": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(_D), _I64,
                                         _D, _P]),
    "gi_ring_store_release": (None, [_P, _I64]),
    "gi_ring_load_acquire": (_I64, [_P]),
    "gi_dwt_db4_l2_1d": (C.c_int, [_P, _I64, _P, C.POINTER(_I64), _P]),
    "gi_dwt_db4_l2_3d": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P,
                                   C.POINTER(C.c_int32 * 3), _P]),
    "gi_dwt_db4_l2_1d_batch": (C.c_int, [_P, _I64, _I64, _I64, _P, _I64, C.POINTER(_I64), _P]),
    "gi_dwt_db4_l2_3d_batch": (C.c_int, [_P, _I64, _I64, C.c_int32, C.c_int32, C.c_int32, _P, _I64,
                                         C.POINTER(C.c_int32 * 3), _P]),
    "gi_csr_count": (C.c_int, [_P, _I64, _I64, _I64, _D, _P, _P]),
    "gi_csr_fill": (C.c_int, [_P, _I64, _I64, _I64, _D, _P, _P, _P, _P]),
    "gi_hmc_set_wavelet": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _I64]),
    "gi_csr_spmv": (C.c_int, [_P, _P, _P, _I64, _P, _P, _P]),
    "gi_stats_create": (C.c_int, [_I64, _I64, C.c_int32, C.POINTER(_P)]),
    "gi_stats_destroy": (C.c_int, [_P]),
    "gi_stats_reset": (C.c_int, [_P]),
    "gi_stats_window": (C.c_int, [_P, _I64, _I64]),
    "gi_stats_add": (C.c_int, [_P, C.c_int32, _P, _P, _P]),
    "gi_stats_result": (C.c_int, [_P, C.c_int32, _P, _P, C.POINTER(_I64), C.POINTER(_I64), _P]),
    "gi_stats_result_dev": (C.c_int, [_P, C.c_int32, _P, _P, _P, C.POINTER(_I64), C.POINTER(_I64), _P]),
    "gi_stats_launch_count": (_I64, [_P]),
    "gi_hmc_attach_stats": (C.c_int, [_P, _P, C.c_int32, _P]),
    "gi_hmcb_attach_stats": (C.c_int, [_P, _P, _P]),
    "gi_cg_create": (C.c_int, [C.POINTER(CgConfig), _P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(_P)]),
    "gi_cg_destroy": (C.c_int, [_P]),
    "gi_cg_set_wavelet": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _I64]),
    "gi_cg_set_shard": (C.c_int, [_P, _I64, _P, _P, _P, _P]),
    "gi_cg_run": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, _P]),
    "gi_cg_get_result": (C.c_int, [_P, _P, _P, _P]),
    "gi_cg_launch_count": (_I64, [_P]),
}

SHARD_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.c_int32, C.c_int32)
CG_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32)

_LIB = None


def lib():
    """The loaded shared library; raises (no fallback) if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise GravInvError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). gravinv3dhmc_b200 has no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.gi_abi_version() != 1:
            raise GravInvError("libgravinv_b200.so ABI version mismatch; rebuild")
        _LIB = L
    return _LIB


def check(rc: int, what: str = ""):
    """Map a GI_ERR_* return code to the exception type the reference raises."""
    if rc == GI_OK:
        return
    msg = lib().gi_last_error().decode("utf-8", "replace")
    if rc == GI_ERR_INVALID:
        raise ValueError(msg or what)
    if rc == GI_ERR_OVERFLOW:
        raise OverflowError(msg or what)
    if rc == GI_ERR_NOMEM:
        raise MemoryError(msg or what)
    raise GravInvError(f"{what}: {msg}")


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise GravInvError("gravinv3dhmc_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    lib()
    return torch


def ptr(t):
    """device (or host) pointer of a torch tensor / numpy array / None as c_void_p"""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def stream_ptr(torch=None):
    import torch as _t

    return C.c_void_p(_t.cuda.current_stream().cuda_stream)


def sync():
    """wait for the current CUDA stream"""
    import torch as _t

    _t.cuda.current_stream().synchronize()


def legacy_randn_scaled(rs, n: int, scale: float, out):
    """out[:n] = rs.randn(n) * scale, bit for bit, for a numpy legacy RandomState `rs` (its state is
    advanced exactly as numpy would); `out` is a C-contiguous float64 numpy array (e.g. a pinned or
    shared-memory slot).  ~2x numpy's pace, GIL released."""
    import numpy as np

    name, key, pos, has_gauss, cached = rs.get_state()
    key = np.ascontiguousarray(key, dtype=np.uint32)
    c_pos, c_has, c_cached = C.c_int32(int(pos)), C.c_int32(int(has_gauss)), C.c_double(float(cached))
    check(lib().gi_legacy_randn_scaled(C.c_void_p(key.ctypes.data), C.byref(c_pos), C.byref(c_has),
                                       C.byref(c_cached), int(n), float(scale), C.c_void_p(out.ctypes.data)),
          "gi_legacy_randn_scaled")
    rs.set_state((name, key, c_pos.value, c_has.value, c_cached.value))
    return out


def padded_ld(M: int) -> int:
    """leading dimension: M rounded up to 32 doubles (256 B) so every row starts sector-aligned"""
    return ((int(M) + 31) // 32) * 32
