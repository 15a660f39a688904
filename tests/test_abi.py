"""The C-ABI library loads on a CPU-only box and exports every symbol the header declares."""
import ctypes as C
import os
import re

import pytest

from gravinv3dhmc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "gravinv_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(gi_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert L.gi_abi_version() == 1


def test_argument_validation_without_gpu():
    """Bad shapes are rejected with GI_ERR_INVALID before any CUDA call."""
    L = _lib.lib()
    assert L.gi_prism_gz_assemble(None, None, None, 4, None, 10, 1.0, None, 7, None) == _lib.GI_ERR_INVALID
    assert b"bad shape" in L.gi_last_error()
    assert L.gi_tess_gz_assemble(None, None, None, None, 1, None, 4, -1.0, 1.0, 1.0, None, 4, None,
                                 None) == _lib.GI_ERR_INVALID
    plan = C.c_void_p()
    assert L.gi_plan_create(0, 4, 4, 1, C.byref(plan)) == _lib.GI_ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(L.gi_plan_create(4, 5, 4, 1, C.byref(plan)), "plan")
    n = C.c_int64()
    assert L.gi_dwt_db4_l2_1d(None, 6000, None, C.byref(n), None) == 0 and n.value == 2 * 1500 + 3000
    shp = (C.c_int32 * 3)()
    assert L.gi_dwt_db4_l2_3d(None, 10, 30, 20, None, C.byref(shp), None) == 0
    assert list(shp) == [11, 31, 20]  # (10,30,20) packs into an (11,31,20) array


def test_reg_and_constraint_names_raise_like_reference():
    from gravinv3dhmc_b200.inversion._engine import reg_params

    with pytest.raises(ValueError, match="regularization"):
        reg_params("Tikhonov", "mandatory", (1, 1, 1), 1, 1, 1)
    with pytest.raises(ValueError, match="boundary constraint"):
        reg_params("Damping", "reflect", (1, 1, 1), 1, 1, 1)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.GravInvError):
        _lib.require_cuda()


def test_row_split_matches_reference_chunking():
    from gravinv3dhmc_b200.gravmag._common import split_rows

    # gravmag/prism.py:986-996: size//nparts rows per part, remainder to the last
    assert split_rows(600, 1) == [(0, 600)]
    assert split_rows(16384, 8) == [(i * 2048, (i + 1) * 2048) for i in range(8)]
    assert split_rows(10, 4) == [(0, 2), (2, 4), (4, 6), (6, 10)]
    assert split_rows(3, 4) == [(0, 0), (0, 0), (0, 0), (0, 3)]
