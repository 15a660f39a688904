"""Host-side bookkeeping of the peer-memory exchange (no GPU): the column slices of a row-sharded
batch must tile [0, ld) in whole 256-column strips of the adjoint kernel, in rank order, for every
world size -- the same table the adjoint epilogue (owner of a strip), the slice update and the
forward kernel's flag waits are driven by (csrc/batched.cu: peer_columns)."""
import ctypes as C

import numpy as np
import pytest

from gravinv3dhmc_b200 import _lib


@pytest.mark.parametrize("ld", [32, 128, 256, 960, 1024, 6016, 72000 + 0, 1 << 20])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8, 16])
def test_column_slices_tile_the_row(ld, world):
    L = _lib.lib()
    col = (C.c_int64 * (world + 1))()
    assert L.gi_peer_columns(ld, world, col) == 0
    col = np.array(col[:])
    assert col[0] == 0 and col[-1] == ld and np.all(np.diff(col) >= 0)
    assert np.all(col[:-1] % 256 == 0)                      # a strip never straddles two slices
    nstrips = -(-ld // 256)
    widths = np.diff(col)
    assert widths.max() - widths[widths > 0].min() <= 256 + (256 - ld % 256) % 256  # balanced to one strip
    # owner lookup as the kernels do it: the last q with col[q] <= v0 (empty slices skipped)
    for strip in {0, nstrips // 2, nstrips - 1}:
        v0 = strip * 256
        o = 0
        while o + 1 < world and v0 >= col[o + 1]:
            o += 1
        assert col[o] <= v0 < col[o + 1]
    if ld == 1 << 20 and world == 8:
        assert list(col) == [k * 131072 for k in range(9)]  # c5 on 8 GPUs: 512 strips each


def test_bad_arguments():
    L = _lib.lib()
    col = (C.c_int64 * 4)()
    assert L.gi_peer_columns(0, 2, col) == _lib.GI_ERR_INVALID
    assert L.gi_peer_columns(1024, 17, col) == _lib.GI_ERR_INVALID
    assert L.gi_peer_columns(1024, 2, None) == _lib.GI_ERR_INVALID
