"""TEST INFRASTRUCTURE ONLY -- the UNMODIFIED reference's tesseroid gz kernel on the deeply subdivided
near-field pairs of config 3 (example/realdata, build container only; needs /root/reference):

    python oracle/make_golden_nearfield.py        # -> tests/golden/nearfield_c3.npz

`gravmag.tesseroid.gz` (numba engine, gravmag/_tesseroid_numba.py:25-72) is run on all 625 x 10 444
pairs; stored are the (observation, cell) indices, leaf counts and kernel values of the pairs the
adaptive subdivision splits into >= 9 leaves (23 587 pairs) -- the entries where
l^2 = r^2 + rc^2 - 2 r rc cos(psi) cancels up to ~1e9-fold.  tests/test_nearfield_oracle.py pins the C
oracle to them bit for bit and tests/test_gpu_nearfield.py compares the CUDA kernel and the reference
with the binary128 evaluation of the same quadrature (oracle/csrc/oracle_tess_quad.c)."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import oracle_np as onp, ref_harness  # noqa: E402
from oracle.make_golden import OUT, in_tmpdir, quiet  # noqa: E402

C3_RANGE = (106.5, 118.5, 16, 28, 2000, -60000)
C3_SPACING = ([-1000, -2000, -5000], 0.5, 0.5)
C3_DIV = [2000, -5000, -15000, -60000]


def main():
    if not ref_harness.available():
        raise SystemExit("reference tree not mounted")
    ns = ref_harness.load()
    e = np.load(os.path.join(OUT, "examples.npz"))
    o, t = e["c3_obs"], e["c3_topo"]
    with quiet(), in_tmpdir():
        m = ns.mesher.TesseroidMeshSegment(C3_RANGE, C3_SPACING, C3_DIV)
        m.carvetopo(t[:, 0], t[:, 1], t[:, 2])
        m.addprop("density", np.zeros(m.size))
        _, K = ns.tesseroid.gz(o[:, 0], o[:, 1], o[:, 2], m, njobs=1)
    # leaf counts: the subdivision bookkeeping (pinned bit-exact elsewhere) picks the stored pairs
    mesh = onp.OracleMesh(C3_RANGE, C3_SPACING, divisionsection=C3_DIV, zdown=False)
    mesh.carvetopo(t[:, 0], t[:, 1], t[:, 2])
    tab, _ = mesh.active_bounds()
    lv = onp.tess_leaves(o[:, 0], o[:, 1], o[:, 2], tab, threads=8)
    oi, ci = np.nonzero(lv >= 9)
    np.savez_compressed(os.path.join(OUT, "nearfield_c3.npz"), obs=oi.astype(np.int32),
                        cell=ci.astype(np.int32), leaves=lv[oi, ci], K=K[oi, ci],
                        shape=np.array(K.shape), ksum=np.array(K.sum()), kmax=np.array(np.abs(K).max()))
    print("nearfield_c3.npz:", oi.size, "pairs with >= 9 leaves of", K.size)


if __name__ == "__main__":
    main()
