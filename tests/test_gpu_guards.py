"""Guard-band checks of the kernels' output ranges (SURVEY.md section 5 "sanitizers"): compute-sanitizer
is closed on this GPU pool (profiles/r02_compute_sanitizer_closed.txt), so every output buffer of the
hot kernels is placed inside a larger allocation whose surroundings hold a sentinel bit pattern; after
the launch the sentinels must be untouched (no out-of-bounds store on ragged shapes) and every output
element must have been written (no sentinel left inside).  Run-to-run bitwise determinism -- the
race-detector proxy -- is asserted by the per-kernel tests (test_gpu_leapfrog / batched / fused)."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from gravinv3dhmc_b200 import _lib  # noqa: E402

SENT = -7.25e300  # never produced by the kernels on these inputs
PAD = 4096


class Guarded:
    """a device buffer of `n` doubles with PAD sentinel doubles on either side"""

    def __init__(self, n):
        self.n = int(n)
        self.buf = torch.full((self.n + 2 * PAD,), SENT, dtype=torch.float64, device="cuda")
        self.view = self.buf[PAD: PAD + self.n]

    def ptr(self):
        return C.c_void_p(self.view.data_ptr())

    def check(self, what, written=True):
        torch.cuda.synchronize()
        assert bool((self.buf[:PAD] == SENT).all()) and bool((self.buf[PAD + self.n:] == SENT).all()), \
            what + ": store outside the output range"
        if written:
            assert not bool((self.view == SENT).any()), what + ": output element never written"
        return self.view


@pytest.mark.parametrize("n,m,nch", [(5, 33, 3), (130, 1000, 9), (257, 4100, 17), (600, 6000, 64)])
def test_passes_stay_inside_their_outputs(n, m, nch):
    L = _lib.lib()
    s = _lib.stream_ptr()
    rng = np.random.RandomState(n + m)
    ld = _lib.padded_ld(m)
    A = torch.zeros((n, ld), dtype=torch.float64, device="cuda")
    A[:, :m] = torch.as_tensor(rng.standard_normal((n, m)))
    x = torch.zeros(ld, dtype=torch.float64, device="cuda")
    x[:m] = torch.as_tensor(rng.standard_normal(m))
    r = torch.as_tensor(rng.standard_normal(n)).cuda()
    # single chain: GEMV passes and the single-pass evaluation
    plan = C.c_void_p()
    _lib.check(L.gi_plan_create(n, m, ld, 1, C.byref(plan)))
    d, g = Guarded(n), Guarded(ld)
    _lib.check(L.gi_gemv_fwd(plan, _lib.ptr(A), _lib.ptr(x), d.ptr(), s))
    _lib.check(L.gi_gemv_adj(plan, _lib.ptr(A), _lib.ptr(r), g.ptr(), s))
    d.check("gemv_fwd"), g.check("gemv_adj")
    L.gi_plan_destroy(plan)
    fh = C.c_void_p()
    if L.gi_fused_create(n, m, ld, _lib.ptr(A), s, C.byref(fh)) == 0:
        d, g = Guarded(n), Guarded(ld)
        _lib.check(L.gi_fused_pass(fh, _lib.ptr(x), _lib.ptr(r), None, 1, d.ptr(), g.ptr(), s))
        d.check("fused_pass d"), g.check("fused_pass g")
        L.gi_fused_destroy(fh)
    # batched chains: DMMA contractions
    plan = C.c_void_p()
    _lib.check(L.gi_plan_create(n, m, ld, nch, C.byref(plan)))
    cp, npad = C.c_int32(), C.c_int64()
    _lib.check(L.gi_plan_batch_info(plan, C.byref(cp), C.byref(npad)))
    X = torch.zeros((cp.value, ld), dtype=torch.float64, device="cuda")
    X[:nch, :m] = torch.as_tensor(rng.standard_normal((nch, m)))
    R = torch.zeros((cp.value, npad.value), dtype=torch.float64, device="cuda")
    R[:nch, :n] = torch.as_tensor(rng.standard_normal((nch, n)))
    D, G = Guarded(cp.value * n), Guarded(cp.value * ld)
    _lib.check(L.gi_gemm_fwd(plan, _lib.ptr(A), _lib.ptr(X), D.ptr(), s))
    _lib.check(L.gi_gemm_adj(plan, _lib.ptr(A), _lib.ptr(R), G.ptr(), s))
    Dv = D.check("gemm_fwd").view(cp.value, n)
    Gv = G.check("gemm_adj").view(cp.value, ld)
    ref = X[:nch, :m] @ A[:, :m].T
    assert float((Dv[:nch] - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    ref = R[:nch, :n] @ A[:, :m]
    assert float((Gv[:nch, :m] - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    L.gi_plan_destroy(plan)


@pytest.mark.parametrize("nobs,shape", [(7, (3, 5, 4)), (33, (4, 9, 11))])
def test_assembly_and_weighting_stay_inside(nobs, shape):
    from gravinv3dhmc_b200 import mesher
    from gravinv3dhmc_b200.constants import G, SI2MGAL

    L = _lib.lib()
    s = _lib.stream_ptr()
    nz, ny, nx = shape
    mesh = mesher.PrismMesh((0, 100.0 * nx, 0, 100.0 * ny, 0, 100.0 * nz), (100, 100, 100))
    tab = torch.as_tensor(mesh.bounds_table()).cuda()
    M = tab.shape[0]
    ld = _lib.padded_ld(M)
    rng = np.random.RandomState(nobs)
    xp, yp = (torch.as_tensor(rng.uniform(0, 100.0 * nx, nobs)).cuda() for _ in range(2))
    zp = torch.full((nobs,), -3.0, dtype=torch.float64, device="cuda")
    Gm = Guarded(nobs * ld)
    _lib.check(L.gi_prism_gz_assemble(_lib.ptr(xp), _lib.ptr(yp), _lib.ptr(zp), nobs, _lib.ptr(tab), M, G * SI2MGAL,
                                      Gm.ptr(), ld, s))
    A = Gm.check("prism_gz_assemble").view(nobs, ld)
    assert bool((A[:, M:] == 0).all()) and bool(torch.isfinite(A).all())
    ss = Guarded(ld)
    _lib.check(L.gi_colsumsq(Gm.ptr(), nobs, M, ld, ss.ptr(), 0, s))
    ss.check("colsumsq")
    wm, wi, w2 = Guarded(ld), Guarded(ld), Guarded(ld)
    _lib.check(L.gi_weights_from_sumsq(ss.ptr(), M, 0.5, wm.ptr(), wi.ptr(), w2.ptr(), s))
    wm.check("weights wm", written=False), wi.check("weights wminv", written=False)
    _lib.check(L.gi_scale_columns(Gm.ptr(), nobs, M, ld, wi.ptr(), s))
    A = Gm.check("scale_columns").view(nobs, ld)
    assert float(((A[:, :M] ** 2).sum(0) - 1.0).abs().max()) < 1e-12


def test_wavelet_kernels_stay_inside():
    L = _lib.lib()
    s = _lib.stream_ptr()
    shape = (5, 7, 9)
    M = int(np.prod(shape))
    shp = (C.c_int32 * 3)()
    _lib.check(L.gi_dwt_db4_l2_3d(None, *shape, None, C.byref(shp), None))
    nc = int(shp[0]) * int(shp[1]) * int(shp[2])
    x = torch.as_tensor(np.random.RandomState(2).standard_normal(M)).cuda()
    out = Guarded(nc)
    _lib.check(L.gi_dwt_db4_l2_3d(_lib.ptr(x), *shape, out.ptr(), None, s))
    out.check("dwt_3d")
    n1 = C.c_int64()
    _lib.check(L.gi_dwt_db4_l2_1d(None, M, None, C.byref(n1), None))
    out = Guarded(n1.value)
    _lib.check(L.gi_dwt_db4_l2_1d(_lib.ptr(x), M, out.ptr(), None, s))
    out.check("dwt_1d")
