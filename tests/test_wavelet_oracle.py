"""The oracle's restatement of the wavelet compressors (oracle/oracle_np.py `dwt_per` ... `modelcompressor_3d`,
following gravmag/compressor1D.py:17-60 and compressor3D.py:17-68) against what CAN be pinned without
PyWavelets (absent from this image, un-pinned by the reference -- the path stays "parity unpinned"):

* the db4 filter bank re-derived here from Daubechies' construction (spectral factorisation of the
  half-band polynomial, extremal phase) instead of trusted as typed-in constants;
* the properties the published transform has whatever the implementation: quadrature-mirror pair,
  orthogonality of the periodised transform, four vanishing moments, perfect reconstruction by the
  transpose, ceil(n/2) output lengths with edge repetition for odd n;
* the coefficient-array layouts the reference relies on (`coeffs_to_array`: concatenation in 1-D, Mallat
  packing with zero gaps in 3-D, SURVEY.md 8c: (10, 30, 20) -> (11, 31, 20));
* the compressor identity: with threshold 0 and nesting lengths, Awcp @ DWT(m) == Aw @ m.

What stays unpinned: pywt's phase convention in 'periodization' mode (which sample the first output is
centred on) and its sign convention for dec_hi -- both only shift / flip coefficients consistently on the
kernel and the model side, so the compressed forward `Awcp @ coef` is unchanged by them when thr = 0."""
from math import comb

import numpy as np
import pytest

from oracle import oracle_np as onp


TOL = 5e-12  # accuracy of the published filter constants (see the first test)


def daubechies_scaling_filter(p):
    """extremal-phase Daubechies filter with p vanishing moments: |H(w)|^2 = 2 cos^2p(w/2) P(sin^2(w/2)),
    P(y) = sum_{k<p} C(p-1+k, k) y^k; each root y of P gives z + 1/z = 2 - 4y, keep |z| < 1"""
    roots_y = np.roots([comb(p - 1 + k, k) for k in range(p - 1, -1, -1)])
    zs = []
    for y in roots_y:
        r = np.roots([1.0, -(2.0 - 4.0 * y), 1.0])
        zs.append(r[np.argmin(np.abs(r))])
    h = np.poly(np.concatenate([-np.ones(p), zs])).real
    return h * (np.sqrt(2.0) / h.sum())


def test_db4_filters_follow_from_daubechies_construction():
    h = daubechies_scaling_filter(4)  # reconstruction low-pass; the decomposition filter is its reverse
    assert np.max(np.abs(h[::-1] - onp.DB4_DEC_LO)) < 5e-13
    k = np.arange(8)
    assert np.array_equal(onp.DB4_DEC_HI, -((-1.0) ** k) * onp.DB4_DEC_LO[::-1])  # quadrature mirror
    assert abs(onp.DB4_DEC_LO.sum() - np.sqrt(2)) < 1e-13 and abs(onp.DB4_DEC_HI.sum()) < 1e-13
    # PyWavelets' published 17-digit constants are orthonormal to 1e-12, not to the last bit
    assert abs(np.dot(onp.DB4_DEC_LO, onp.DB4_DEC_LO) - 1.0) < TOL
    for s in range(1, 4):  # double-shift orthogonality
        assert abs(np.dot(onp.DB4_DEC_LO[2 * s:], onp.DB4_DEC_LO[: 8 - 2 * s])) < TOL
    for mom in range(4):  # four vanishing moments of the wavelet
        assert abs(np.dot(onp.DB4_DEC_HI, k.astype(float) ** mom)) < 1e-10


def analysis_matrix(n):
    W = np.empty((2 * ((n + 1) // 2), n))
    for j in range(n):
        e = np.zeros(n)
        e[j] = 1.0
        W[:, j] = np.concatenate(onp.dwt_per(e))
    return W


@pytest.mark.parametrize("n", [8, 10, 16, 30, 64])
def test_periodised_transform_is_orthogonal_for_even_lengths(n):
    W = analysis_matrix(n)
    assert np.max(np.abs(W @ W.T - np.eye(n))) < TOL  # so W.T reconstructs perfectly
    x = np.random.RandomState(n).standard_normal(n)
    cA, cD = onp.dwt_per(x)
    assert np.max(np.abs(W.T @ np.concatenate([cA, cD]) - x)) < TOL * n


@pytest.mark.parametrize("n", [9, 11, 37])
def test_odd_lengths_repeat_the_last_sample(n):
    x = np.random.RandomState(n).standard_normal(n)
    cA, cD = onp.dwt_per(x)
    assert cA.size == cD.size == (n + 1) // 2
    eA, eD = onp.dwt_per(np.append(x, x[-1]))
    assert np.array_equal(cA, eA) and np.array_equal(cD, eD)


def test_details_of_a_cubic_vanish_away_from_the_wrap():
    t = np.arange(64, dtype=float)
    cA, cD = onp.dwt_per(0.3 - 0.2 * t + 0.01 * t ** 2 - 1e-4 * t ** 3)
    assert np.max(np.abs(cD[4:-4])) < 1e-9 and np.max(np.abs(cD)) > 1e-3  # only the periodic seam sees a jump
    cA, cD = onp.dwt_per(np.full(32, 2.5))
    assert np.max(np.abs(cA - 2.5 * np.sqrt(2))) < 1e-13 and np.max(np.abs(cD)) < 1e-13


def test_coefficient_array_layouts():
    x = np.random.RandomState(0).standard_normal(37)
    c = onp.wavedec_1d(x)
    assert [v.size for v in c] == [10, 10, 19]  # [cA2, cD2, cD1]: 37 -> 19 -> 10
    assert np.array_equal(onp.coeffs_to_array_1d(c), np.concatenate(c))
    v = np.random.RandomState(1).standard_normal((10, 30, 20))
    co = onp.wavedecn_3d(v)
    arr = onp.coeffs_to_array_3d(co)
    assert arr.shape == (11, 31, 20)  # 10 -> 5 -> 3: 3 + 3 + 5; 30 -> 15 -> 8: 8 + 8 + 15; 20 -> 10 -> 5
    assert np.array_equal(arr[:3, :8, :5], co[0])
    assert np.array_equal(arr[3:6, 8:16, 5:10], co[1]["ddd"]) and np.array_equal(arr[6:, 16:, 10:], co[2]["ddd"])
    assert np.array_equal(arr[:3, 8:16, :5], co[1]["ada"]) and np.array_equal(arr[6:, :15, :10], co[2]["daa"])
    filled = co[0].size + sum(b.size for d in co[1:] for b in d.values())
    assert np.count_nonzero(arr) == filled < arr.size  # the gaps of non-nesting shapes stay zero
    # separability: the 3-D transform is the 1-D one along each axis in turn
    a0 = onp.dwt_per(onp.dwt_per(onp.dwt_per(v, axis=0)[0], axis=1)[1], axis=2)[0]
    assert np.array_equal(a0, co[2]["ada"])


@pytest.mark.parametrize("kind", ["1D", "3D"])
def test_compressed_forward_equals_dense_at_zero_threshold(kind):
    rng = np.random.RandomState(5)
    shape = (4, 8, 12)  # every axis nests twice (divisible by 4): the transform is orthogonal
    M = int(np.prod(shape))
    Aw, m = rng.standard_normal((7, M)), rng.uniform(0, 0.3, M)
    if kind == "1D":
        got = onp.modelcompressor_1d(m, onp.kernelcompressor_1d(Aw, thr=0.0))
    else:
        got = onp.modelcompressor_3d(m, onp.kernelcompressor_3d(Aw, shape, thr=0.0), shape)
    assert np.max(np.abs(got - Aw @ m)) < 4 * TOL * np.max(np.abs(Aw @ m))
    # the reference's threshold drops only coefficients below 1e-3: bounded forward error
    if kind == "1D":
        cp = onp.kernelcompressor_1d(Aw * 1e-2)
        err = onp.modelcompressor_1d(m, cp) - (Aw * 1e-2) @ m
        assert cp.nnz < Aw.size and np.max(np.abs(err)) <= 1e-3 * np.sum(np.abs(onp.coeffs_to_array_1d(onp.wavedec_1d(m))))
