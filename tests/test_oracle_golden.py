"""Pins the CPU oracle (oracle/fir_oracle.c) against every known-answer vector the reference
holds for the batched-FIR path and against SciPy for the API SciPy specifies.

CPU-only; no CUDA involved.  If this file is red, no GPU parity claim means anything.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O

scipy_signal = pytest.importorskip("scipy.signal")


# --- (1) reference golden vector: crates/scir-gpu/src/lib.rs:1251-1260 -------------------------
X24 = np.array([[1.0, 2.0, 3.0, 4.0], [0.5, 0.0, -0.5, -1.0]], dtype=np.float32)
TAPS3 = np.array([0.25, 0.5, 0.25], dtype=np.float32)
Y24 = np.array([[0.25, 1.0, 2.0, 3.0], [0.125, 0.25, 0.0, -0.5]], dtype=np.float64)


def test_reference_golden_2x4_f32():
    y = O.fir1d_batched_f32(X24, TAPS3)
    assert y.shape == (2, 4)
    np.testing.assert_allclose(y.astype(np.float64), Y24, atol=1e-7, rtol=1e-7)


def test_reference_golden_2x4_f64_and_acc64():
    np.testing.assert_allclose(O.fir1d_batched_f64(X24, TAPS3), Y24, atol=1e-12, rtol=1e-12)
    np.testing.assert_allclose(O.fir1d_batched_f32_acc64(X24, TAPS3), Y24, atol=1e-12, rtol=1e-12)


# --- (3)/(4) the naive restatement the reference's tests inline (gpu/lib.rs:1279-1290,
#             sig/lib.rs:679-691), written independently in Python -------------------------------
def naive_fir(x, taps, dtype):
    b, n = x.shape
    k = taps.size
    y = np.zeros((b, n), dtype=dtype)
    for bi in range(b):
        for i in range(n):
            acc = dtype(0)
            start = max(i + 1 - k, 0)
            for t_idx, xi in enumerate(range(i, start - 1, -1)):
                acc = dtype(acc + dtype(taps[k - 1 - t_idx] * x[bi, xi]))
            y[bi, i] = acc
    return y


def test_matches_naive_bit_exact_f32():
    # sig/lib.rs:670-696 asserts tol = 0.0 between dispatch(Device::Cpu) and the naive loop
    y = O.fir1d_batched_f32(X24, TAPS3)
    assert np.array_equal(y, naive_fir(X24, TAPS3, np.float32))
    rng = np.random.RandomState(0)
    x = (rng.rand(3, 40).astype(np.float32) * 2 - 1)
    taps = (1.0 / (np.arange(7, dtype=np.float32) + 1.0)).astype(np.float32)
    assert np.array_equal(O.fir1d_batched_f32(x, taps), naive_fir(x, taps, np.float32))


def test_random_f64_matches_naive():
    # gpu/lib.rs:1263-1298: 3x32, k=5, taps 1/(i+1), 1e-12
    rng = np.random.RandomState(1)
    x = rng.rand(3, 32) * 2 - 1
    taps = 1.0 / (np.arange(5, dtype=np.float64) + 1.0)
    np.testing.assert_allclose(O.fir1d_batched_f64(x, taps), naive_fir(x, taps, np.float64),
                               atol=1e-12, rtol=1e-12)


def test_mt_equals_single_thread():
    rng = np.random.RandomState(2)
    x = (rng.rand(13, 300).astype(np.float32) * 2 - 1)
    taps = rng.randn(31).astype(np.float32)
    y1 = O.fir1d_batched_f32(x, taps)
    ymt, used = O.fir1d_batched_f32_mt(x, taps, 4)
    assert used == 4 and np.array_equal(y1, ymt)


# --- tap order: fir1d_batched(x, taps) == lfilter(taps[::-1], [1], x) --------------
def test_tap_order_is_reversed_lfilter():
    rng = np.random.RandomState(3)
    x = (rng.rand(4, 200).astype(np.float32) * 2 - 1)
    taps = (1.0 / (np.arange(31) + 1.0)).astype(np.float32)          # fir_bench.rs:14-18
    want = scipy_signal.lfilter(taps[::-1].astype(np.float64), [1.0], x.astype(np.float64), axis=-1)
    got = O.fir1d_batched_f32_acc64(x, taps)
    np.testing.assert_allclose(got, want, atol=1e-12, rtol=1e-12)
    wrong = scipy_signal.lfilter(taps.astype(np.float64), [1.0], x.astype(np.float64), axis=-1)
    assert np.abs(got - wrong).max() > 0.1
    # and the reference-order f32 sum is within the north-star tolerance of the f64 judge
    tol = 1e-5 * np.abs(taps).sum() * np.abs(x).max()
    assert np.abs(O.fir1d_batched_f32(x, taps) - got).max() <= tol


def test_edge_shapes():
    taps = np.array([1.0, -2.0, 3.0], dtype=np.float32)
    assert O.fir1d_batched_f32(np.zeros((0, 5), np.float32), taps).shape == (0, 5)
    assert O.fir1d_batched_f32(np.zeros((3, 0), np.float32), taps).shape == (3, 0)
    x = np.array([[2.0]], dtype=np.float32)            # n < k
    assert O.fir1d_batched_f32(x, taps)[0, 0] == 6.0   # last tap times newest sample
    y = O.fir1d_batched_f32(np.array([[1, 1, 1, 1]], np.float32), np.array([5.0], np.float32))
    assert np.array_equal(y, np.full((1, 4), 5.0, np.float32))


# --- (5) legacy 2/3 resampler fixture: sig/lib.rs:655-668 ---------------------------------------
def test_legacy_taps_match_reference_literals(golden_dir):
    ref = np.load(os.path.join(golden_dir, "legacy_resample_taps.npy"))
    np.testing.assert_allclose(O.legacy_resample_taps(), ref, atol=1e-15, rtol=0)
    np.testing.assert_allclose(ref, 2 * scipy_signal.firwin(31, 1.0 / 3.0, window="hamming"),
                               atol=1e-15, rtol=0)


def test_legacy_resample_poly_matches_fixture(golden_dir):
    fx = os.path.join(golden_dir, "reference_fixtures")
    x = np.load(os.path.join(fx, "sosfilt_input.npy"))
    want = np.load(os.path.join(fx, "resample_poly_output.npy"))
    np.testing.assert_allclose(x, np.linspace(0, 1, 32, endpoint=False))
    ref_taps = np.load(os.path.join(golden_dir, "legacy_resample_taps.npy"))
    for h in (None, ref_taps):
        got = O.legacy_resample_poly_2_3(x, h)
        assert got.shape == want.shape == (22,)
        # the reference's own tolerance (atol = 2e-2, rtol = 1e-6): its filter is not SciPy's
        np.testing.assert_allclose(got, want, atol=2e-2, rtol=1e-6)
    # and the SciPy-exact path reproduces the fixture to rounding
    h = scipy_signal.firwin(2 * 10 * 3 + 1, 1.0 / 3.0, window=("kaiser", 5.0)).astype(np.float32)
    got = O.resample_poly(x[None, :].astype(np.float32), 2, 3, h)[0]
    np.testing.assert_allclose(got, want, atol=2e-6, rtol=0)


# --- (6) filtfilt structure: fixture is sosfilt(sos, sosfilt(sos,x)[::-1])[::-1]
#         (scripts/gen_signal_fixtures.py:28); same structure with an FIR numerator ---------------
def test_filtfilt_refstyle_structure(scipy_vectors, golden_dir):
    v = scipy_vectors
    got = O.filtfilt_fir_nopad(v["filtfilt_b"], v["filtfilt_x"])
    np.testing.assert_allclose(got, v["filtfilt_refstyle"], atol=1e-12, rtol=1e-12)
    # the IIR fixture itself pins only the structure: an FIR cannot reproduce an IIR output,
    # but sosfilt-as-lfilter forward/backward/zero-state is what filtfilt_output.npy holds.
    fx = os.path.join(golden_dir, "reference_fixtures")
    sos = np.load(os.path.join(fx, "butter_sos.npy"))
    x = np.load(os.path.join(fx, "sosfilt_input.npy"))
    want = np.load(os.path.join(fx, "filtfilt_output.npy"))
    got = scipy_signal.sosfilt(sos, scipy_signal.sosfilt(sos, x)[::-1])[::-1]
    np.testing.assert_allclose(got, want, atol=1e-12)


# --- SciPy-specified API ------------------------------------------------------------------------
def test_lfilter_fir_vs_scipy(scipy_vectors):
    v = scipy_vectors
    y = O.lfilter_fir(v["lfilter_b"], v["lfilter_x"])
    np.testing.assert_allclose(y, v["lfilter_y64"], atol=1e-12, rtol=1e-12)
    tol = 1e-5 * np.abs(v["lfilter_b"]).sum() * np.abs(v["lfilter_x"]).max()
    assert np.abs(y - v["lfilter_y"]).max() <= tol            # SciPy's own f32 result
    y, zf = O.lfilter_fir(v["lfilter_b"], v["lfilter_x"], zi=v["lfilter_zi"])
    np.testing.assert_allclose(y, v["lfilter_y_zi"], atol=1e-12, rtol=1e-12)
    np.testing.assert_allclose(zf, v["lfilter_zf"], atol=1e-12, rtol=1e-12)
    # sp/tests/test_signaltools.py:1848-1853
    y = O.lfilter_fir(np.array([1, 1], np.float32), np.arange(6, dtype=np.float32)[None, :])
    assert np.array_equal(y[0], [0, 1, 3, 5, 7, 9])
    # a0 normalisation
    y2 = O.lfilter_fir(np.array([2, 2], np.float32), np.arange(6, dtype=np.float32)[None, :], a0=2.0)
    assert np.array_equal(y2[0], [0, 1, 3, 5, 7, 9])


def test_fir_oracle_equals_lfilter_oracle_reversed():
    rng = np.random.RandomState(5)
    x = (rng.rand(2, 100).astype(np.float32) * 2 - 1)
    b = rng.randn(17).astype(np.float32)
    np.testing.assert_allclose(O.fir1d_batched_f32_acc64(x, b[::-1].copy()), O.lfilter_fir(b, x),
                               atol=1e-12, rtol=1e-12)


@pytest.mark.parametrize("len_h,len_x,up,down,expected", [
    (2, 2, 5, 2, [1, 0, 0, 0]), (2, 3, 6, 3, [1, 0, 1, 0, 1]), (2, 4, 4, 3, [1, 0, 0, 0, 1]),
    (3, 2, 6, 2, [1, 0, 0, 1, 0]), (4, 11, 3, 5, [1, 0, 0, 1, 0, 0, 1])])
def test_upfirdn_length_factors(len_h, len_x, up, down, expected):
    # MATLAB-derived: sp/tests/test_upfirdn.py:155-169
    h = np.zeros(len_h, np.float32); h[0] = 1
    y = O.upfirdn(h, np.ones((1, len_x), np.float32), up, down)
    assert np.array_equal(y[0], np.asarray(expected, dtype=np.float64))


def test_upfirdn_output_len():
    # sp/tests/test_upfirdn.py:311-322
    assert O.upfirdn_out_len(1001, 10**8, 320, 441) == 72562360
    from scipy.signal._upfirdn_apply import _output_len
    rng = np.random.RandomState(7)
    for _ in range(300):
        a = [int(v) for v in (rng.randint(1, 500), rng.randint(1, 10**6), rng.randint(1, 50),
                              rng.randint(1, 50))]
        assert O.upfirdn_out_len(*a) == _output_len(*a)


@pytest.mark.parametrize("len_h", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("len_x", [1, 2, 3, 4, 5])
def test_upfirdn_singleton(len_h, len_x):
    # sp/tests/test_upfirdn.py:132-141
    h = np.zeros(len_h, np.float32); h[len_h // 2] = 1
    y = O.upfirdn(h, np.ones((1, len_x), np.float32), 1, 1)[0]
    want = np.pad(np.ones(len_x), (len_h // 2, (len_h - 1) // 2))
    assert np.array_equal(y, want)


def test_upfirdn_shift_x_and_docstring_vectors():
    # sp/tests/test_upfirdn.py:143-150 and sp/_upfirdn.py:171-182
    assert np.array_equal(O.upfirdn([1, 1], [[1.0]], 1, 1)[0], [1.0, 1.0])
    assert np.array_equal(O.upfirdn([1, 1], [[0.0, 1.0]], 1, 1)[0], [0.0, 1.0, 1.0])
    assert np.array_equal(O.upfirdn([1, 1, 1], [[1, 1, 1]], 1, 1)[0], [1, 2, 3, 2, 1])
    assert np.array_equal(O.upfirdn([1], [[1, 2, 3]], 3, 1)[0], [1, 0, 0, 2, 0, 0, 3])
    assert np.array_equal(O.upfirdn([1, 1, 1], [[1, 2, 3]], 3, 1)[0], [1, 1, 1, 2, 2, 2, 3, 3, 3])
    assert np.array_equal(O.upfirdn([.5, 1, .5], [[1, 1, 1]], 2, 1)[0], [.5, 1, 1, 1, 1, 1, .5])
    assert np.array_equal(O.upfirdn([1], [np.arange(10.0)], 1, 3)[0], [0, 3, 6, 9])
    np.testing.assert_allclose(O.upfirdn([.5, 1, .5], [np.arange(10.0)], 2, 3)[0],
                               [0., 1., 2.5, 4., 5.5, 7., 8.5])


def test_upfirdn_vs_scipy_sweep(scipy_vectors):
    v = scipy_vectors
    for i, (lh, lx, up, down) in enumerate(v["upfirdn_cases"]):
        y = O.upfirdn(v[f"upfirdn_{i}_h"], v[f"upfirdn_{i}_x"], int(up), int(down))
        want = v[f"upfirdn_{i}_y64"]
        assert y.shape == want.shape
        np.testing.assert_allclose(y, want, atol=1e-12, rtol=1e-12)
        y32 = O.upfirdn(v[f"upfirdn_{i}_h"], v[f"upfirdn_{i}_x"], int(up), int(down), acc64=False)
        tol = 1e-5 * np.abs(v[f"upfirdn_{i}_h"]).sum() * np.abs(v[f"upfirdn_{i}_x"]).max()
        assert np.abs(y32 - want).max() <= tol


@pytest.mark.parametrize("down,want_len", [(2, 5015), (11, 912), (79, 127)])
def test_upfirdn_vs_convolve_matlab_lengths(scipy_vectors, down, want_len):
    # sp/tests/test_upfirdn.py:171-201
    v = scipy_vectors
    x = v["vs_convolve_x"]
    h = v[f"vs_convolve_{down}_h"]
    y = O.upfirdn(h, x[None, :], 1, down)[0]
    assert y.shape == (want_len,)
    np.testing.assert_allclose(y, v[f"vs_convolve_{down}_y64"], atol=1e-12, rtol=1e-12)
    np.testing.assert_allclose(y, np.convolve(h.astype(np.float64), x.astype(np.float64))[::down],
                               atol=1e-12, rtol=1e-12)


def test_resample_poly_plan_bit_exact_vs_scipy_source_formula():
    # re-derive the plan with SciPy's own _output_len and the arithmetic of
    # _signaltools.py:3882-3918, for many shapes, incl. BASELINE config 4
    from math import gcd
    from scipy.signal._upfirdn_apply import _output_len
    rng = np.random.RandomState(11)
    cases = [(2**20, 96, 3, 2), (32, 31, 2, 3), (32, 61, 2, 3)]
    cases += [tuple(int(v) for v in (rng.randint(1, 5000), rng.randint(1, 400), rng.randint(1, 30),
                                     rng.randint(1, 30))) for _ in range(500)]
    for n_in, len_h, up, down in cases:
        g = gcd(up, down); u, d = up // g, down // g
        n_out = n_in * u; n_out = n_out // d + bool(n_out % d)
        half_len = (len_h - 1) // 2
        n_pre_pad = d - half_len % d
        n_post_pad = 0
        n_pre_remove = (half_len + n_pre_pad) // d
        while _output_len(len_h + n_pre_pad + n_post_pad, n_in, u, d) < n_out + n_pre_remove:
            n_post_pad += 1
        p = O.resample_poly_plan(n_in, len_h, up, down)
        assert (p["up"], p["down"], p["n_out"], p["half_len"], p["n_pre_pad"], p["n_post_pad"],
                p["n_pre_remove"]) == (u, d, n_out, half_len, n_pre_pad, n_post_pad, n_pre_remove)
        assert p["len_h_padded"] == len_h + n_pre_pad + n_post_pad
    p = O.resample_poly_plan(2**20, 96, 3, 2)      # SURVEY 8(a10) probed values
    assert (p["half_len"], p["n_pre_pad"], p["n_post_pad"], p["n_pre_remove"], p["upfirdn_len"],
            p["n_out"]) == (47, 1, 0, 24, 1572911, 1572864)


def test_resample_poly_vs_scipy(scipy_vectors):
    v = scipy_vectors
    for i, (up, down, lh, n) in enumerate(v["resample_cases"]):
        h, x = v[f"resample_{i}_h"], v[f"resample_{i}_x"]
        y = O.resample_poly(x, int(up), int(down), h)
        want64, want32 = v[f"resample_{i}_y64"], v[f"resample_{i}_y32"]
        assert y.shape == want64.shape == want32.shape
        tol = 1e-5 * np.abs(h * up).sum() * np.abs(x).max()
        # SciPy-f64 scales the window in f64; ours scales in f32 like SciPy-f32: compare to both
        assert np.abs(y - want64).max() <= tol
        assert np.abs(y - want32).max() <= tol
        y32 = O.resample_poly(x, int(up), int(down), h, acc64=False)
        assert np.abs(y32 - want32).max() <= tol


@pytest.mark.parametrize("name,padtype,padlen", [
    ("filtfilt_odd", O.PAD_ODD, -1), ("filtfilt_even", O.PAD_EVEN, -1),
    ("filtfilt_const", O.PAD_CONSTANT, -1), ("filtfilt_none", O.PAD_NONE, -1),
    ("filtfilt_odd_pad10", O.PAD_ODD, 10)])
def test_filtfilt_fir_vs_scipy(scipy_vectors, name, padtype, padlen):
    v = scipy_vectors
    y = O.filtfilt_fir(v["filtfilt_b"], v["filtfilt_x"], padtype, padlen)
    np.testing.assert_allclose(y, v[name], atol=1e-10, rtol=1e-10)


def test_filtfilt_identity_and_short_input():
    # sp/tests/test_signaltools.py:2797-2804: filtfilt with b=[1] is the identity
    x = np.arange(12, dtype=np.float32)[None, :]
    np.testing.assert_allclose(O.filtfilt_fir(np.array([1.0], np.float32), x)[0], x[0], atol=1e-12)
    with pytest.raises(ValueError):
        O.filtfilt_fir(np.ones(5, np.float32), np.ones((1, 15), np.float32))   # n <= 3*ntaps


# ---- signal-extension modes of upfirdn / resample_poly ------------------------------------
@pytest.fixture(scope="module")
def scipy_modes(golden_dir):
    return np.load(os.path.join(golden_dir, "scipy_modes.npz"))


def test_oracle_upfirdn_modes_pinned_on_scipy(scipy_modes):
    """The C restatement of _extend_left / _extend_right / _apply_impl (_upfirdn_apply.pyx:110-231, :421-481)
    against SciPy's own f64 outputs for all nine modes, including inputs shorter than the filter's reach."""
    v = scipy_modes
    for i, (lh, lx, up, down) in enumerate(v["cases"]):
        h, x = v[f"u{i}_h"], v[f"u{i}_x"]
        for m in [str(s) for s in v["modes"]]:
            want = v[f"u{i}_{m}"]
            got64 = O.upfirdn_mode(h, x, int(up), int(down), m, f64=True)          # all-f64 twin: to rounding
            assert got64.shape == want.shape
            np.testing.assert_allclose(got64, want, rtol=1e-12, atol=1e-12 * np.abs(h).sum() * np.abs(want).max())
            got = O.upfirdn_mode(h, x, int(up), int(down), m)                       # f32 extension values, f64 sums
            scale = np.abs(h.astype(np.float64)).sum() * max(np.abs(want).max(), 1.0)
            assert np.abs(got - want).max() <= 1e-5 * scale, (i, m)
        got = O.upfirdn_mode(h, x, int(up), int(down), "constant", 0.75, f64=True)
        np.testing.assert_allclose(got, v[f"u{i}_constant_cval"], rtol=1e-12, atol=1e-12)
    # mode='constant', cval=0 is the path every other test uses
    h, x = v["u0_h"], v["u0_x"]
    np.testing.assert_allclose(O.upfirdn_mode(h, x, 3, 2, "constant"), O.upfirdn(h, x, 3, 2), rtol=0, atol=0)


def test_oracle_resample_poly_padtypes_pinned_on_scipy(scipy_modes):
    v = scipy_modes
    for i, (up, down, lh, n) in enumerate(v["rcases"]):
        h, x = v[f"r{i}_h"], v[f"r{i}_x"]
        for pt in [str(s) for s in v["padtypes"]]:
            want = v[f"r{i}_{pt}"]
            got = O.resample_poly_padtype(x, int(up), int(down), h, pt)
            assert got.shape == want.shape
            tol = 1e-5 * np.abs(h.astype(np.float64) * up).sum() * np.abs(x).max() * 4
            assert np.abs(got - want).max() <= tol, (i, pt)
        got = O.resample_poly_padtype(x, int(up), int(down), h, "constant", 0.5)
        assert np.abs(got - v[f"r{i}_constant_cval"]).max() <= 1e-5 * np.abs(h * up).sum() * np.abs(x).max() * 4
