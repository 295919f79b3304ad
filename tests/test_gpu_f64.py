"""GPU parity of the f64 routes: upfirdn / resample_poly / filtfilt on float64 rows, through the
C ABI (scir_b200_upfirdn_mode_f64, scir_b200_resample_poly_pad_f64, scir_b200_filtfilt_fir_f64), against the oracle's
all-f64 restatement and against SciPy's own f64 results, and the reference's f64 fixture
(crates/scir-signal/src/lib.rs:655-668) served in f64 instead of by down-casting.  Tolerance: 1e-12 * sum|h| * max|x|
(the reference's f64 convention, crates/scir-gpu/src/lib.rs:1263-1298)."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]

torch = pytest.importorskip("torch")

from oracle import oracle as O                      # noqa: E402  (checker only)
from scir_b200 import _lib as L                     # noqa: E402
from scir_b200 import gpu, signal                   # noqa: E402

MODES = ["constant", "symmetric", "edge", "smooth", "wrap", "reflect", "antisymmetric", "antireflect", "line"]


def tol64(h, x, scale=1.0):
    return 1e-12 * float(np.abs(h).sum()) * max(float(np.abs(x).max()), 1e-300) * scale


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("up,down,len_h,n", [(1, 1, 5, 40), (3, 2, 97, 5000), (2, 3, 31, 32), (160, 147, 401, 3000), (4, 7, 9, 3),
                                             (5, 1, 64, 1025), (1, 9, 200, 2049)])
def test_upfirdn_f64_vs_oracle_and_scipy(mode, up, down, len_h, n):
    from scipy.signal import upfirdn as sp_upfirdn
    rng = np.random.RandomState(up * 1000 + down * 10 + len_h)
    h = rng.randn(len_h)
    x = rng.randn(3, n)
    want = O.upfirdn_mode(h, x, up, down, mode, 0.25 if mode == "constant" else 0.0, f64=True)
    cval = 0.25 if mode == "constant" else 0.0
    for xin in (x, torch.from_numpy(x).cuda()):
        y = signal.upfirdn_f64(h, xin, up, down, mode=mode, cval=cval)
        y = y.cpu().numpy() if hasattr(y, "cpu") else y
        assert y.dtype == np.float64 and y.shape == want.shape
        scale = max(1.0, float(np.abs(want).max()) / (float(np.abs(h).sum()) * float(np.abs(x).max())))   # smooth/line grow
        assert np.abs(y - want).max() <= tol64(h, x, scale)
    sp = np.stack([sp_upfirdn(h, r, up, down, mode=mode, cval=cval) for r in x])
    assert np.abs(y - sp).max() <= tol64(h, x, 10 * scale)


@pytest.mark.parametrize("padtype", MODES + ["mean", "median", "minimum", "maximum"])
@pytest.mark.parametrize("up,down,n", [(3, 2, 4001), (2, 3, 32), (160, 147, 2000), (1, 4, 999)])
def test_resample_poly_f64_vs_scipy(padtype, up, down, n):
    from scipy.signal import resample_poly as sp_resample
    spt = {"wrap": "wrap"}.get(padtype, padtype)
    rng = np.random.RandomState(n + up)
    x = rng.randn(2, n) + 3.0
    w = signal.kaiser_lowpass(up, down)
    want = sp_resample(x, up, down, axis=-1, window=w, padtype=spt)
    for xin in (x, torch.from_numpy(x).cuda()):
        y = signal.resample_poly_f64(xin, up, down, w, padtype=padtype)
        y = y.cpu().numpy() if hasattr(y, "cpu") else y
        assert y.dtype == np.float64 and y.shape == want.shape
        assert np.abs(y - want).max() <= tol64(w * up, x, 20.0), padtype
    y1 = signal.resample_poly_f64(x[0], up, down, w, padtype=padtype)          # 1-D in, 1-D out
    assert y1.shape == (want.shape[1],) and np.array_equal(y1, y[0])


def test_reference_f64_fixture_served_in_f64(golden_dir):
    """sig/lib.rs:655-668: resample_poly(linspace(0,1,32,endpoint=False), 2, 3) == resample_poly_output.npy.  The fixture
    is SciPy's f64 output; the f64 route reproduces it to 1e-13 (the f32 route: 1e-5), and the reference's own 31-tap
    resampler (sig/lib.rs:313-362) is reproduced in f64 through the same kernel."""
    fx = os.path.join(golden_dir, "reference_fixtures")
    x = np.load(os.path.join(fx, "sosfilt_input.npy"))
    want = np.load(os.path.join(fx, "resample_poly_output.npy"))
    assert x.dtype == np.float64 and want.dtype == np.float64
    y = signal.resample_poly_f64(x, 2, 3)
    assert y.dtype == np.float64 and y.shape == want.shape == (22,)
    np.testing.assert_allclose(y, want, atol=1e-13, rtol=0)
    taps = np.load(os.path.join(golden_dir, "legacy_resample_taps.npy"))
    y_legacy = signal.resample_poly_f64(x, 2, 3, taps / 2.0)
    np.testing.assert_allclose(y_legacy, O.legacy_resample_poly_2_3(x), atol=1e-14, rtol=0)
    np.testing.assert_allclose(y_legacy, want, atol=2e-2, rtol=1e-6)               # the reference's own tolerance


@pytest.mark.parametrize("padtype,padlen", [("odd", None), ("even", None), ("constant", None), (None, None), ("odd", 10), ("even", 0),
                                            ("odd", 200)])
@pytest.mark.parametrize("k,n", [(31, 500), (255, 4000), (2, 10), (64, 20000)])
def test_filtfilt_f64_vs_scipy(padtype, padlen, k, n):
    from scipy.signal import filtfilt as sp_filtfilt, firwin
    rng = np.random.RandomState(k + n)
    b = firwin(k, 0.2) if k > 2 else np.array([0.75, 0.25])
    x = rng.randn(3, n)
    edge = 0 if padtype is None else (3 * k if padlen is None else padlen)
    if n <= edge:
        with pytest.raises(ValueError, match="padlen"):
            signal.filtfilt_f64(b, [1.0], x, padtype=padtype, padlen=padlen)
        return
    want = sp_filtfilt(b, [1.0], x, axis=-1, padtype=padtype, padlen=padlen)
    for xin in (x, torch.from_numpy(x).cuda()):
        y = signal.filtfilt_f64(b, [1.0], xin, padtype=padtype, padlen=padlen)
        y = y.cpu().numpy() if hasattr(y, "cpu") else y
        assert y.dtype == np.float64
        assert np.abs(y - want).max() <= tol64(np.convolve(b, b[::-1]), x, 3.0 * 10), (padtype, padlen)


def test_filtfilt_f64_zero_state_is_the_reference_structure():
    """sig/lib.rs:278-291: zero-state forward, reverse, zero-state forward, reverse (no padding), in f64."""
    from scipy.signal import firwin
    rng = np.random.RandomState(5)
    b32 = firwin(63, 0.3).astype(np.float32)
    x32 = (rng.rand(2, 3000).astype(np.float32) * 2 - 1)
    want = O.filtfilt_fir_nopad(b32, x32)                                    # f64 sums over f32-representable data
    y = signal.filtfilt_f64(b32.astype(np.float64), [1.0], x32.astype(np.float64), padtype="zero_state")
    assert np.abs(y - want).max() <= tol64(np.convolve(b32.astype(np.float64), b32[::-1].astype(np.float64)), x32, 10.0)
    # and SciPy's structure on the same data (odd padding) against the oracle's filtfilt
    y2 = signal.filtfilt_f64(b32.astype(np.float64), [1.0], x32.astype(np.float64))
    assert np.abs(y2 - O.filtfilt_fir(b32, x32, O.PAD_ODD, -1)).max() <= 1e-11


def test_fir_f64_non_finite_samples_stay_local():
    """the zero-padded tail of the last tap group must not turn an Inf that lies OUTSIDE an output's k-tap
    window into NaN (0 * Inf).  Same locality as the reference loop (lib.rs:1166-1184)."""
    rng = np.random.RandomState(17)
    for k in (1, 3, 5, 8, 13, 63):
        x = rng.randn(2, 5000)
        x[0, 100] = np.inf
        x[1, 4000] = np.nan
        taps = np.abs(rng.randn(k)) + 0.1
        with np.errstate(invalid="ignore"):
            want = O.fir1d_batched_f64(x, taps)
        y = gpu.fir1d_batched_f64_cuda(torch.from_numpy(x).cuda(), taps).cpu().numpy()
        assert np.array_equal(np.isnan(y), np.isnan(want)), k
        assert np.array_equal(np.isposinf(y), np.isposinf(want)), k
        fin = np.isfinite(want)
        assert np.abs(y[fin] - want[fin]).max() <= tol64(taps, x[np.isfinite(x)])


def test_mg_device_resident_shards_and_gather_to_one_device():
    """north_star (d): rows sharded over the devices with device-resident shards (no host traffic), then the OPTIONAL
    fan-in of the whole output to one device (scir_b200_mg_gather_rows_f32: peer-to-peer copies).  On a one-GPU box the
    same device is listed three times; on the 8-GPU box every device takes part."""
    import ctypes as C
    lib = L.lib()
    ndev = gpu.device_count()
    ids = list(range(ndev)) if ndev >= 2 else [0, 0, 0]
    w = len(ids)
    devs = (C.c_int * w)(*ids)
    mg = C.c_void_p()
    assert lib.scir_b200_mg_create(devs, w, C.byref(mg)) == 0, L.last_error()
    rng = np.random.RandomState(31)
    batch, n = 5 * w + 2, 40000
    x = (rng.rand(batch, n).astype(np.float32) * 2 - 1)
    taps = rng.randn(63).astype(np.float32)
    xs, ys, keep = (C.c_void_p * w)(), (C.c_void_p * w)(), []
    ld = (C.c_int64 * w)(*([n] * w))
    cur = torch.cuda.current_device()
    for s in range(w):
        a, b = C.c_int64(), C.c_int64()
        assert lib.scir_b200_shard_rows(batch, w, s, C.byref(a), C.byref(b)) == 0
        xt = torch.from_numpy(x[a.value:b.value]).to(f"cuda:{ids[s]}")
        yt = torch.empty_like(xt)
        keep += [xt, yt]
        xs[s], ys[s] = xt.data_ptr(), yt.data_ptr()
    for d in set(ids):
        torch.cuda.synchronize(d)
    rc = lib.scir_b200_mg_fir1d_batched_f32(mg, xs, ld, taps.ctypes.data, taps.size, L.TAPS_SCIR, ys, ld, batch, n)
    assert rc == 0, L.last_error()
    assert torch.cuda.current_device() == cur                      # no ABI call moves the caller's device
    dst = w - 1
    full = torch.empty((batch, n + 4), dtype=torch.float32, device=f"cuda:{ids[dst]}")     # padded pitch
    rc = lib.scir_b200_mg_gather_rows_f32(mg, ys, ld, dst, C.c_void_p(full.data_ptr()), n + 4, batch, n)
    assert rc == 0, L.last_error()
    got = full[:, :n].cpu().numpy()
    want = O.fir1d_batched_f32_acc64(x, taps)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(taps).sum() * np.abs(x).max()
    assert lib.scir_b200_mg_gather_rows_f32(mg, ys, ld, w, C.c_void_p(full.data_ptr()), n + 4, batch, n) == L.ERR_INVALID_ARG
    lib.scir_b200_mg_destroy(mg)
    assert torch.cuda.current_device() == cur


def test_abi_calls_do_not_move_the_current_device():
    """ctx_bind used to leave the ctx's device current.  A ctx on another device -- or, on a
    one-GPU box, any ctx -- must leave cudaGetDevice / torch's current device where the caller had it."""
    ndev = gpu.device_count()
    other = 1 if ndev >= 2 else 0
    torch.cuda.set_device(0)
    ctx = gpu.Context(other)
    x = (np.random.RandomState(0).rand(2, 1000).astype(np.float32))
    y = gpu.fir1d_batched_f32_cuda(x, np.ones(3, np.float32), ctx=ctx)
    assert torch.cuda.current_device() == 0 and gpu.current_device() == 0
    assert y.shape == x.shape
    if ndev >= 2:
        t = torch.ones((2, 64), device="cuda:1")
        z = gpu.fir1d_batched_f32_cuda(t, np.ones(3, np.float32))
        assert z.device.index == 1 and torch.cuda.current_device() == 0
        # a fresh tensor after the call still lands on the caller's device
        assert torch.empty(1, device="cuda").device.index == 0
