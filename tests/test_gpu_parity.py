"""GPU parity: the CUDA path, called through the C ABI (scir_b200.gpu / scir_b200.signal ->
ctypes -> libscir_b200.so), against the CPU oracle on the same seeded inputs and against the
committed golden vectors.

Tolerance (BASELINE.json north_star): max|err| <= 1e-5 * sum|h| * max|x| against the f64 judge;
shapes and integer plans bit-exact.  Every test asserts VALUES (the reference's CUDA test passed on
Err and its bench compared only shapes: gpu/lib.rs:1319-1321, fir_bench.rs:75).
"""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import oracle as O                      # noqa: E402  (checker only)
from scir_b200 import _lib as L                     # noqa: E402
from scir_b200 import gpu, signal                   # noqa: E402
from scir_b200.gpu import Device                    # noqa: E402


from parity_util import assert_filtfilt_close, tol   # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_native_library_is_loaded_and_device_is_b200():
    assert gpu.device_count() >= 1
    assert os.path.exists(L.LIB_PATH)
    ctx = gpu.Context(0)
    assert ctx.launch_count() == 0                                       # a fresh ctx has launched nothing ...
    y = gpu.fir1d_batched_f32_cuda(dev(np.ones((2, 64), np.float32)), np.ones(3, np.float32), ctx=ctx)
    ctx.sync()
    assert ctx.launch_count() >= 1 and float(y[0, 5]) == 3.0             # ... and the FIR call is a real kernel launch
    assert torch.cuda.get_device_capability(0)[0] == 10
    loaded = open("/proc/self/maps").read()
    assert "libscir_b200.so" in loaded                                   # the in-tree native library is what ran


# ---- reference known-answer vector, gpu/lib.rs:1300-1323 (cuda_fir1d_batched_f32_parity_small) ----
X24 = np.array([[1.0, 2.0, 3.0, 4.0], [0.5, 0.0, -0.5, -1.0]], dtype=np.float32)
TAPS3 = np.array([0.25, 0.5, 0.25], dtype=np.float32)
Y24 = np.array([[0.25, 1.0, 2.0, 3.0], [0.125, 0.25, 0.0, -0.5]], dtype=np.float64)


def test_reference_golden_host_path():
    y = gpu.fir1d_batched_f32_cuda(X24, TAPS3)               # Result::Ok expected: Err is a failure here
    assert y.shape == (2, 4) and y.dtype == np.float32
    np.testing.assert_allclose(y.astype(np.float64), Y24, atol=1e-7, rtol=1e-7)      # lib.rs:1259-1260
    y_cpu = O.fir1d_batched_f32(X24, TAPS3)
    np.testing.assert_allclose(y, y_cpu, atol=1e-5, rtol=1e-6)                        # lib.rs:1317
    y2 = signal.gpu.fir1d_batched_f32(X24, TAPS3, Device.Cuda)                        # sig/lib.rs:372
    assert np.array_equal(y, y2)
    y3 = gpu.fir1d_batched_f32_auto(X24, TAPS3, Device.Cuda)
    assert np.array_equal(y, y3)


def test_reference_golden_device_path():
    y = gpu.fir1d_batched_f32_cuda(dev(X24), TAPS3)
    assert y.is_cuda and tuple(y.shape) == (2, 4)
    np.testing.assert_allclose(y.cpu().numpy().astype(np.float64), Y24, atol=1e-7, rtol=1e-7)


SHAPES = [  # (batch, n, k)
    (1, 1, 1), (1, 1, 5), (3, 2, 7), (2, 4, 3), (3, 32, 5), (5, 100, 31), (4, 1000, 32), (4, 1001, 33),
    (2, 5119, 63), (2, 5120, 63), (2, 5121, 64), (3, 10240, 65), (2, 20000, 127), (2, 20003, 255),
    (2, 12345, 256), (1, 30001, 257), (2, 9000, 1000), (1, 40000, 4097), (64, 16384, 31), (7, 16385, 63),
]


@pytest.mark.parametrize("batch,n,k", SHAPES)
def test_fir_vs_oracle_device(batch, n, k):
    rng = np.random.RandomState(batch * 7919 + n * 31 + k)
    x = (rng.rand(batch, n).astype(np.float32) * 2 - 1)
    taps = rng.randn(k).astype(np.float32)
    want = O.fir1d_batched_f32_acc64(x, taps)
    y = gpu.fir1d_batched_f32_cuda(dev(x), taps).cpu().numpy()
    assert y.shape == x.shape
    err = np.abs(y - want).max()
    assert err <= tol(taps, x), (err, tol(taps, x))
    # the reference-order f32 CPU result is itself within tolerance of the judge; report both
    ref_err = np.abs(O.fir1d_batched_f32(x, taps) - want).max()
    assert ref_err <= tol(taps, x) * max(1.0, k / 64.0)


@pytest.mark.parametrize("batch,n,k", [(3, 777, 31), (2, 6000, 63), (16, 4096, 255), (1, 11000, 1500)])
def test_fir_vs_oracle_host_path(batch, n, k):
    rng = np.random.RandomState(n + k)
    x = (rng.rand(batch, n).astype(np.float32) * 2 - 1)
    taps = (1.0 / (np.arange(k, dtype=np.float32) + 1.0)).astype(np.float32)       # fir_bench.rs:14-18
    want = O.fir1d_batched_f32_acc64(x, taps)
    y = gpu.fir1d_batched_f32_cuda(x, taps)
    assert np.abs(y - want).max() <= tol(taps, x)


def test_fir_bench_default_config_full():
    """BASELINE configs[0]: fir_bench defaults 64 x 16384, k=31, taps 1/(i+1), U[-1,1), seed 42."""
    rng = np.random.RandomState(42)
    x = (rng.rand(64, 1 << 14).astype(np.float32) * 2 - 1)
    taps = (1.0 / (np.arange(31, dtype=np.float32) + 1.0)).astype(np.float32)
    want = O.fir1d_batched_f32_acc64(x, taps)
    for y in (gpu.fir1d_batched_f32_cuda(x, taps), gpu.fir1d_batched_f32_cuda(dev(x), taps).cpu().numpy()):
        assert y.shape == (64, 1 << 14)
        assert np.abs(y - want).max() <= tol(taps, x)


def test_empty_and_degenerate_shapes():
    taps = np.array([1.0, 2.0, 3.0], np.float32)
    assert gpu.fir1d_batched_f32_cuda(np.zeros((0, 5), np.float32), taps).shape == (0, 5)
    assert gpu.fir1d_batched_f32_cuda(np.zeros((3, 0), np.float32), taps).shape == (3, 0)
    assert tuple(gpu.fir1d_batched_f32_cuda(torch.zeros((0, 8), device="cuda"), taps).shape) == (0, 8)
    y = gpu.fir1d_batched_f32_cuda(np.array([[2.0]], np.float32), taps)
    assert y[0, 0] == 6.0                                  # last tap times newest sample
    with pytest.raises(ValueError):
        gpu.fir1d_batched_f32_cuda(np.zeros((1, 4), np.float32), np.zeros(0, np.float32))
    with pytest.raises(gpu.GpuError):
        gpu.fir1d_batched_f32_cuda(np.zeros((1, 4), np.float32), np.zeros(L.MAX_TAPS + 1, np.float32))
    with pytest.raises(gpu.GpuError):
        gpu.fir1d_batched_f32_cuda(np.zeros(4, np.float32), taps)         # not 2-D: ShapeMismatch


def test_unaligned_and_strided_device_views():
    """Row pitch != n and 4-byte-aligned-only bases force the generic tile IO path."""
    rng = np.random.RandomState(3)
    big = (rng.rand(5, 12007).astype(np.float32) * 2 - 1)
    taps = rng.randn(63).astype(np.float32)
    xb = dev(big)
    for view, ref in ((xb[:, 1:], big[:, 1:]), (xb[:, 3:9000], big[:, 3:9000]), (xb[::2, :], big[::2, :]),
                      (xb[:, 4:], big[:, 4:])):
        y = gpu.fir1d_batched_f32_cuda(view, taps).cpu().numpy()
        want = O.fir1d_batched_f32_acc64(np.ascontiguousarray(ref), taps)
        assert np.abs(y - want).max() <= tol(taps, big)
    out = torch.zeros((5, 12007 + 5), device="cuda")[:, 1:12008]          # unaligned OUTPUT rows
    gpu.fir1d_batched_f32_cuda(xb, taps, out=out)
    assert np.abs(out.cpu().numpy() - O.fir1d_batched_f32_acc64(big, taps)).max() <= tol(taps, big)


@pytest.mark.parametrize("k,variant", [(31, 0), (63, 0), (255, 0), (700, 0), (63, 4), (255, 4),
                                       (700, 3), (31, 3)])
def test_stream_kernel_many_tiles_per_cta(k, variant):
    """More tiles than resident CTAs (148 SMs x 3): every persistent CTA walks several tiles, both
    pipeline stages wrap, edge tiles (non-bulk) interleave with bulk tiles.  Checked against the
    independent naive kernel everywhere and against the oracle on a row subset."""
    rng = np.random.RandomState(k)
    batch, n = 96, 150000                                   # many tiles per warp / CTA
    x = (rng.rand(batch, n).astype(np.float32) * 2 - 1)
    taps = rng.randn(k).astype(np.float32)
    xd = dev(x)
    main = gpu.Context(0)
    main.set_option("variant", variant)
    torch.cuda.synchronize()
    y = gpu.fir1d_batched_f32_cuda(xd, taps, ctx=main)
    main.sync()
    naive = gpu.Context(0)
    naive.set_option("variant", 2)
    torch.cuda.synchronize()
    y2 = gpu.fir1d_batched_f32_cuda(xd, taps, ctx=naive)
    naive.sync()
    assert float((y - y2).abs().max()) <= 2 * tol(taps, x)
    rows = [0, 1, 47, 95]
    want = O.fir1d_batched_f32_acc64(x[rows], taps)
    assert np.abs(y.cpu().numpy()[rows] - want).max() <= tol(taps, x)
    # twice in a row on the same ctx/stream: no state leaks between launches
    y3 = gpu.fir1d_batched_f32_cuda(xd, taps, ctx=main)
    main.sync()
    assert torch.equal(y, y3)


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_kernel_variants_agree(variant):
    """variant 1 = generic (non-bulk) IO, 2 = naive 1-thread/output kernel, 3 = one-tile-per-CTA kernel,
    4 = CTA-streaming kernel (scalar FFMA core)."""
    rng = np.random.RandomState(5)
    x = (rng.rand(3, 15000).astype(np.float32) * 2 - 1)
    taps = rng.randn(100).astype(np.float32)
    want = O.fir1d_batched_f32_acc64(x, taps)
    ctx = gpu.Context(0)
    ctx.set_option("variant", variant)
    assert ctx.get_option("variant") == variant
    xd = dev(x)
    torch.cuda.synchronize()
    y = gpu.fir1d_batched_f32_cuda(xd, taps, ctx=ctx)
    ctx.sync()
    assert np.abs(y.cpu().numpy() - want).max() <= tol(taps, x)
    assert ctx.launch_count() >= 1
    with pytest.raises(ValueError):
        ctx.set_option("no_such_option", 1)


def test_impulse_linearity_and_dc_properties_at_scale():
    """Size-independent properties on a BASELINE-config-2-shaped slab (rows reduced to keep the
    oracle out of it): impulse response == taps, linearity, DC gain, time invariance."""
    n, k = 1 << 20, 63
    from scipy.signal import firwin
    b = firwin(k, 0.25).astype(np.float32)
    taps = b[::-1].copy()                                   # kernel order = reversed lfilter order
    x = torch.zeros((4, n), device="cuda")
    pos = [0, 5119, 5120, n - 1]
    for r, p in enumerate(pos):
        x[r, p] = 1.0
    direct = gpu.Context(0)
    direct.set_option("long_tap_path", 1)                  # FP32 direct family
    y = gpu.fir1d_batched_f32_cuda(x, taps, ctx=direct)
    direct.sync()
    y = y.cpu().numpy()
    # the launch fills the GPU, so auto dispatch takes the tcgen05 path: taps are reproduced to the split's
    # 22 bits (x = 1 is exact in FP16, c = ch + cm + O(2^-22 c)), far inside the path's tolerance
    yt = gpu.fir1d_batched_f32_cuda(x, taps).cpu().numpy()
    for r, p in enumerate(pos):
        want = np.zeros(n, np.float32)
        seg = b[: max(0, min(k, n - p))]
        want[p:p + seg.size] = seg
        assert np.array_equal(y[r], want), r               # exact: one product per output
        assert np.abs(yt[r] - want).max() <= 2.0 ** -21 * np.abs(b).max(), r
        assert np.array_equal(yt[r] != 0, want != 0) or np.abs(yt[r][want == 0]).max() == 0.0
    g = torch.Generator(device="cuda").manual_seed(42)
    a = torch.rand((8, n), device="cuda", generator=g) * 2 - 1
    c = torch.rand((8, n), device="cuda", generator=g) * 2 - 1
    ya, yc = gpu.fir1d_batched_f32_cuda(a, taps), gpu.fir1d_batched_f32_cuda(c, taps)
    ys = gpu.fir1d_batched_f32_cuda(2.0 * a - 3.0 * c, taps)
    t = 1e-5 * float(np.abs(b).sum()) * 5.0
    assert float((ys - (2.0 * ya - 3.0 * yc)).abs().max()) <= 2 * t
    ones = torch.ones((2, n), device="cuda")
    yo = gpu.fir1d_batched_f32_cuda(ones, taps).cpu().numpy()
    assert np.abs(yo[:, k:] - b.astype(np.float64).sum()).max() <= 1e-5 * np.abs(b).sum()
    np.testing.assert_allclose(yo[0, :k], np.cumsum(b.astype(np.float64)), atol=1e-5 * np.abs(b).sum())
    shifted = torch.zeros_like(a)
    shifted[:, 1000:] = a[:, :-1000]
    ysh = gpu.fir1d_batched_f32_cuda(shifted, taps)
    assert float((ysh[:, 1000:] - ya[:, :-1000]).abs().max()) <= 2 * t


def test_config2_rows_vs_oracle():
    """BASELINE configs[1] shape (1M samples, 63 taps, lfilter a=[1]): full rows against the oracle."""
    from scipy.signal import firwin
    n = 1 << 20
    b = firwin(63, 0.25).astype(np.float32)
    rng = np.random.RandomState(42)
    x = (rng.rand(6, n).astype(np.float32) * 2 - 1)
    y = signal.lfilter(b, [1.0], dev(x)).cpu().numpy()
    want = O.lfilter_fir(b, x)
    assert np.abs(y - want).max() <= tol(b, x)


# ---- lfilter -------------------------------------------------------------------------------------------
def test_lfilter_golden(scipy_vectors):
    v = scipy_vectors
    b, x = v["lfilter_b"], v["lfilter_x"]
    for xin in (x, dev(x)):
        y = signal.lfilter(b, np.ones(1, np.float32), xin)
        y = y.cpu().numpy() if hasattr(y, "cpu") else y
        assert np.abs(y - v["lfilter_y64"]).max() <= tol(b, x)
        assert np.abs(y - v["lfilter_y"]).max() <= tol(b, x)
    y = signal.lfilter([1, 1], [1], np.arange(6, dtype=np.float32))         # test_signaltools.py:1848-1853
    assert np.array_equal(y, [0, 1, 3, 5, 7, 9])
    y = signal.lfilter([2, 2], [2], np.arange(6, dtype=np.float32))
    assert np.array_equal(y, [0, 1, 3, 5, 7, 9])
    with pytest.raises(gpu.GpuError):
        signal.lfilter([1, 1], [1, 0.5], np.arange(6, dtype=np.float32))    # IIR is not this path


def test_lfilter_streaming_state(scipy_vectors):
    v = scipy_vectors
    b, x, zi = v["lfilter_b"], v["lfilter_x"], v["lfilter_zi"]
    for xin, ziin in ((x, zi), (dev(x), dev(zi))):
        y, zf = signal.lfilter(b, [1.0], xin, zi=ziin)
        y = y.cpu().numpy() if hasattr(y, "cpu") else y
        zf = zf.cpu().numpy() if hasattr(zf, "cpu") else zf
        assert np.abs(y - v["lfilter_y_zi"]).max() <= tol(b, x) + 1e-6
        assert np.abs(zf - v["lfilter_zf"]).max() <= tol(b, x) + 1e-6
    # chunked filtering with carried state == one-shot filtering (the point of zi/zf)
    rng = np.random.RandomState(9)
    xs = (rng.rand(2, 3000).astype(np.float32) * 2 - 1)
    bb = rng.randn(40).astype(np.float32)
    whole = signal.lfilter(bb, [1.0], dev(xs)).cpu().numpy()
    state = torch.zeros((2, 39), device="cuda")
    parts = []
    for lo, hi in ((0, 17), (17, 1000), (1000, 1020), (1020, 3000)):          # incl. chunks shorter than k-1
        yp, state = signal.lfilter(bb, [1.0], dev(xs[:, lo:hi]), zi=state)
        parts.append(yp.cpu().numpy())
    assert np.abs(np.concatenate(parts, axis=1) - whole).max() <= 2 * tol(bb, xs)


# ---- upfirdn / resample_poly ---------------------------------------------------------------------------
def test_upfirdn_golden_sweep(scipy_vectors):
    v = scipy_vectors
    for i, (lh, lx, up, down) in enumerate(v["upfirdn_cases"]):
        h, x = v[f"upfirdn_{i}_h"], v[f"upfirdn_{i}_x"]
        want = v[f"upfirdn_{i}_y64"]
        for xin in (x, dev(x)):
            y = signal.upfirdn(h, xin, int(up), int(down))
            y = y.cpu().numpy() if hasattr(y, "cpu") else y
            assert y.shape == want.shape                                    # integer length logic: exact
            assert np.abs(y - want).max() <= tol(h, x)


@pytest.mark.parametrize("len_h,len_x,up,down,expected", [
    (2, 2, 5, 2, [1, 0, 0, 0]), (2, 3, 6, 3, [1, 0, 1, 0, 1]), (2, 4, 4, 3, [1, 0, 0, 0, 1]),
    (3, 2, 6, 2, [1, 0, 0, 1, 0]), (4, 11, 3, 5, [1, 0, 0, 1, 0, 0, 1])])
def test_upfirdn_length_factors(len_h, len_x, up, down, expected):
    h = np.zeros(len_h, np.float32); h[0] = 1                               # test_upfirdn.py:155-169
    y = signal.upfirdn(h, np.ones(len_x, np.float32), up, down)
    assert np.array_equal(y, np.asarray(expected, np.float32))


@pytest.mark.parametrize("down,want_len", [(2, 5015), (11, 912), (79, 127)])
def test_upfirdn_vs_convolve(scipy_vectors, down, want_len):
    v = scipy_vectors                                                       # test_upfirdn.py:171-201
    x, h = v["vs_convolve_x"], v[f"vs_convolve_{down}_h"]
    y = signal.upfirdn(h, dev(x), 1, down).cpu().numpy()
    assert y.shape == (want_len,)
    assert np.abs(y - v[f"vs_convolve_{down}_y64"]).max() <= tol(h, x)


@pytest.mark.parametrize("up,down,len_h,n,batch", [
    (3, 2, 97, 40000, 3), (3, 2, 96, 7777, 2), (2, 3, 31, 9001, 2), (1, 2, 63, 30000, 2), (4, 1, 33, 5000, 2),
    (5, 7, 121, 12000, 2), (7, 5, 64, 3000, 1), (1, 1, 63, 12000, 2), (2, 1, 255, 9000, 1), (160, 147, 801, 5000, 1),
    (3, 8, 50, 20000, 2), (9, 4, 200, 6000, 1), (3, 2, 5, 100, 1), (13, 11, 40, 997, 2)])
def test_upfirdn_vs_oracle_random(up, down, len_h, n, batch):
    rng = np.random.RandomState(up * 100 + down + len_h)
    h = rng.randn(len_h).astype(np.float32)
    x = (rng.rand(batch, n).astype(np.float32) * 2 - 1)
    want = O.upfirdn(h, x, up, down)
    y = signal.upfirdn(h, dev(x), up, down).cpu().numpy()
    assert y.shape == want.shape
    assert np.abs(y - want).max() <= tol(h, x)
    # generic kernel agrees too
    ctx = gpu.Context(0)
    ctx.set_option("upfirdn_variant", 1)
    xd = dev(x)
    torch.cuda.synchronize()
    y1 = signal.upfirdn(h, xd, up, down, ctx=ctx)
    ctx.sync()
    assert np.abs(y1.cpu().numpy() - want).max() <= tol(h, x)


RATES = [(3, 2), (2, 3), (2, 1), (1, 2), (3, 1), (1, 3), (4, 1), (1, 4), (4, 3), (3, 4)]


@pytest.mark.parametrize("up,down", RATES)
@pytest.mark.parametrize("lead,trail,len_h", [(0, 0, 33), (1, 0, 97), (2, 3, 64), (3, 1, 200), (5, 0, 31), (0, 0, 3)])
def test_upfirdn_tile_kernel_rates(up, down, lead, trail, len_h):
    """Every templated rate of the polyphase tile kernel x leading/trailing structural zeros
    (resample_poly's padding) x 1..3 tap chunks, several tiles per row, vs the oracle; the generic
    one-thread-per-output kernel and the non-bulk tile IO path must agree."""
    rng = np.random.RandomState(up * 1000 + down * 100 + lead * 10 + len_h)
    h = rng.randn(len_h).astype(np.float32)
    h[:lead] = 0
    if trail:
        h[-trail:] = 0
    n = 30011
    x = (rng.rand(3, n).astype(np.float32) * 2 - 1)
    want = O.upfirdn(h, x, up, down)
    xd = dev(x)
    y = signal.upfirdn(h, xd, up, down).cpu().numpy()
    assert y.shape == want.shape
    assert np.abs(y - want).max() <= tol(h, x)
    for variant in (1, 2):
        ctx = gpu.Context(0)
        ctx.set_option("upfirdn_variant", variant)
        torch.cuda.synchronize()
        yv = signal.upfirdn(h, xd, up, down, ctx=ctx)
        ctx.sync()
        assert np.abs(yv.cpu().numpy() - want).max() <= tol(h, x)


@pytest.mark.parametrize("variant", [3, 4, 5, 6, 7, 8, 9, 0, 20])
@pytest.mark.parametrize("up,down,len_h", [(3, 2, 97), (2, 3, 61), (1, 2, 33), (4, 3, 40), (1, 4, 120), (3, 1, 90), (2, 1, 64)])
def test_upfirdn_kernel_variants_agree(variant, up, down, len_h):
    """upfirdn_variant: 0 warp-specialised + tap-reuse FFMA2 core (default; 20 = the same with 2 input stages),
    7 streaming + tap-reuse core, 8 streaming + one tap-pair load per FFMA2 (round-1 default), 3 one tile per CTA (FFMA2),
    4 streaming scalar FFMA, 5 tile scalar, 6 per-warp pipelines: same results to rounding, all inside the oracle
    tolerance, on aligned and unaligned views, many more tiles than resident CTAs (every stage and barrier phase cycles)."""
    if variant == 20:
        variant, stages = 0, 2
    else:
        stages = 3
    rng = np.random.RandomState(variant * 10 + up)
    h = rng.randn(len_h).astype(np.float32)
    x = (rng.rand(5, 70001).astype(np.float32) * 2 - 1)
    want = O.upfirdn(h, x, up, down)
    ctx = gpu.Context(0)
    ctx.set_option("upfirdn_variant", variant)
    ctx.set_option("upfirdn_ws_stages", stages)
    for view in (dev(x), dev(x)[:, 1:], dev(x)[::2, 3:60000]):
        y = signal.upfirdn(h, view, up, down, ctx=ctx)
        ctx.sync()
        w = O.upfirdn(h, view.cpu().numpy(), up, down)
        assert y.shape == w.shape
        assert np.abs(y.cpu().numpy() - w).max() <= tol(h, x)
    assert want.shape[0] == 5
    # a launch with far more tiles than resident CTAs, and an extension mode (edge tiles synthesised by the producer)
    big = (np.random.RandomState(up + down).rand(64, 400003).astype(np.float32) * 2 - 1)
    yb = signal.upfirdn(h, dev(big), up, down, mode="symmetric", ctx=ctx)
    ctx.sync()
    rows = [0, 31, 63]
    wb = O.upfirdn_mode(h, big[rows], up, down, "symmetric")
    assert np.abs(yb[rows].cpu().numpy() - wb).max() <= tol(h, big)


@pytest.mark.parametrize("up,down,len_h,n", [(160, 147, 3201, 30000), (147, 160, 3201, 30011), (5, 4, 101, 50000), (7, 3, 141, 20000),
                                            (2, 5, 101, 40000), (441, 320, 2001, 9000), (5, 1, 33, 7000), (1, 7, 50, 60000),
                                            (11, 13, 7, 5000)])
def test_upfirdn_any_rate_tiled_kernel(up, down, len_h, n):
    """Rates outside the template grid run on the tiled any-rate kernel (phase-transposed taps + input span in
    shared memory), not on the one-thread-per-output fallback; values against the oracle, with and without a mode,
    and equal to the fallback kernel's to rounding."""
    rng = np.random.RandomState(up * 7 + down)
    h = rng.randn(len_h).astype(np.float32)
    x = (rng.rand(3, n).astype(np.float32) * 2 - 1)
    ctx = gpu.Context(0)
    g0 = ctx.get_option("gen_tiled_launches")
    y = signal.upfirdn(h, dev(x), up, down, ctx=ctx)
    ctx.sync()
    assert ctx.get_option("gen_tiled_launches") == g0 + 1
    want = O.upfirdn(h, x, up, down)
    assert y.shape == want.shape
    assert np.abs(y.cpu().numpy() - want).max() <= tol(h, x)
    ym = signal.upfirdn(h, dev(x), up, down, mode="reflect", ctx=ctx)
    ctx.sync()                                             # the ctx owns its stream: torch's .cpu() does not wait on it
    ym = ym.cpu().numpy()
    wm = O.upfirdn_mode(h, x, up, down, "reflect")
    assert np.abs(ym - wm).max() <= tol(h, x, 2.0)
    old = gpu.Context(0)
    old.set_option("upfirdn_variant", 1)                   # one thread per output
    y1 = signal.upfirdn(h, dev(x), up, down, ctx=old)
    old.sync()
    assert old.get_option("gen_tiled_launches") == 0
    assert float((y1 - y).abs().max()) <= 2 * tol(h, x)
    # resample_poly window (n_pre_remove offset) through the same kernel
    w = signal.kaiser_lowpass(up, down) if max(up, down) <= 160 else h
    yr = signal.resample_poly(dev(x), up, down, w, ctx=ctx)
    ctx.sync()
    assert np.abs(yr.cpu().numpy() - O.resample_poly(x, up, down, w)).max() <= tol(w * up, x, 2.0)


def test_upfirdn_tile_kernel_is_the_one_that_runs():
    """The templated rates must be served by ONE tile-kernel launch (not the generic fallback)."""
    rng = np.random.RandomState(1)
    h = rng.randn(97).astype(np.float32)
    xd = dev(rng.rand(4, 100000).astype(np.float32))
    ctx = gpu.torch_context(xd)
    n0, p0 = ctx.launch_count(), ctx.get_option("poly_launches")
    signal.upfirdn(h, xd, 3, 2)
    assert ctx.launch_count() == n0 + 1 and ctx.get_option("poly_launches") == p0 + 1
    signal.upfirdn(h, xd, 160, 147)                         # not in the template grid: generic kernel
    assert ctx.get_option("poly_launches") == p0 + 1


# ---- signal-extension modes (_upfirdn_apply.pyx:110-231) -----------------------------------
MODES = ["constant", "symmetric", "edge", "smooth", "wrap", "reflect", "antisymmetric", "antireflect", "line"]


@pytest.fixture(scope="module")
def scipy_modes(golden_dir):
    return np.load(os.path.join(golden_dir, "scipy_modes.npz"))


def test_upfirdn_modes_golden(scipy_modes):
    """All nine modes against SciPy's outputs (committed golden vectors), host and device arrays, including inputs
    shorter than the filter's reach (multiple reflections)."""
    v = scipy_modes
    for i, (lh, lx, up, down) in enumerate(v["cases"]):
        h, x = v[f"u{i}_h"], v[f"u{i}_x"]
        for m in MODES:
            want = v[f"u{i}_{m}"]
            scale = max(np.abs(want).max() / max(np.abs(h).sum(), 1e-30), np.abs(x).max())     # extension can exceed max|x|
            for xin in (x, dev(x)):
                y = signal.upfirdn(h, xin, int(up), int(down), mode=m)
                y = y.cpu().numpy() if hasattr(y, "cpu") else y
                assert y.shape == want.shape and y.dtype == np.float32
                assert np.abs(y - want).max() <= tol(h, np.asarray([scale]), 2.0), (i, m)
        y = signal.upfirdn(h, dev(x), int(up), int(down), mode="constant", cval=0.75).cpu().numpy()
        assert np.abs(y - v[f"u{i}_constant_cval"]).max() <= tol(h, x, 2.0)
    with pytest.raises(ValueError):
        signal.upfirdn(np.ones(3, np.float32), np.ones(5, np.float32), 1, 1, mode="bogus")


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("up,down,len_h,n", [(3, 2, 97, 30000), (2, 3, 64, 40001), (1, 2, 33, 25000), (5, 3, 41, 9000),
                                            (4, 1, 40, 12000)])
def test_upfirdn_modes_vs_oracle_many_tiles(mode, up, down, len_h, n):
    """Tile kernels (rates in the template grid) and the generic kernel (5/3) at sizes with interior tiles: only the
    edge tiles may see the mode."""
    rng = np.random.RandomState(up * 100 + down + len_h)
    h = rng.randn(len_h).astype(np.float32)
    x = (rng.rand(3, n).astype(np.float32) * 2 - 1)
    want = O.upfirdn_mode(h, x, up, down, mode, 0.0)
    y = signal.upfirdn(h, dev(x), up, down, mode=mode).cpu().numpy()
    assert y.shape == want.shape
    scale = max(np.abs(want).max() / np.abs(h).sum(), 1.0)
    assert np.abs(y - want).max() <= tol(h, np.asarray([scale]), 2.0)
    # interior outputs do not depend on the mode at all
    y0 = signal.upfirdn(h, dev(x), up, down).cpu().numpy()
    lo, hi = (len_h // down) + 2, y.shape[1] - (len_h // down) - 2 - (len_h * 1) // down
    assert np.array_equal(y[:, lo:hi], y0[:, lo:hi])


def test_resample_poly_padtypes_golden(scipy_modes):
    v = scipy_modes
    for i, (up, down, lh, n) in enumerate(v["rcases"]):
        h, x = v[f"r{i}_h"], v[f"r{i}_x"]
        for pt in [str(s) for s in v["padtypes"]]:
            want = v[f"r{i}_{pt}"]
            for xin in (x, dev(x)):
                y = signal.resample_poly(xin, int(up), int(down), h, padtype=pt)
                y = y.cpu().numpy() if hasattr(y, "cpu") else y
                assert y.shape == want.shape
                assert np.abs(y - want).max() <= tol(h * up, x, 4.0), (i, pt)
        y = signal.resample_poly(dev(x), int(up), int(down), h, padtype="constant", cval=0.5).cpu().numpy()
        assert np.abs(y - v[f"r{i}_constant_cval"]).max() <= tol(h * up, x, 4.0)
    with pytest.raises(ValueError):
        signal.resample_poly(np.ones(8, np.float32), 2, 1, np.ones(5, np.float32), padtype="edge", cval=1.0)
    with pytest.raises(ValueError):
        signal.resample_poly(np.ones(8, np.float32), 2, 1, np.ones(5, np.float32), padtype="nope")


def test_resample_poly_stat_padtypes_at_scale():
    rng = np.random.RandomState(77)
    from scipy.signal import firwin
    h = firwin(96, 1.0 / 3.0, window=("kaiser", 5.0)).astype(np.float32)
    x = (rng.rand(5, 50000).astype(np.float32) + 2.0)
    x[2, ::3] -= 5.0                                         # negative values, ties and an even / odd split below
    x[3, :1000] = 2.5
    for pt in ("mean", "median", "minimum", "maximum", "line", "edge"):
        want = O.resample_poly_padtype(x, 3, 2, h, pt)
        y = signal.resample_poly(dev(x), 3, 2, h, padtype=pt).cpu().numpy()
        assert np.abs(y - want).max() <= tol(h * 3, x, 4.0), pt


@pytest.mark.parametrize("n", [1, 2, 3, 10, 1001, 4096, 100000])
def test_row_median_is_exact(n):
    """The median behind padtype='median' is an exact order statistic (radix select), including negative values,
    ties, zeros of both signs and even lengths (f32 mean of the two middle values, like numpy).  A zero filter
    makes resample_poly return the statistic itself: y = upfirdn(0, x - med) + med = med."""
    rng = np.random.RandomState(n)
    x = rng.randn(6, n).astype(np.float32)
    x[1] = np.round(x[1] * 2) / 2                           # many ties
    x[2] = -np.abs(x[2])
    x[3, : n // 2] = 0.0
    x[4, ::2] = -0.0
    x[5] *= 1e30
    zero = np.zeros(3, np.float32)
    y = signal.resample_poly(dev(x), 2, 1, zero, padtype="median").cpu().numpy()
    want = np.median(x, axis=1).astype(np.float32)
    assert y.shape == (6, 2 * n)
    for r in range(6):
        assert np.all(y[r] == want[r]), (r, y[r][:3], want[r])
    for pt, fn in (("minimum", np.amin), ("maximum", np.amax)):
        y = signal.resample_poly(dev(x), 2, 1, zero, padtype=pt).cpu().numpy()
        assert np.array_equal(y[:, 0], fn(x, axis=1))


def test_upfirdn_windowed_output():
    rng = np.random.RandomState(12)
    h = rng.randn(97).astype(np.float32)
    x = (rng.rand(2, 9000).astype(np.float32) * 2 - 1)
    want = O.upfirdn(h, x, 3, 2)
    xd = dev(x)
    lib = L.lib()
    ctx = gpu.torch_context(xd)
    for m0, cnt in ((0, 10), (24, 13476), (5, 4001), (want.shape[1] - 7, 7)):
        y = torch.full((2, cnt + 3), -7.0, device="cuda")
        rc = lib.scir_b200_upfirdn_f32(ctx.handle, h.ctypes.data, h.size, 3, 2, xd.data_ptr(), xd.stride(0), 2, 9000,
                                       y.data_ptr(), y.stride(0), m0, cnt)
        assert rc == 0, L.last_error()
        yh = y.cpu().numpy()
        assert np.abs(yh[:, :cnt] - want[:, m0:m0 + cnt]).max() <= tol(h, x)
        assert np.all(yh[:, cnt:] == -7.0)                                   # nothing written past the window
    rc = lib.scir_b200_upfirdn_f32(ctx.handle, h.ctypes.data, h.size, 3, 2, xd.data_ptr(), xd.stride(0), 2, 9000,
                                   xd.data_ptr(), 1 << 20, want.shape[1] - 1, 2)
    assert rc == L.ERR_INVALID_ARG


def test_resample_poly_golden(scipy_vectors):
    v = scipy_vectors
    for i, (up, down, lh, n) in enumerate(v["resample_cases"]):
        h, x = v[f"resample_{i}_h"], v[f"resample_{i}_x"]
        t = tol(h * up, x)
        for xin in (x, dev(x)):
            y = signal.resample_poly(xin, int(up), int(down), h)
            y = y.cpu().numpy() if hasattr(y, "cpu") else y
            assert y.shape == v[f"resample_{i}_y64"].shape
            assert np.abs(y - v[f"resample_{i}_y64"]).max() <= t
            assert np.abs(y - v[f"resample_{i}_y32"]).max() <= t


def test_resample_poly_reference_fixture(golden_dir):
    """sig/lib.rs:655-668: resample_poly(linspace(0,1,32,endpoint=False), 2, 3) vs resample_poly_output.npy."""
    fx = os.path.join(golden_dir, "reference_fixtures")
    x = np.load(os.path.join(fx, "sosfilt_input.npy")).astype(np.float32)
    want = np.load(os.path.join(fx, "resample_poly_output.npy"))
    y = signal.resample_poly(x, 2, 3)                      # default Kaiser design == SciPy's
    assert y.shape == want.shape == (22,)
    np.testing.assert_allclose(y, want, atol=2e-2, rtol=1e-6)          # the reference's own tolerance
    np.testing.assert_allclose(y, want, atol=1e-5)                      # and ours
    # legacy filter (2*firwin(31,1/3,hamming), sig/lib.rs:315-347) through the same kernel
    taps = np.load(os.path.join(golden_dir, "legacy_resample_taps.npy"))
    y_legacy = signal.resample_poly(x, 2, 3, (taps / 2.0).astype(np.float32))
    np.testing.assert_allclose(y_legacy, O.legacy_resample_poly_2_3(x.astype(np.float64)), atol=1e-5)
    np.testing.assert_allclose(y_legacy, want, atol=2e-2, rtol=1e-6)


def test_resample_poly_identity_and_gcd():
    x = np.random.RandomState(4).rand(2, 100).astype(np.float32)
    h = np.ones(5, np.float32)
    assert np.array_equal(signal.resample_poly(x, 4, 4, h), x)           # :3885-3886 copy
    assert np.array_equal(signal.resample_poly(dev(x), 3, 3, h).cpu().numpy(), x)
    from scipy.signal import firwin
    w = firwin(41, 0.5, window=("kaiser", 5.0)).astype(np.float32)
    a = signal.resample_poly(x, 4, 2, w)
    b = signal.resample_poly(x, 2, 1, w)
    assert np.array_equal(a, b)                                          # gcd reduction


def test_config4_slab_vs_oracle():
    """BASELINE configs[3] shape: up=3, down=2, 96-tap Kaiser, 1M-sample rows."""
    from scipy.signal import firwin
    n = 1 << 20
    h = firwin(96, 1.0 / 3.0, window=("kaiser", 5.0)).astype(np.float32)
    rng = np.random.RandomState(42)
    x = (rng.rand(3, n).astype(np.float32) * 2 - 1)
    y = signal.resample_poly(dev(x), 3, 2, h).cpu().numpy()
    assert y.shape == (3, 1572864)
    want = O.resample_poly(x, 3, 2, h)
    assert np.abs(y - want).max() <= tol(h * 3, x)


# ---- filtfilt -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,padtype,padlen", [
    ("filtfilt_odd", "odd", None), ("filtfilt_even", "even", None), ("filtfilt_const", "constant", None),
    ("filtfilt_none", None, None), ("filtfilt_odd_pad10", "odd", 10)])
def test_filtfilt_golden(scipy_vectors, name, padtype, padlen):
    v = scipy_vectors
    b, x = v["filtfilt_b"], v["filtfilt_x"]
    for xin in (x, dev(x)):
        y = signal.filtfilt(b, [1.0], xin, padtype=padtype, padlen=padlen)
        y = y.cpu().numpy() if hasattr(y, "cpu") else y
        assert y.shape == x.shape
        assert_filtfilt_close(y, v[name], b, x, padtype, padlen, what=name)


def test_filtfilt_zero_state_matches_reference_structure(scipy_vectors):
    v = scipy_vectors
    b, x = v["filtfilt_b"], v["filtfilt_x"]
    t = 2 * tol(b, x) * max(1.0, float(np.abs(b).sum()))
    for xin in (x, dev(x)):
        y = signal.filtfilt_zero_state(b, xin)
        y = y.cpu().numpy() if hasattr(y, "cpu") else y
        assert np.abs(y - v["filtfilt_refstyle"]).max() <= t
        assert np.abs(y - O.filtfilt_fir_nopad(b, x)).max() <= t


@pytest.mark.parametrize("batch,n,k,padtype", [(3, 5000, 31, "odd"), (2, 20000, 255, "odd"), (2, 7001, 64, "even"),
                                               (1, 3000, 100, "constant"), (2, 9000, 63, None), (2, 1200, 255, "odd")])
def test_filtfilt_vs_oracle_random(batch, n, k, padtype):
    from scipy.signal import firwin
    rng = np.random.RandomState(n + k)
    b = firwin(k, 0.2).astype(np.float32)
    x = (rng.rand(batch, n).astype(np.float32) * 2 - 1)
    mode = {"odd": O.PAD_ODD, "even": O.PAD_EVEN, "constant": O.PAD_CONSTANT, None: O.PAD_NONE}[padtype]
    want = O.filtfilt_fir(b, x, mode, -1)
    y = signal.filtfilt(b, [1.0], dev(x), padtype=padtype).cpu().numpy()
    assert_filtfilt_close(y, want, b, x, padtype)
    yz = signal.filtfilt_zero_state(b, dev(x)).cpu().numpy()
    assert_filtfilt_close(yz, O.filtfilt_fir_nopad(b, x), b, x, "zero_state")


@pytest.mark.parametrize("padtype", ["odd", "even", "constant"])
def test_filtfilt_single_pass_equals_two_pass(padtype):
    """Padded filtfilt runs as one zero-phase pass with b (*) flip(b) when the pad covers k-1 samples; a shorter
    pad (padlen < k-1) must fall back to two passes.  Both against the oracle, and against each other."""
    rng = np.random.RandomState(12)
    from scipy.signal import firwin
    b = firwin(63, 0.3).astype(np.float32)
    x = (rng.rand(4, 30000).astype(np.float32) * 2 - 1)
    code = {"odd": O.PAD_ODD, "even": O.PAD_EVEN, "constant": O.PAD_CONSTANT}[padtype]
    for padlen in (None, 62, 61, 10):
        one, two = gpu.Context(0), gpu.Context(0)
        two.set_option("filtfilt_fused", 0)
        y1 = signal.filtfilt(b, [1.0], dev(x), padtype=padtype, padlen=padlen, ctx=one)
        y2 = signal.filtfilt(b, [1.0], dev(x), padtype=padtype, padlen=padlen, ctx=two)
        one.sync(); two.sync()
        fused = one.get_option("filtfilt_fused_calls")
        assert fused == (1 if (padlen is None or padlen >= 62) else 0), padlen
        assert two.get_option("filtfilt_fused_calls") == 0
        want = O.filtfilt_fir(b, x, code, -1 if padlen is None else padlen)
        assert_filtfilt_close(y1.cpu().numpy(), want, b, x, padtype, padlen, fused=bool(fused), what=f"default padlen={padlen}")
        assert_filtfilt_close(y2.cpu().numpy(), want, b, x, padtype, padlen, fused=False, what=f"two-pass padlen={padlen}")


def test_in_place_filtering_is_rejected():
    """x and y may not overlap (tiles read their halo while other CTAs store): INVALID_ARG, not silent garbage."""
    lib = L.lib()
    ctx = gpu.Context(0)
    t = dev(np.ones((4, 1000), np.float32))
    taps = np.ones(5, np.float32)
    p = C.c_void_p(t.data_ptr())
    rc = lib.scir_b200_fir1d_batched_f32(ctx.handle, p, 1000, taps.ctypes.data_as(C.c_void_p), 5, L.TAPS_SCIR, p, 1000, 4, 1000)
    assert rc == L.ERR_INVALID_ARG and "in place" in L.last_error()
    p2 = C.c_void_p(t.data_ptr() + 4 * 500)                            # partial overlap
    rc = lib.scir_b200_fir1d_batched_f32(ctx.handle, p, 1000, taps.ctypes.data_as(C.c_void_p), 5, L.TAPS_SCIR, p2, 1000, 2, 1000)
    assert rc == L.ERR_INVALID_ARG
    rc = lib.scir_b200_filtfilt_fir_f32(ctx.handle, taps.ctypes.data_as(C.c_void_p), 5, L.PAD_ODD, -1, p, 1000, p, 1000, 4, 1000)
    assert rc == L.ERR_INVALID_ARG


def test_filtfilt_identity_and_errors():
    x = np.arange(12, dtype=np.float32)
    np.testing.assert_allclose(signal.filtfilt([1.0], [1.0], x), x, atol=1e-6)      # test_signaltools.py:2797-2804
    with pytest.raises(ValueError, match="padlen"):
        signal.filtfilt(np.ones(5, np.float32), [1.0], np.ones(15, np.float32))     # n <= 3*ntaps
    with pytest.raises(ValueError):
        signal.filtfilt(np.ones(5, np.float32), [1.0], np.ones(50, np.float32), padtype="bogus")


def test_config5_slab_vs_oracle():
    """BASELINE configs[4] shape: 255-tap FIR numerator, 256k-sample rows, both pad modes."""
    from scipy.signal import firwin
    n = 1 << 18
    b = firwin(255, 0.2).astype(np.float32)
    rng = np.random.RandomState(42)
    x = (rng.rand(3, n).astype(np.float32) * 2 - 1)
    y = signal.filtfilt(b, [1.0], dev(x)).cpu().numpy()
    assert_filtfilt_close(y, O.filtfilt_fir(b, x, O.PAD_ODD, -1), b, x, "odd")
    yz = signal.filtfilt_zero_state(b, dev(x)).cpu().numpy()
    assert_filtfilt_close(yz, O.filtfilt_fir_nopad(b, x), b, x, "zero_state")


# ---- DeviceArray (lib.rs:77-190) -----------------------------------------------------------------------
def test_device_array_roundtrip_and_chained_fir():
    data = [1.0, 2.0, 3.0, 4.0, 0.5, 0.0, -0.5, -1.0]
    arr = gpu.DeviceArray.from_cpu_slice([2, 4], gpu.DType.F32, data)
    assert arr.shape() == [2, 4] and arr.dtype() == gpu.DType.F32 and arr.device() == Device.Cpu
    arr.to_device(Device.Cuda)
    assert arr.device() == Device.Cuda
    assert arr.to_cpu_vec() == data
    y = arr.fir1d_batched(TAPS3)
    assert y.device() == Device.Cuda
    np.testing.assert_allclose(np.asarray(y.to_cpu_vec()).reshape(2, 4), Y24, atol=1e-7)
    arr.to_device(Device.Cpu)
    assert arr.device() == Device.Cpu and arr.to_cpu_vec() == data


# ---- multi-GPU front end (needs >= 2 devices) ------------------------------------------------------------
def test_fir_f64_device_twin():
    """fir1d_batched_f64 (lib.rs:1166-1184) on the device: the reference's own random property test
    (3 x 32, k = 5, taps 1/(i+1), atol = rtol = 1e-12: lib.rs:1263-1298), the doc-test shape (:1160-1164), then
    shapes with halos, many tiles, long filters and both tap orders against the f64 oracle."""
    rng = np.random.RandomState(1)
    x = rng.rand(3, 32) * 2 - 1
    taps = 1.0 / (np.arange(5, dtype=np.float64) + 1.0)
    y = gpu.fir1d_batched_f64_cuda(x, taps)
    np.testing.assert_allclose(y, O.fir1d_batched_f64(x, taps), atol=1e-12, rtol=1e-12)
    y = gpu.fir1d_batched_f64_cuda(np.array([[1.0, 2.0, 3.0, 4.0]]), np.array([0.25, 0.5, 0.25]))
    np.testing.assert_allclose(y, [[0.25, 1.0, 2.0, 3.0]], atol=1e-15)
    for b, n, k in ((2, 5000, 63), (3, 2048, 1), (1, 2049, 300), (4, 10000, 4097), (2, 7, 20)):
        x = rng.randn(b, n)
        taps = rng.randn(k)
        want = O.fir1d_batched_f64(x, taps)
        yd = gpu.fir1d_batched_f64_cuda(torch.from_numpy(x).cuda(), taps)
        assert yd.dtype == torch.float64
        scale = np.abs(taps).sum() * np.abs(x).max()
        assert np.abs(yd.cpu().numpy() - want).max() <= 1e-13 * scale, (b, n, k)
        yl = gpu.fir1d_batched_f64_cuda(x, taps[::-1].copy(), tap_order=L.TAPS_LFILTER)
        assert np.abs(yl - want).max() <= 1e-13 * scale
    with pytest.raises(gpu.GpuError):
        gpu.fir1d_batched_f64_cuda(torch.zeros((2, 8), device="cuda"), taps)        # f32 tensor: wrong dtype


def test_device_array_elementwise_ops_bit_exact():
    """add_scalar_auto / mul_scalar_auto / add_auto (lib.rs:268-377) on device-resident arrays: the reference's
    doc-test vectors (lib.rs:258-262, :293-298, :345-350), then random data of awkward lengths bit-exact against
    the CPU loops, chained without leaving the device."""
    a = gpu.DeviceArray.from_cpu_slice([3], gpu.DType.F32, [1.0, 2.0, 3.0])
    a.to_device(Device.Cuda)
    assert a.add_scalar_auto(1.0).to_cpu_vec() == [2.0, 3.0, 4.0]
    assert a.mul_scalar_auto(2.0).to_cpu_vec() == [2.0, 4.0, 6.0]
    b = gpu.DeviceArray.from_cpu_slice([3], gpu.DType.F32, [0.5, 1.5, 2.5])
    b.to_device(Device.Cuda)
    assert a.add_auto(b).to_cpu_vec() == [1.5, 3.5, 5.5]
    c = gpu.DeviceArray.from_cpu_slice([2], gpu.DType.F32, [1.0, 2.0])
    c.to_device(Device.Cuda)
    with pytest.raises(gpu.GpuError) as ei:
        a.add_auto(c)
    assert ei.value.kind == "ShapeMismatch"
    cpu = gpu.DeviceArray.from_cpu_slice([3], gpu.DType.F32, [1.0, 2.0, 3.0])
    # a Device.Cpu array runs the crate's own loop (lib.rs:206-221) and stays on the CPU; the array's device decides
    r = cpu.add_scalar_auto(1.0)
    assert r.device() == Device.Cpu and r.to_cpu_vec() == [2.0, 3.0, 4.0]
    with pytest.raises(gpu.GpuError):                       # operands on different devices: never silently moved
        a.add_auto(cpu)
    rng = np.random.RandomState(9)
    for n in (1, 3, 4, 1023, 4096 * 4 * 4 + 5, 3_000_001):
        x = (rng.randn(n) * 10).astype(np.float32)
        z = (rng.randn(n) * 1e-3).astype(np.float32)
        dx = gpu.DeviceArray.from_cpu_slice([n], gpu.DType.F32, x)
        dz = gpu.DeviceArray.from_cpu_slice([n], gpu.DType.F32, z)
        dx.to_device(Device.Cuda)
        dz.to_device(Device.Cuda)
        got = dx.mul_scalar_auto(1.7).add_scalar_auto(-0.3).add_auto(dz)          # stays in HBM between ops
        want = O.add_f32(O.add_scalar_f32(O.mul_scalar_f32(x, 1.7), -0.3), z)
        assert np.array_equal(np.asarray(got.to_cpu_vec(), np.float32), want), n


def test_elementwise_unaligned_views_and_aliasing():
    lib = L.lib()
    ctx = gpu.Context(0)
    rng = np.random.RandomState(10)
    x = rng.randn(100_003).astype(np.float32)
    t = dev(x)
    out = torch.empty_like(t)
    for off in (0, 1, 2, 3):                                # 4-byte aligned only: scalar path
        n = x.size - off - 5
        rc = lib.scir_b200_add_scalar_f32(ctx.handle, C.c_void_p(t.data_ptr() + 4 * off), 2.5,
                                          C.c_void_p(out.data_ptr() + 4 * off), n)
        assert rc == 0, L.last_error()
        ctx.sync()
        assert np.array_equal(out.cpu().numpy()[off:off + n], O.add_scalar_f32(x[off:off + n], 2.5))
    rc = lib.scir_b200_add_f32(ctx.handle, C.c_void_p(t.data_ptr()), C.c_void_p(t.data_ptr()), C.c_void_p(t.data_ptr()), x.size)
    assert rc == 0
    ctx.sync()
    assert np.array_equal(t.cpu().numpy(), x + x)           # in place, y aliases a and b
    assert lib.scir_b200_add_f32(ctx.handle, None, None, None, 0) == 0            # empty
    assert lib.scir_b200_mul_scalar_f32(ctx.handle, None, 1.0, None, 5) != 0      # NULL with n > 0


def test_mg_row_sharding_matches_single_device():
    """In-process multi-GPU front end (one ctx + host thread per shard, no collective).  On a one-GPU box the
    same device is listed three times: the sharding, threading and error plumbing are identical."""
    ndev = gpu.device_count()
    ids = list(range(ndev)) if ndev >= 2 else [0, 0, 0]
    lib = L.lib()
    devs = (C.c_int * len(ids))(*ids)
    mg = C.c_void_p()
    assert lib.scir_b200_mg_create(devs, len(ids), C.byref(mg)) == 0, L.last_error()
    rng = np.random.RandomState(8)
    x = (rng.rand(4 * len(ids) + 1, 30000).astype(np.float32) * 2 - 1)      # uneven shards
    taps = rng.randn(63).astype(np.float32)
    y = np.empty_like(x)
    rc = lib.scir_b200_mg_fir1d_batched_f32_host(mg, x.ctypes.data, x.shape[1], taps.ctypes.data, taps.size, 0,
                                                 y.ctypes.data, x.shape[1], x.shape[0], x.shape[1])
    assert rc == 0, L.last_error()
    assert np.array_equal(y, gpu.fir1d_batched_f32_cuda(x, taps))
    assert np.abs(y - O.fir1d_batched_f32_acc64(x, taps)).max() <= tol(taps, x)
    # resample_poly and filtfilt through the same front end
    from scipy.signal import firwin
    h = firwin(96, 1.0 / 3.0, window=("kaiser", 5.0)).astype(np.float32)
    n_out = -(-x.shape[1] * 3 // 2)
    yr = np.empty((x.shape[0], n_out), np.float32)
    rc = lib.scir_b200_mg_resample_poly_f32_host(mg, h.ctypes.data, h.size, 3, 2, x.ctypes.data, x.shape[1], x.shape[0],
                                                 x.shape[1], yr.ctypes.data, n_out)
    assert rc == 0, L.last_error()
    assert np.array_equal(yr, signal.resample_poly(x, 3, 2, h))
    b = firwin(31, 0.2).astype(np.float32)
    yf = np.empty_like(x)
    rc = lib.scir_b200_mg_filtfilt_fir_f32_host(mg, b.ctypes.data, b.size, L.PAD_ODD, -1, x.ctypes.data, x.shape[1],
                                                yf.ctypes.data, x.shape[1], x.shape[0], x.shape[1])
    assert rc == 0, L.last_error()
    assert np.array_equal(yf, signal.filtfilt(b, [1.0], x))
    # a failing shard reports which one
    rc = lib.scir_b200_mg_filtfilt_fir_f32_host(mg, b.ctypes.data, b.size, L.PAD_ODD, 10 ** 6, x.ctypes.data, x.shape[1],
                                                yf.ctypes.data, x.shape[1], x.shape[0], x.shape[1])
    assert rc == L.ERR_SHAPE and "shard" in L.last_error()
    lib.scir_b200_mg_destroy(mg)
