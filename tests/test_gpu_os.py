"""GPU parity of the overlap-save FFT FIR path (scir_b200/csrc/fir_os.cu), forced on through the ctx option
`long_tap_path=3` so that short filters and small launches exercise it too.  Same tolerance as every other kernel of the
path: max|err| <= 1e-5 * sum|h| * max|x| against the f64-accumulating oracle (BASELINE.json north_star), with an asserted
margin on BASELINE-shaped data (the f32 FFT's error is ~log2(N) * 2^-24 of the block RMS)."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]

torch = pytest.importorskip("torch")

from oracle import oracle as O                      # noqa: E402  (checker only)
from scir_b200 import gpu, signal                   # noqa: E402
from parity_util import assert_filtfilt_close, tol  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def os_ctx(packed=0):
    ctx = gpu.Context(0)
    ctx.set_option("long_tap_path", 3)
    ctx.set_option("os_packed", packed)        # N = 16384 blocks: 0 scalar kernel (default), 1 packed-lane kernel (A/B arm)
    return ctx


def run(ctx, fn):
    torch.cuda.synchronize()
    y = fn()
    ctx.sync()
    return y


@pytest.mark.parametrize("batch,n,k", [(1, 16384, 2), (2, 20000, 63), (3, 40000, 255), (2, 33000, 509), (2, 9000, 1025),
                                       (1, 100000, 640), (1, 100000, 641), (1, 100000, 1537), (2, 70000, 4097), (1, 40000, 7936), (1, 5, 3),
                                       (2, 127, 200), (5, 16385, 64), (1, 3588, 509), (1, 3589, 509), (1, 7176, 509), (1, 7177, 509)])
def test_os_fir_vs_oracle(batch, n, k):
    rng = np.random.RandomState(batch * 131 + n + k)
    x = (rng.rand(batch, n).astype(np.float32) * 2 - 1)
    taps = rng.randn(k).astype(np.float32)
    want = O.fir1d_batched_f32_acc64(x, taps)
    for packed in ((0, 1) if k > 640 else (0,)):
        ctx = os_ctx(packed)
        y = run(ctx, lambda: gpu.fir1d_batched_f32_cuda(dev(x), taps, ctx=ctx)).cpu().numpy()
        assert ctx.get_option("os_launches") == 1                # the FFT kernel is the one that ran
        err = np.abs(y - want).max()
        assert err <= tol(taps, x), (packed, err / tol(taps, x), "of tolerance")


@pytest.mark.parametrize("k,cutoff", [(255, 0.2), (509, 0.2), (1025, 0.05), (4097, 0.01)])
def test_os_firwin_margin(k, cutoff):
    """BASELINE-shaped data (firwin taps; uniform noise, a constant, a full-scale tone): <= 0.25 of the tolerance."""
    from scipy.signal import firwin
    rng = np.random.RandomState(k)
    b = firwin(k, cutoff).astype(np.float32)
    n = 1 << 17
    t = np.arange(n)
    for name, x in (("uniform", (rng.rand(3, n).astype(np.float32) * 2 - 1)),
                    ("const0.7", np.full((1, n), 0.7, np.float32)),
                    ("tone", (0.9 * np.sin(2 * np.pi * 0.003 * t)).astype(np.float32)[None, :]),
                    ("tone+noise", (0.9 * np.sin(2 * np.pi * 0.003 * t) + 1e-3 * rng.randn(n)).astype(np.float32)[None, :])):
        want = O.lfilter_fir(b, x)
        for packed in ((0, 1) if k > 640 else (0,)):
            ctx = os_ctx(packed)
            y = run(ctx, lambda: signal.lfilter(b, [1.0], dev(x), ctx=ctx)).cpu().numpy()
            assert ctx.get_option("os_launches") == 1
            frac = float(np.abs(y - want).max()) / tol(b, x)
            assert frac <= 0.25, (k, name, packed, frac, "of tolerance")


def test_os_views_offsets_and_many_blocks():
    rng = np.random.RandomState(3)
    taps = rng.randn(300).astype(np.float32)
    big = (rng.rand(40, 100003).astype(np.float32) * 2 - 1)
    xb = dev(big)
    ctx = os_ctx()
    naive = gpu.Context(0)
    naive.set_option("variant", 2)
    for view in (xb[:, :98304], xb[:, 1:], xb[::3, 3:90001]):
        y = run(ctx, lambda: gpu.fir1d_batched_f32_cuda(view, taps, ctx=ctx))
        y2 = run(naive, lambda: gpu.fir1d_batched_f32_cuda(view, taps, ctx=naive))
        assert float((y - y2).abs().max()) <= 2 * tol(taps, big)


@pytest.mark.parametrize("padtype", ["odd", "even", "constant", None])
def test_os_filtfilt_both_directions(padtype):
    """The fused single pass (extension in the loader) and the two-pass form (causal + ANTICAUSAL block direction, held
    boundary) both run on the FFT kernel."""
    from scipy.signal import firwin
    rng = np.random.RandomState(11)
    b = firwin(255, 0.2).astype(np.float32)
    x = (rng.rand(3, 40001).astype(np.float32) * 2 - 1)
    want = O.filtfilt_fir(b, x, padtype={"odd": O.PAD_ODD, "even": O.PAD_EVEN, "constant": O.PAD_CONSTANT, None: O.PAD_NONE}[padtype])
    ctx = os_ctx()
    y = run(ctx, lambda: signal.filtfilt(b, [1.0], dev(x), padtype=padtype, ctx=ctx)).cpu().numpy()
    assert ctx.get_option("os_launches") == (2 if padtype is None else 1)
    assert_filtfilt_close(y, want, b, x, padtype, what="default")
    two = os_ctx()
    two.set_option("filtfilt_fused", 0)
    y2 = run(two, lambda: signal.filtfilt(b, [1.0], dev(x), padtype=padtype, ctx=two)).cpu().numpy()
    assert two.get_option("os_launches") == 2
    assert_filtfilt_close(y2, want, b, x, padtype, fused=False, what="two-pass")
    yz = run(ctx, lambda: signal.filtfilt_zero_state(b, dev(x), ctx=ctx)).cpu().numpy()
    assert_filtfilt_close(yz, O.filtfilt_fir_nopad(b, x), b, x, "zero_state")


@pytest.mark.parametrize("k,packed", [(63, 0), (300, 0), (2000, 0), (2000, 1)])
def test_os_non_finite_samples_stay_local(k, packed):
    """A NaN / Inf sample must reach exactly the outputs whose window contains it, not its whole FFT block:
    flagged block pairs are redone by os_fixup_kernel the reference's way."""
    rng = np.random.RandomState(k)
    x = (rng.rand(3, 60000).astype(np.float32) * 2 - 1)
    taps = np.abs(rng.randn(k)).astype(np.float32) + 0.1
    bad = {0: [(100, np.nan)], 1: [(16384 - 5, np.inf), (40000, -np.inf)], 2: []}
    for r, lst in bad.items():
        for pos, v in lst:
            x[r, pos] = v
    ctx = os_ctx(packed)
    y = run(ctx, lambda: gpu.fir1d_batched_f32_cuda(dev(x), taps, ctx=ctx)).cpu().numpy()
    with np.errstate(invalid="ignore"):
        ref = O.fir1d_batched_f32(x, taps)
    assert np.array_equal(np.isnan(y), np.isnan(ref))
    assert np.array_equal(np.isposinf(y), np.isposinf(ref)) and np.array_equal(np.isneginf(y), np.isneginf(ref))
    for r, lst in bad.items():
        mask = np.ones(x.shape[1], bool)
        for pos, _ in lst:
            mask[pos:pos + k] = False
        xs = np.where(np.isfinite(x[r]), x[r], 0).astype(np.float32)[None, :]
        want = O.fir1d_batched_f32_acc64(xs, taps)[0]
        assert np.abs(y[r][mask] - want[mask]).max() <= tol(taps, xs)


def test_os_dynamic_range_error_model():
    """Like the tensor path, the FFT path's error is absolute per block (relative to the block's energy): a quiet stretch
    sharing an FFT block with a loud burst is accurate to the burst's scale -- inside the path's tolerance, which is stated
    relative to max|x| -- and a quiet block on its own is accurate to its own scale."""
    from scipy.signal import firwin
    rng = np.random.RandomState(99)
    b = firwin(509, 0.2).astype(np.float32)
    n = 1 << 16
    x = (rng.rand(2, n).astype(np.float32) * 2 - 1) * np.float32(1e-6)
    x[0, 20000:20064] = 1.0
    want = O.lfilter_fir(b, x)
    ctx = os_ctx()
    y = run(ctx, lambda: signal.lfilter(b, [1.0], dev(x), ctx=ctx)).cpu().numpy()
    err = np.abs(y - want)
    assert err.max() <= tol(b, x)
    assert err[1].max() <= tol(b, x[1])                       # a uniformly quiet row: relative to ITS scale
    assert err[0, 40000:].max() <= tol(b, x[1])               # quiet blocks of the loud row, away from the burst


def test_auto_dispatch_costs_the_fft_path_against_the_tensor_kernel():
    """api.cu: launch_fir takes the FFT path when its cost model beats the tensor kernel's: long filters on rows that fill
    block pairs (config 3's shape) -- not short filters, not the 509 fused taps of config 5, not rows much shorter than a
    block pair (which would pay for whole 16384-point transforms)."""
    rng = np.random.RandomState(5)
    ctx = gpu.Context(0)
    big = dev(rng.rand(64, 1 << 17).astype(np.float32))
    run(ctx, lambda: gpu.fir1d_batched_f32_cuda(big, rng.randn(4097).astype(np.float32), ctx=ctx))
    assert ctx.get_option("os_launches") == 1                 # long taps, long rows: FFT
    run(ctx, lambda: gpu.fir1d_batched_f32_cuda(big, rng.randn(63).astype(np.float32), ctx=ctx))
    run(ctx, lambda: gpu.fir1d_batched_f32_cuda(big, rng.randn(509).astype(np.float32), ctx=ctx))
    assert ctx.get_option("os_launches") == 1                 # short and mid-length filters stay on the tensor / direct kernels
    short_rows = dev(rng.rand(4096, 5000).astype(np.float32))
    run(ctx, lambda: gpu.fir1d_batched_f32_cuda(short_rows, rng.randn(1025).astype(np.float32), ctx=ctx))
    assert ctx.get_option("os_launches") == 1                 # a 5000-sample row would pay for a whole 16384-point pair
    ctx.set_option("long_tap_path", 2)
    run(ctx, lambda: gpu.fir1d_batched_f32_cuda(big, rng.randn(4097).astype(np.float32), ctx=ctx))
    assert ctx.get_option("os_launches") == 1 and ctx.get_option("toeplitz_launches") >= 1
    ctx.set_option("long_tap_path", 3)
    run(ctx, lambda: gpu.fir1d_batched_f32_cuda(short_rows, rng.randn(1025).astype(np.float32), ctx=ctx))
    assert ctx.get_option("os_launches") == 2                 # forced
