"""GPU parity at BASELINE.json's FULL sizes (configs 2-5), through the default (auto-dispatched) kernels.

The oracle cannot run the whole problem in seconds, so each config is checked on (a) a strided subset of rows
(first / last / middle: every row-tile and both edges of the sharding unit) against the f64-accumulating CPU
oracle on exactly the same samples, and (b) size-independent properties over ALL rows: a checksum of
checksums against the filter's DC gain (sum of outputs = sum of inputs x sum of taps, up to edge terms that
are computed exactly), and equality of duplicated rows (row independence: no cross-row state).
Tolerance as everywhere: max|err| <= 1e-5 * sum|h| * max|x| (BASELINE.json north_star).
"""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]

torch = pytest.importorskip("torch")

from oracle import oracle as O                      # noqa: E402  (checker only)
from scir_b200 import gpu, signal                   # noqa: E402
from bench import CONFIGS, make_taps                # noqa: E402  (the bench's own workload definitions)
from parity_util import assert_filtfilt_close       # noqa: E402


def tol(h, xmax, scale=1.0):
    return 1e-5 * float(np.abs(np.asarray(h, np.float64)).sum()) * float(xmax) * scale


def synth(rows, n, seed=42):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.rand((rows, n), device="cuda", generator=g) * 2 - 1
    x[rows // 2] = x[0]                             # duplicated row: outputs must be identical bit for bit
    return x


def pick_rows(rows):
    return sorted({0, 1, rows // 2, rows // 3, rows - 2, rows - 1})


def test_config2_full_size():
    cfg = CONFIGS["c2"]
    rows, n = cfg["rows"], cfg["n"]
    b = make_taps(cfg)
    x = synth(rows, n)
    ctx = gpu.torch_context(x)
    t0 = ctx.get_option("toeplitz_launches")
    y = signal.lfilter(b, [1.0], x)
    torch.cuda.synchronize()
    assert ctx.get_option("toeplitz_launches") == t0 + 1           # the tensor-core kernel is the default here
    sel = pick_rows(rows)
    xs, ys = x[sel].cpu().numpy(), y[sel].cpu().numpy()
    assert np.abs(ys - O.lfilter_fir(b, xs)).max() <= tol(b, 1.0)
    assert torch.equal(y[rows // 2], y[0])
    # checksum of checksums: sum_i y[i] = sum_d b[d] * (sum_i x[i] - tail_d), tail_d = last d samples
    sx = x.double().sum(dim=1)
    sy = y.double().sum(dim=1)
    k = b.size
    tails = torch.cumsum(torch.flip(x[:, n - k:].double(), dims=[1]), dim=1)   # tails[:, d-1] = sum of the last d samples
    bd = torch.from_numpy(b.astype(np.float64)).cuda()
    want = bd.sum() * sx - (tails[:, : k - 1] * bd[1:]).sum(dim=1)
    assert float((sy - want).abs().max()) <= 1e-5 * float(np.abs(b).sum()) * np.sqrt(n) * 4 + n * 2e-8
    # the FP32 direct family on the same data (A/B arm): both inside the tolerance of each other
    direct = gpu.Context(0)
    direct.set_option("long_tap_path", 1)
    yd = signal.lfilter(b, [1.0], x[: 64], ctx=direct)
    direct.sync()
    assert float((yd - y[:64]).abs().max()) <= 2 * tol(b, 1.0)


def test_config3_full_size_segments():
    cfg = CONFIGS["c3"]
    rows, n = cfg["rows"], cfg["n"]
    b = make_taps(cfg)
    k = b.size
    x = synth(rows, n)
    y = signal.lfilter(b, [1.0], x)
    torch.cuda.synchronize()
    assert torch.equal(y[rows // 2], y[0])
    seg = 20000
    for r in pick_rows(rows):
        for start in (0, 16384 - 100, n // 2 + 777, n - seg):     # row start, a tile boundary, the middle, the row end
            lo = max(0, start - (k - 1))
            xs = x[r, lo:start + seg].cpu().numpy()[None, :]
            want = O.lfilter_fir(b, xs)[0, start - lo:]
            got = y[r, start:start + seg].cpu().numpy()
            if lo > 0:
                # the oracle started from zero state at `lo`; outputs from `start` on have a full window
                assert start - lo == k - 1
            assert np.abs(got - want).max() <= tol(b, 1.0), (r, start)
    # DC property on all rows (sum of taps = 1 for firwin): mean of a long row is preserved
    assert float((y.double().mean(dim=1) - x.double().mean(dim=1)).abs().max()) <= 1e-4


def test_config4_full_size():
    cfg = CONFIGS["c4"]
    rows, n = cfg["rows"], cfg["n"]
    h = make_taps(cfg)
    x = synth(rows, n)
    y = signal.resample_poly(x, cfg["up"], cfg["down"], h)
    torch.cuda.synchronize()
    n_out = -(-n * cfg["up"] // cfg["down"])
    assert tuple(y.shape) == (rows, n_out)                          # integer plan, bit-exact
    assert torch.equal(y[rows // 2], y[0])
    sel = pick_rows(rows)
    xs, ys = x[sel].cpu().numpy(), y[sel].cpu().numpy()
    want = O.resample_poly(xs, cfg["up"], cfg["down"], h)
    assert want.shape == ys.shape
    assert np.abs(ys - want).max() <= tol(h * cfg["up"], 1.0)
    # linearity over ALL rows at full size: resample(2a - 3c) == 2 resample(a) - 3 resample(c)
    c = synth(rows, n, seed=7)
    yc = signal.resample_poly(c, cfg["up"], cfg["down"], h)
    ys2 = signal.resample_poly(2.0 * x - 3.0 * c, cfg["up"], cfg["down"], h)
    ys2 -= 2.0 * y
    ys2 += 3.0 * yc
    assert float(ys2.abs().max()) <= 2 * tol(h * cfg["up"], 5.0)


@pytest.mark.parametrize("padtype", ["odd", None])
def test_config5_full_size(padtype):
    cfg = CONFIGS["c5"]
    rows, n = cfg["rows"], cfg["n"]
    b = make_taps(cfg)
    x = synth(rows, n)
    ctx = gpu.torch_context(x)
    t0 = ctx.get_option("toeplitz_launches") + ctx.get_option("os_launches")
    y = signal.filtfilt(b, [1.0], x, padtype=padtype)
    torch.cuda.synchronize()
    # padded: ONE zero-phase pass with b (*) flip(b) (509 taps); unpadded: forward + anticausal pass -- on the
    # overlap-save FFT kernel or the tensor kernel, whichever the dispatch (api.cu: launch_fir) costs cheaper
    assert ctx.get_option("toeplitz_launches") + ctx.get_option("os_launches") == t0 + (1 if padtype == "odd" else 2)
    assert torch.equal(y[rows // 2], y[0])
    sel = pick_rows(rows)
    xs, ys = x[sel].cpu().numpy(), y[sel].cpu().numpy()
    want = O.filtfilt_fir(b, xs, O.PAD_ODD if padtype == "odd" else O.PAD_NONE, -1)
    hc = np.convolve(b.astype(np.float64), b[::-1].astype(np.float64))
    # interior: flat 1e-5 * sum|hc| * max|x| (one fused pass) / 2e-5 * sum|b|^2 * max|x| (two passes); x3 only where
    # the odd extension is touched (tests/parity_util.py)
    assert_filtfilt_close(ys, want, b, xs, padtype)
    # zero phase: a symmetric-filtered signal reversed equals the filtered reversed signal (interior)
    yr = signal.filtfilt(b, [1.0], torch.flip(x[:8], dims=[1]).contiguous(), padtype=padtype)
    k = b.size                                                       # FIR: edge rules reach 2(k-1) samples inwards at most
    d = (torch.flip(yr, dims=[1]) - y[:8])[:, 2 * k:n - 2 * k]
    assert float(d.abs().max()) <= 2 * tol(hc, 1.0, 2.0)            # two results, each within its own (interior) tolerance
