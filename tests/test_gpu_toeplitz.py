"""GPU parity for the tcgen05 block-Toeplitz FIR path (scir_b200/csrc/fir_toeplitz.cu), forced on
through the ctx option `long_tap_path=2` so that short filters exercise it too.  Same tolerance as
every other kernel of the path: max|err| <= 1e-5 * sum|h| * max|x| against the f64-accumulating
oracle (BASELINE.json north_star).  A hung tensor-core pipeline must not hang the suite: every test
carries a hard timeout (thread method: the process is killed)."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]

torch = pytest.importorskip("torch")

from oracle import oracle as O                      # noqa: E402  (checker only)
from scir_b200 import gpu, signal                   # noqa: E402


def tol(h, x):
    return 1e-5 * float(np.abs(np.asarray(h, np.float64)).sum()) * float(np.abs(x).max()) + 1e-30


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def toep_ctx(terms=4, loader=0):
    ctx = gpu.Context(0)
    ctx.set_option("long_tap_path", 2)
    ctx.set_option("toeplitz_terms", terms)
    ctx.set_option("toeplitz_loader", loader)      # 0: TMA raw ring when it fits, 1: register-prefetch loader
    return ctx


def run(ctx, fn):
    torch.cuda.synchronize()
    y = fn()
    ctx.sync()
    return y


@pytest.mark.parametrize("batch,n,k", [(1, 16384, 1), (2, 16384, 63), (2, 20000, 63), (3, 40000, 255), (2, 33000, 129),
                                       (1, 100000, 1500), (2, 70000, 4097), (1, 5, 3), (2, 127, 200), (5, 16385, 64)])
@pytest.mark.parametrize("terms,loader", [(3, 0), (4, 0), (6, 0), (4, 1)])
def test_toeplitz_fir_vs_oracle(batch, n, k, terms, loader):
    if terms == 6 and k > 1500:
        pytest.skip("three split terms per operand do not fit shared memory for very long filters")
    rng = np.random.RandomState(batch * 131 + n + k)
    x = (rng.rand(batch, n).astype(np.float32) * 2 - 1)
    taps = rng.randn(k).astype(np.float32)
    want = O.fir1d_batched_f32_acc64(x, taps)
    ctx = toep_ctx(terms, loader)
    t0 = ctx.get_option("toeplitz_launches")
    y = run(ctx, lambda: gpu.fir1d_batched_f32_cuda(dev(x), taps, ctx=ctx)).cpu().numpy()
    assert ctx.get_option("toeplitz_launches") == t0 + 1          # the tensor-core kernel is the one that ran
    err = np.abs(y - want).max()
    assert err <= tol(taps, x), (err / tol(taps, x), "of tolerance")


def test_toeplitz_error_budget_reported():
    """How much of the 1e-5 tolerance each split uses on BASELINE-shaped data (firwin taps, U[-1,1))
    and on a coherent worst-ish case (constant input, all-positive taps)."""
    from scipy.signal import firwin
    rng = np.random.RandomState(7)
    rows = []
    for k, cutoff in ((63, 0.25), (255, 0.2), (4097, 0.01)):
        b = firwin(k, cutoff).astype(np.float32)
        for name, x in (("uniform", (rng.rand(2, 50000).astype(np.float32) * 2 - 1)),
                        ("const0.7", np.full((1, 50000), 0.7, np.float32))):
            want = O.lfilter_fir(b, x)
            for terms in (3, 4, 6):
                if terms == 6 and k > 1500:
                    continue
                ctx = toep_ctx(terms)
                y = run(ctx, lambda: signal.lfilter(b, np.ones(1, np.float32), dev(x), ctx=ctx)).cpu().numpy()
                frac = np.abs(y - want).max() / tol(b, x)
                rows.append((k, name, terms, frac))
                assert frac <= 1.0, (k, name, terms, frac)
    print("\n".join(f"k={k} {name} terms={t}: {f:.3f} of tolerance" for k, name, t, f in rows))


@pytest.mark.parametrize("loader", [0, 1])
@pytest.mark.parametrize("padtype", ["odd", "even", "constant", None])
def test_toeplitz_filtfilt_matches_direct_and_oracle(padtype, loader):
    """Anticausal pass, held boundary and signal extension go through the Toeplitz loader too."""
    from scipy.signal import firwin
    rng = np.random.RandomState(11)
    b = firwin(255, 0.2).astype(np.float32)
    x = (rng.rand(3, 40001).astype(np.float32) * 2 - 1)
    want = O.filtfilt_fir(b, x, padtype={"odd": O.PAD_ODD, "even": O.PAD_EVEN, "constant": O.PAD_CONSTANT,
                                         None: O.PAD_NONE}[padtype])
    ctx = toep_ctx(4, loader)
    y = run(ctx, lambda: signal.filtfilt(b, [1.0], dev(x), padtype=padtype, ctx=ctx)).cpu().numpy()
    assert ctx.get_option("toeplitz_launches") == 2
    hc = np.convolve(b.astype(np.float64), b[::-1].astype(np.float64))
    assert np.abs(y - want).max() <= tol(hc, x) * 2
    yz = run(ctx, lambda: signal.filtfilt_zero_state(b, dev(x), ctx=ctx)).cpu().numpy()
    assert np.abs(yz - O.filtfilt_fir_nopad(b, x)).max() <= tol(hc, x) * 2


def test_toeplitz_many_tiles_and_views():
    """More tiles than SMs (persistent CTAs wrap both pipelines), unaligned rows (scalar loader path)."""
    rng = np.random.RandomState(3)
    taps = rng.randn(300).astype(np.float32)
    big = (rng.rand(40, 100003).astype(np.float32) * 2 - 1)
    xb = dev(big)
    _many_tiles(taps, big, xb, toep_ctx(4, 0))
    _many_tiles(taps, big, xb, toep_ctx(4, 1))


def _many_tiles(taps, big, xb, ctx):
    naive = gpu.Context(0)
    naive.set_option("variant", 2)
    for view in (xb[:, :98304], xb[:, 1:], xb[::3, 3:90001]):
        y = run(ctx, lambda: gpu.fir1d_batched_f32_cuda(view, taps, ctx=ctx))
        y2 = run(naive, lambda: gpu.fir1d_batched_f32_cuda(view, taps, ctx=naive))
        assert float((y - y2).abs().max()) <= 2 * tol(taps, big)
    y = run(ctx, lambda: gpu.fir1d_batched_f32_cuda(xb[:2, :98304], taps, ctx=ctx)).cpu().numpy()
    assert np.abs(y - O.fir1d_batched_f32_acc64(big[:2, :98304], taps)).max() <= tol(taps, big)


def test_auto_dispatch_uses_tensor_path_for_long_taps():
    rng = np.random.RandomState(5)
    x = dev(rng.rand(2, 50000).astype(np.float32))
    ctx = gpu.Context(0)
    run(ctx, lambda: gpu.fir1d_batched_f32_cuda(x, rng.randn(63).astype(np.float32), ctx=ctx))
    assert ctx.get_option("toeplitz_launches") == 0
    run(ctx, lambda: gpu.fir1d_batched_f32_cuda(x, rng.randn(2000).astype(np.float32), ctx=ctx))
    assert ctx.get_option("toeplitz_launches") == 1
    ctx.set_option("long_tap_path", 1)
    run(ctx, lambda: gpu.fir1d_batched_f32_cuda(x, rng.randn(2000).astype(np.float32), ctx=ctx))
    assert ctx.get_option("toeplitz_launches") == 1
