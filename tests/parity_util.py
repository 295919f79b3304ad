"""Shared tolerance rules of the parity tests (BASELINE.json north_star: max|err| <= 1e-5 * sum|h| * max|x|).

filtfilt: the zero-phase filter is hc = b (*) flip(b).  Kept outputs i in [k-1, n-(k-1)) see only samples of x
itself, so they hold the FLAT tolerance 1e-5 * sum|hc| * max|x|; only outputs within k-1 of either end touch the
extension, and only the odd extension (2*x[0] - x[i]) exceeds max|x| (by up to 3x).  The fused single-pass form
(scir_b200/csrc/api.cu: filtfilt_device) is ONE kernel pass and gets no factor for "two passes"; the two-pass form is
two kernels, each with its own 1e-5 budget, the first one's error filtered by the second: 2e-5 * sum|b|^2 * max|x|.
Reference tolerance convention: crates/scir-gpu/src/lib.rs:1317.
"""
import numpy as np


def tol(h, x, scale=1.0):
    return 1e-5 * float(np.abs(np.asarray(h, np.float64)).sum()) * float(np.abs(x).max() if np.size(x) else 0.0) * scale + 1e-30


def assert_filtfilt_close(y, want, b, x, padtype, padlen=None, fused=None, what=""):
    """padtype in {'odd','even','constant',None,'zero_state'}; fused: None = what the default ctx does."""
    b64 = np.asarray(b, np.float64)
    k, n = b64.size, x.shape[-1]
    edge = 0 if padtype in (None, "zero_state") else (3 * k if padlen is None else int(padlen))
    if fused is None:
        fused = padtype in ("odd", "even", "constant") and edge > 0 and edge >= k - 1
    hc = np.convolve(b64, b64[::-1])
    xmax = float(np.abs(x).max())
    flat = 1e-5 * float(np.abs(hc).sum()) * xmax if fused else 2e-5 * float(np.abs(b64).sum()) ** 2 * xmax
    flat += 1e-30
    amp = 3.0 if (padtype == "odd" and edge > 0) else 1.0
    d = np.abs(np.asarray(y, np.float64) - np.asarray(want, np.float64))
    lo, hi = min(k - 1, n), max(n - (k - 1), 0)
    if hi > lo:
        e_int = float(d[..., lo:hi].max())
        assert e_int <= flat, (what, "interior", e_int / flat, "of tolerance")
    e_edge = float(max(d[..., :lo].max() if lo > 0 else 0.0, d[..., hi:].max() if hi < n else 0.0))
    assert e_edge <= amp * flat, (what, "edges", e_edge / (amp * flat), "of tolerance")
    return float(d.max()) / flat
