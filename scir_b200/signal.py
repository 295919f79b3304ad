"""Host-side mirror of crate `scir-signal`'s FIR routes onto the GPU path.

    gpu.fir1d_batched_f32(x, taps, device)   crates/scir-signal/src/lib.rs:365-375 (forwarder)
    lfilter(b, a, x)           a = [a0]      SciPy spec: scipy/signal/_signaltools.py:2181-2242
    upfirdn(h, x, up, down)                  scipy/signal/_upfirdn.py:107-216, _upfirdn_apply.pyx:421-481
    resample_poly(x, up, down, window=h)     scipy/signal/_signaltools.py:3865-3957
    filtfilt(b, a, x, padtype, padlen)       scipy/signal/_signaltools.py:4745-4826 (FIR numerator)
    filtfilt_zero_state(b, x)                the reference's own structure, sig/lib.rs:278-291

The reference has only the forwarder; `resample_poly` there is f64, 2/3-only (sig/lib.rs:313-362)
and `filtfilt` is SOS-only, so SciPy -- which the reference's fixtures are generated from
(scripts/gen_signal_fixtures.py) -- is the behavioural spec.  All dispatch logic
(tap reversal, gcd reduction, pad plan, odd extension, zi) lives in C++ behind the C ABI; this file
only marshals arrays.  Inputs are (batch, n) or (n,) float32; numpy => *_host ABI calls, torch CUDA
tensors => device-pointer ABI calls on torch's current stream.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib as L
from . import gpu as G
from .gpu import Device, GpuError, _check, _is_torch, _load, _ptr, _taps_f32


class gpu:  # noqa: N801  (mirrors `pub mod gpu` in scir-signal)
    @staticmethod
    def fir1d_batched_f32(x, taps, device: Device, **kw):
        """sig/lib.rs:372-374"""
        return G.fir1d_batched_f32_auto(x, taps, device, **kw)


def _as_2d(x):
    """Returns (array2d, was_1d)."""
    if _is_torch(x):
        return (x[None, :], True) if x.dim() == 1 else (x, False)
    a = np.asarray(x, dtype=np.float32)
    return (a[None, :], True) if a.ndim == 1 else (a, False)


def _prep(x2):
    """(device?, obj, ptr, ld, batch, n, ctx)"""
    if _is_torch(x2):
        xt, ld = G._torch_matrix(x2)
        return True, xt, xt.data_ptr(), ld, xt.shape[0], xt.shape[1], G.torch_context(xt)
    a = G._host_matrix(x2)
    return False, a, _ptr(a), max(a.shape[1], 1), a.shape[0], a.shape[1], G.default_context()


def _alloc_like(dev, ref, batch, n_out):
    if dev:
        import torch
        y = torch.empty((batch, n_out), dtype=torch.float32, device=ref.device)
        return y, y.data_ptr(), max(n_out, 1)
    y = np.empty((batch, n_out), dtype=np.float32)
    return y, _ptr(y), max(n_out, 1)


class _DevBuf:
    """Library-owned device buffer for the host-array variants that have no *_host ABI twin."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.p = ctx, C.c_void_p()
        _check(_load().scir_b200_malloc(ctx.handle, max(int(nbytes), 4), C.byref(self.p)))

    def upload(self, a: np.ndarray):
        _check(_load().scir_b200_memcpy_h2d(self.ctx.handle, self.p, _ptr(a), a.nbytes))
        return self

    def download(self, a: np.ndarray):
        _check(_load().scir_b200_memcpy_d2h(self.ctx.handle, _ptr(a), self.p, a.nbytes))
        return a

    def free(self):
        _check(_load().scir_b200_free(self.ctx.handle, self.p))


def lfilter(b, a, x, zi=None, *, ctx=None):
    """lfilter(b, [a0], x[, zi]) along the last axis.  Returns y, or (y, zf) when zi is given."""
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    if a.size != 1:
        raise GpuError.backend_unavailable("only FIR (len(a) == 1) filters run on this path; IIR stays on the CPU")
    bt = _taps_f32(b)
    x2, was1d = _as_2d(x)
    dev, xo, xp, ld, batch, n, c = _prep(x2)
    c = ctx or c
    y, yp, ldy = _alloc_like(dev, xo, batch, n)
    lib = _load()
    if zi is None:
        if dev:
            _check(lib.scir_b200_lfilter_fir_f32(c.handle, _ptr(bt), bt.size, float(a[0]), xp, ld, None, None,
                                                 yp, ldy, batch, n))
        else:
            # host arrays: b / a[0] in f32 (_signaltools.py:2223), then the streaming *_host hot path
            if float(a[0]) == 0.0:
                raise ValueError("a[0] must be nonzero")
            bs = (bt / np.float32(a[0])).astype(np.float32) if float(a[0]) != 1.0 else bt
            _check(lib.scir_b200_fir1d_batched_f32_host(c.handle, xp, ld, _ptr(bs), bs.size, L.TAPS_LFILTER,
                                                        yp, ldy, batch, n))
        return y[0] if was1d else y
    if not dev:
        km1 = bt.size - 1
        zih = np.ascontiguousarray(zi, dtype=np.float32).reshape(batch, km1)
        zfh = np.empty_like(zih)
        bufs = [_DevBuf(c, xo.nbytes).upload(xo), _DevBuf(c, zih.nbytes).upload(zih),
                _DevBuf(c, zih.nbytes), _DevBuf(c, xo.nbytes)]
        try:
            _check(lib.scir_b200_lfilter_fir_f32(c.handle, _ptr(bt), bt.size, float(a[0]), bufs[0].p, ld,
                                                 bufs[1].p if km1 else None, bufs[2].p if km1 else None,
                                                 bufs[3].p, ldy, batch, n))
            c.sync()
            bufs[3].download(y)
            if km1:
                bufs[2].download(zfh)
        finally:
            for bf in bufs:
                bf.free()
        return (y[0], zfh[0]) if was1d else (y, zfh)
    import torch
    zit = zi.reshape(batch, bt.size - 1).contiguous().to(torch.float32)
    zf = torch.empty_like(zit)
    _check(lib.scir_b200_lfilter_fir_f32(c.handle, _ptr(bt), bt.size, float(a[0]), xp, ld,
                                         zit.data_ptr() if zit.numel() else None,
                                         zf.data_ptr() if zf.numel() else None, yp, ldy, batch, n))
    return (y[0], zf[0]) if was1d else (y, zf)


def upfirdn_output_len(len_h, in_len, up, down) -> int:
    """_output_len, _upfirdn_apply.pyx:59-67"""
    return int(_load().scir_b200_upfirdn_out_len(int(len_h), int(in_len), int(up), int(down)))


UPFIRDN_MODES = {"constant": L.EXT_CONSTANT, "symmetric": L.EXT_SYMMETRIC, "edge": L.EXT_EDGE, "smooth": L.EXT_SMOOTH,
                 "wrap": L.EXT_PERIODIC, "reflect": L.EXT_REFLECT, "antisymmetric": L.EXT_ANTISYMMETRIC,
                 "antireflect": L.EXT_ANTIREFLECT, "line": L.EXT_LINE}       # mode_enum, _upfirdn_apply.pyx:89-105


def _mode_code(mode):
    try:
        return UPFIRDN_MODES[mode]
    except (KeyError, TypeError):
        raise ValueError(f"Unknown mode: {mode}") from None                  # pyx:104


def upfirdn(h, x, up=1, down=1, mode="constant", cval=0.0, *, ctx=None):
    """upfirdn(h, x, up, down, mode, cval) along the last axis (scipy/signal/_upfirdn.py:107-216); f32 in, f32 out."""
    if int(up) != up or int(down) != down:
        raise ValueError("up and down must be integers")                     # _upfirdn.py:95-97
    up, down = int(up), int(down)
    if up < 1 or down < 1:
        raise ValueError("Both up and down must be >= 1")                    # _upfirdn.py:98
    m = _mode_code(mode)
    ht = _taps_f32(h)
    x2, was1d = _as_2d(x)
    dev, xo, xp, ld, batch, n, c = _prep(x2)
    c = ctx or c
    if n < 1:
        raise ValueError("x must have at least one sample")
    lo = upfirdn_output_len(ht.size, n, up, down)
    if dev:
        y, yp, ldy = _alloc_like(dev, xo, batch, lo)
        _check(_load().scir_b200_upfirdn_mode_f32(c.handle, _ptr(ht), ht.size, up, down, m, float(cval), xp, ld, batch, n,
                                                  yp, ldy, 0, lo))
        return y[0] if was1d else y
    y = np.empty((batch, lo), dtype=np.float32)
    bx, by = _DevBuf(c, xo.nbytes).upload(xo), _DevBuf(c, y.nbytes)
    try:
        _check(_load().scir_b200_upfirdn_mode_f32(c.handle, _ptr(ht), ht.size, up, down, m, float(cval), bx.p, ld, batch, n,
                                                  by.p, max(lo, 1), 0, lo))
        c.sync()
        by.download(y)
    finally:
        bx.free()
        by.free()
    return y[0] if was1d else y


def resample_poly_plan(n_in, len_h, up, down) -> dict:
    """The int64 plan of resample_poly (_signaltools.py:3882-3918); bit-exact with SciPy."""
    p = L.ResamplePlan()
    _check(_load().scir_b200_resample_poly_plan(int(n_in), int(len_h), int(up), int(down), C.byref(p)))
    return p.as_dict()


def kaiser_lowpass(up: int, down: int) -> np.ndarray:
    """SciPy's default resample_poly design (_signaltools.py:3898-3904):
    firwin(2*10*max(up,down)+1, 1/max(up,down), window=('kaiser', 5.0)), restated with numpy so the
    default works without SciPy.  Host-side, a few hundred taps."""
    g = math.gcd(int(up), int(down))
    max_rate = max(int(up) // g, int(down) // g)
    ntaps = 2 * 10 * max_rate + 1
    cutoff = 1.0 / max_rate
    m = np.arange(ntaps) - (ntaps - 1) / 2.0
    h = cutoff * np.sinc(cutoff * m) * np.kaiser(ntaps, 5.0)
    return (h / h.sum()).astype(np.float64)


_PAD_STATS = {"mean": L.PAD_STAT_MEAN, "median": L.PAD_STAT_MEDIAN, "minimum": L.PAD_STAT_MINIMUM,
              "maximum": L.PAD_STAT_MAXIMUM}


def resample_poly(x, up, down, window=("kaiser", 5.0), padtype="constant", cval=None, *, ctx=None):
    """resample_poly(x, up, down, window, padtype, cval) along the last axis (scipy/signal/_signaltools.py:3865-3957).
    padtype: an upfirdn extension mode, or 'mean' / 'median' / 'minimum' / 'maximum'."""
    if int(up) != up or int(down) != down:
        raise ValueError("up and down must be integers")
    up, down = int(up), int(down)
    if up < 1 or down < 1:
        raise ValueError("up and down must be >= 1")                         # :3876-3877
    if cval is not None and padtype != "constant":
        raise ValueError("cval has no effect when padtype is " + str(padtype))   # :3878-3879
    if padtype in _PAD_STATS:
        code = _PAD_STATS[padtype]
    elif padtype in UPFIRDN_MODES:
        code = UPFIRDN_MODES[padtype]
    else:
        raise ValueError("padtype must be one of: " + ", ".join(list(UPFIRDN_MODES) + list(_PAD_STATS)))   # :3932-3934
    if isinstance(window, (tuple, str)):
        if window != ("kaiser", 5.0):
            raise GpuError.backend_unavailable("pass the filter as an array; only the default "
                                               "('kaiser', 5.0) design is built in")
        window = kaiser_lowpass(up, down)
    w = _taps_f32(window)
    x2, was1d = _as_2d(x)
    dev, xo, xp, ld, batch, n, c = _prep(x2)
    c = ctx or c
    plan = resample_poly_plan(n, w.size, up, down)
    n_out = n if (plan["up"] == 1 and plan["down"] == 1) else plan["n_out"]
    plain = (code == L.EXT_CONSTANT and not cval)
    if dev or plain:
        y, yp, ldy = _alloc_like(dev, xo, batch, n_out)
        if plain:
            fn = _load().scir_b200_resample_poly_f32 if dev else _load().scir_b200_resample_poly_f32_host
            _check(fn(c.handle, _ptr(w), w.size, up, down, xp, ld, batch, n, yp, ldy))
        else:
            _check(_load().scir_b200_resample_poly_pad_f32(c.handle, _ptr(w), w.size, up, down, code, float(cval or 0.0),
                                                           xp, ld, batch, n, yp, ldy))
        return y[0] if was1d else y
    # host arrays with a non-default padtype: library-owned device buffers around the device entry point
    y = np.empty((batch, n_out), dtype=np.float32)
    bx, by = _DevBuf(c, xo.nbytes).upload(xo), _DevBuf(c, y.nbytes)
    try:
        _check(_load().scir_b200_resample_poly_pad_f32(c.handle, _ptr(w), w.size, up, down, code, float(cval or 0.0),
                                                       bx.p, ld, batch, n, by.p, max(n_out, 1)))
        c.sync()
        by.download(y)
    finally:
        bx.free()
        by.free()
    return y[0] if was1d else y


_PAD = {"odd": L.PAD_ODD, "even": L.PAD_EVEN, "constant": L.PAD_CONSTANT, None: L.PAD_SCIPY_NONE}


def _filtfilt(b, x, mode, padlen, ctx):
    bt = _taps_f32(b)
    x2, was1d = _as_2d(x)
    dev, xo, xp, ld, batch, n, c = _prep(x2)
    c = ctx or c
    y, yp, ldy = _alloc_like(dev, xo, batch, n)
    fn = _load().scir_b200_filtfilt_fir_f32 if dev else _load().scir_b200_filtfilt_fir_f32_host
    rc = fn(c.handle, _ptr(bt), bt.size, mode, -1 if padlen is None else int(padlen), xp, ld, yp, ldy, batch, n)
    if rc == L.ERR_SHAPE:
        raise ValueError(L.last_error())                                     # SciPy raises ValueError, :4809
    _check(rc)
    return y[0] if was1d else y


def filtfilt(b, a, x, padtype="odd", padlen=None, *, ctx=None):
    """filtfilt(b, [1], x, padtype, padlen), method='pad', along the last axis."""
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    if a.size != 1:
        raise GpuError.backend_unavailable("only FIR numerators (a = [a0]) run on this path")
    if padtype not in _PAD:
        raise ValueError(f"Unknown value '{padtype}' given to padtype.")       # :4795-4797
    bt = _taps_f32(b) / np.float32(a[0])
    if padlen is not None:
        if int(padlen) != padlen:
            raise ValueError("padlen must be an integer or None")
        # the ABI's "default 3*ntaps" sentinel is padlen < 0; SciPy's _validate_pad (:4802-4816) takes a negative
        # padlen as "no extension" (`edge > 0` is false), so that is what a negative value means here too
        padlen = max(int(padlen), 0)
    return _filtfilt(bt, x, _PAD[padtype], padlen, ctx)


def filtfilt_zero_state(b, x, *, ctx=None):
    """The reference's filtfilt structure (sig/lib.rs:278-291) with an FIR numerator: zero-state
    forward pass, zero-state pass over the reversed result, reverse.  No padding."""
    return _filtfilt(b, x, L.PAD_ZERO_STATE, None, ctx)


# ---- f64 twins ------------------------------------------------------------------------------
# The reference's own `resample_poly` and `filtfilt` take and return Array1<f64> (sig/lib.rs:278-291, :313-362) and its
# fixtures are f64 (:638-668).  These functions serve float64 rows in float64 on the device (IEEE DFMA kernels,
# scir_b200/csrc/f64_routes.cu, fir_f64.cu) instead of down-casting them to the f32 path.  Inputs: float64 numpy arrays
# (uploaded / downloaded around the device call) or float64 CUDA tensors (device-resident).
def _as_2d_f64(x):
    if _is_torch(x):
        import torch
        if x.dtype != torch.float64 or not x.is_cuda:
            raise GpuError.shape_mismatch("device input must be a float64 CUDA tensor")
        x2, was1d = (x[None, :], True) if x.dim() == 1 else (x, False)
        if x2.dim() != 2:
            raise GpuError.shape_mismatch("x must be 1-D or 2-D")
        return x2.contiguous(), was1d
    a = np.ascontiguousarray(x, dtype=np.float64)
    if a.ndim not in (1, 2):
        raise GpuError.shape_mismatch("x must be 1-D or 2-D")
    return (a[None, :], True) if a.ndim == 1 else (a, False)


def _run_f64(x2, n_out, ctx, call):
    """call(ctx, x_ptr, ld_x, y_ptr, ld_y) -> rc on device pointers; handles numpy <-> device staging."""
    batch, n = x2.shape
    if _is_torch(x2):
        import torch
        c = ctx or G.torch_context(x2)
        y = torch.empty((batch, n_out), dtype=torch.float64, device=x2.device)
        rc = call(c, x2.data_ptr(), max(n, 1), y.data_ptr(), max(n_out, 1))
        return y, rc
    c = ctx or G.default_context()
    y = np.empty((batch, n_out), dtype=np.float64)
    bx, by = _DevBuf(c, x2.nbytes).upload(x2), _DevBuf(c, y.nbytes)
    try:
        rc = call(c, bx.p, max(n, 1), by.p, max(n_out, 1))
        if rc == L.OK:
            c.sync()
            by.download(y)
    finally:
        bx.free()
        by.free()
    return y, rc


def _taps_f64(t) -> np.ndarray:
    if _is_torch(t):
        t = t.detach().cpu().numpy()
    a = np.ascontiguousarray(t, dtype=np.float64)
    if a.ndim != 1 or a.size < 1:
        raise GpuError.shape_mismatch("taps must be 1-D with at least one element")
    return a


def upfirdn_f64(h, x, up=1, down=1, mode="constant", cval=0.0, *, ctx=None):
    """upfirdn(h, x, up, down, mode, cval) on float64 rows (scipy/signal/_upfirdn.py:107-216)."""
    if int(up) != up or int(down) != down:
        raise ValueError("up and down must be integers")
    up, down = int(up), int(down)
    if up < 1 or down < 1:
        raise ValueError("Both up and down must be >= 1")
    m = _mode_code(mode)
    ht = _taps_f64(h)
    x2, was1d = _as_2d_f64(x)
    n = x2.shape[1]
    if n < 1:
        raise ValueError("x must have at least one sample")
    lo = upfirdn_output_len(ht.size, n, up, down)
    y, rc = _run_f64(x2, lo, ctx, lambda c, xp, ldx, yp, ldy: _load().scir_b200_upfirdn_mode_f64(
        c.handle, _ptr(ht), ht.size, up, down, m, float(cval), xp, ldx, x2.shape[0], n, yp, ldy, 0, lo))
    _check(rc)
    return y[0] if was1d else y


def resample_poly_f64(x, up, down, window=("kaiser", 5.0), padtype="constant", cval=None, *, ctx=None):
    """resample_poly on float64 rows (scipy/signal/_signaltools.py:3865-3957), every padtype.  With the reference's
    31 legacy taps / 2 as `window` and (up, down) = (2, 3) this is scir_signal::resample_poly (sig/lib.rs:313-362)."""
    if int(up) != up or int(down) != down:
        raise ValueError("up and down must be integers")
    up, down = int(up), int(down)
    if up < 1 or down < 1:
        raise ValueError("up and down must be >= 1")
    if cval is not None and padtype != "constant":
        raise ValueError("cval has no effect when padtype is " + str(padtype))
    if padtype in _PAD_STATS:
        code = _PAD_STATS[padtype]
    elif padtype in UPFIRDN_MODES:
        code = UPFIRDN_MODES[padtype]
    else:
        raise ValueError("padtype must be one of: " + ", ".join(list(UPFIRDN_MODES) + list(_PAD_STATS)))
    if isinstance(window, (tuple, str)):
        if window != ("kaiser", 5.0):
            raise GpuError.backend_unavailable("pass the filter as an array; only the default ('kaiser', 5.0) design is built in")
        window = kaiser_lowpass(up, down)
    w = _taps_f64(window)
    x2, was1d = _as_2d_f64(x)
    n = x2.shape[1]
    plan = resample_poly_plan(n, w.size, up, down)
    n_out = n if (plan["up"] == 1 and plan["down"] == 1) else plan["n_out"]
    y, rc = _run_f64(x2, n_out, ctx, lambda c, xp, ldx, yp, ldy: _load().scir_b200_resample_poly_pad_f64(
        c.handle, _ptr(w), w.size, up, down, code, float(cval or 0.0), xp, ldx, x2.shape[0], n, yp, ldy))
    _check(rc)
    return y[0] if was1d else y


def filtfilt_f64(b, a, x, padtype="odd", padlen=None, *, ctx=None):
    """filtfilt(b, [a0], x, padtype, padlen), method='pad', on float64 rows (_signaltools.py:4745-4826).
    padtype='zero_state' is the reference's own structure (sig/lib.rs:278-291) with an FIR numerator."""
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    if a.size != 1:
        raise GpuError.backend_unavailable("only FIR numerators (a = [a0]) run on this path")
    if padtype == "zero_state":
        mode = L.PAD_ZERO_STATE
    elif padtype in _PAD:
        mode = _PAD[padtype]
    else:
        raise ValueError(f"Unknown value '{padtype}' given to padtype.")
    bt = _taps_f64(b) / a[0]
    if padlen is not None:
        padlen = max(int(padlen), 0)
    x2, was1d = _as_2d_f64(x)
    n = x2.shape[1]
    y, rc = _run_f64(x2, n, ctx, lambda c, xp, ldx, yp, ldy: _load().scir_b200_filtfilt_fir_f64(
        c.handle, _ptr(bt), bt.size, mode, -1 if padlen is None else padlen, xp, ldx, yp, ldy, x2.shape[0], n))
    if rc == L.ERR_SHAPE:
        raise ValueError(L.last_error())
    _check(rc)
    return y[0] if was1d else y
