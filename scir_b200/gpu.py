"""Host-side mirror of crate `scir-gpu` for the batched-FIR path (crates/scir-gpu/src/lib.rs).

Same names, argument meaning and error behaviour as the reference's Rust surface, so the parity
tests read like the reference's own tests (lib.rs:1249-1323):

    Device, DType, GpuError, DeviceArray            lib.rs:16-35, 57-84, 95-190
    fir1d_batched_f32_cuda(x, taps) -> Result       lib.rs:1036-1113
    fir1d_batched_f32_auto(x, taps, device)         lib.rs:515-531

Differences, all required by BASELINE.json's north_star:
  * no silent CPU fallback (lib.rs:520-523): Device.Cuda raises GpuError if the CUDA library or a
    B200 is missing; `fir1d_batched_f32_auto` is infallible in Rust, so the binding panics there --
    here it raises.
  * Device.Cpu, when asked for EXPLICITLY, runs the crate's own CPU function `fir1d_batched_f32`
    (lib.rs:1134-1152, restated below) exactly as `_auto` does in the reference (lib.rs:517-518).  It is
    never a fallback: nothing on the Device.Cuda arm can reach it.
  * DeviceArray is genuinely device-backed after `to_device(Device.Cuda)` (lib.rs:177 is a placeholder).

Arrays: numpy (host; goes through the *_host ABI calls: H2D + kernel + D2H) or torch CUDA tensors
(device-resident; the kernel runs on torch's current stream).  PyTorch is plumbing here -- device
memory and streams -- never the compute path.
"""
from __future__ import annotations

import ctypes as C
import enum
import threading
import weakref

import numpy as np

from . import _lib as L


class DType(enum.Enum):
    """lib.rs:16-22"""
    F32 = "f32"
    F64 = "f64"


class Device(enum.Enum):
    """lib.rs:25-35 (the wgpu arm is out of scope: north_star says no multi-backend dispatch)"""
    Cpu = "cpu"
    Cuda = "cuda"


class GpuError(Exception):
    """lib.rs:57-74: BackendUnavailable(String) | ShapeMismatch"""

    def __init__(self, kind: str, message: str = "", code: int = 0):
        self.kind, self.message, self.code = kind, message, code
        super().__init__(f"backend not available: {message}" if kind == "BackendUnavailable" else
                         ("shape mismatch" + (f": {message}" if message else "")))

    @staticmethod
    def backend_unavailable(msg, code=L.ERR_NO_DEVICE):
        return GpuError("BackendUnavailable", msg, code)

    @staticmethod
    def shape_mismatch(msg=""):
        return GpuError("ShapeMismatch", msg, L.ERR_SHAPE)


def _check(rc: int):
    """Map an ABI return code to GpuError the way the Rust binding does."""
    if rc == L.OK:
        return
    msg = L.last_error()
    if rc == L.ERR_SHAPE:
        raise GpuError.shape_mismatch(msg)
    if rc == L.ERR_INVALID_ARG:
        raise ValueError(msg)
    raise GpuError.backend_unavailable(msg, rc)


def _load():
    try:
        return L.lib()
    except (L.LibraryMissing, OSError) as e:
        raise GpuError.backend_unavailable(str(e)) from e


def device_count() -> int:
    n = C.c_int(0)
    _check(_load().scir_b200_device_count(C.byref(n)))
    return n.value


class Context:
    """A long-lived handle (device + stream + scratch): replaces the per-call CudaCtx, lib.rs:601-622."""

    def __init__(self, device: int = 0, stream: int | None = None):
        lib = _load()
        h = C.c_void_p()
        if stream is None:
            _check(lib.scir_b200_ctx_create(device, C.byref(h)))
        else:
            _check(lib.scir_b200_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h)))
        self.handle, self.device_index = h, device
        self._fin = weakref.finalize(self, lib.scir_b200_ctx_destroy, h)

    def sync(self):
        _check(_load().scir_b200_ctx_sync(self.handle))

    def set_option(self, key: str, value: int):
        _check(_load().scir_b200_ctx_set_option(self.handle, key.encode(), int(value)))

    def get_option(self, key: str) -> int:
        v = C.c_int64()
        _check(_load().scir_b200_ctx_get_option(self.handle, key.encode(), C.byref(v)))
        return v.value

    def launch_count(self) -> int:
        v = C.c_uint64()
        _check(_load().scir_b200_ctx_launch_count(self.handle, C.byref(v)))
        return v.value

    def close(self):
        self._fin()


_ctx_lock = threading.Lock()
_ctx_cache: dict = {}


def current_device() -> int:
    """The calling thread's current CUDA device (cudaGetDevice): what torch.cuda.set_device / torchrun's LOCAL_RANK
    binding selected.  Raises GpuError when there is no GPU."""
    d = C.c_int(-1)
    _check(_load().scir_b200_current_device(C.byref(d)))
    return d.value


def default_context(device: int | None = None) -> Context:
    """One library-owned-stream ctx per (thread, device).  `device=None` means the thread's CURRENT device, so host
    arrays in a multi-GPU process (one rank per GPU) are filtered on the rank's own GPU, not on GPU 0."""
    if device is None:
        device = current_device()
    key = ("own", threading.get_ident(), device)
    with _ctx_lock:
        if key not in _ctx_cache:
            _ctx_cache[key] = Context(device)
        return _ctx_cache[key]


def torch_context(tensor) -> Context:
    """A ctx borrowing torch's CURRENT stream on the tensor's device (kernels are stream-ordered
    with the surrounding torch work; nothing is synchronised)."""
    import torch
    dev = tensor.device.index if tensor.device.index is not None else torch.cuda.current_device()
    stream = torch.cuda.current_stream(dev).cuda_stream
    key = ("torch", threading.get_ident(), dev, stream)
    with _ctx_lock:
        if key not in _ctx_cache:
            _ctx_cache[key] = Context(dev, stream)
        return _ctx_cache[key]


def _is_torch(a) -> bool:
    return type(a).__module__.startswith("torch")


def _taps_f32(taps) -> np.ndarray:
    if _is_torch(taps):
        taps = taps.detach().cpu().numpy()
    t = np.ascontiguousarray(taps, dtype=np.float32)
    if t.ndim != 1:
        raise GpuError.shape_mismatch("taps must be 1-D")
    return t


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def _host_matrix(x) -> np.ndarray:
    a = np.ascontiguousarray(x, dtype=np.float32)       # the reference deep-copies too, lib.rs:1042
    if a.ndim != 2:
        raise GpuError.shape_mismatch("x must be 2-D (batch, n)")
    return a


def _check_out_host(out, shape) -> int:
    """Validates a caller-provided host result array and returns its row pitch in elements."""
    if not isinstance(out, np.ndarray) or out.dtype != np.float32 or out.ndim != 2 or tuple(out.shape) != tuple(shape):
        raise GpuError.shape_mismatch(f"out must be a float32 numpy array of shape {tuple(shape)}")
    if not out.flags.writeable:
        raise GpuError.shape_mismatch("out is read-only")
    b, n = shape
    if n > 1 and out.strides[1] != 4:
        raise GpuError.shape_mismatch("out rows must be contiguous (stride 1 along the samples)")
    if b > 1 and (out.strides[0] % 4 != 0 or out.strides[0] < 4 * n):
        raise GpuError.shape_mismatch("out row pitch must be a non-negative multiple of 4 bytes >= the row length")
    return out.strides[0] // 4 if b > 1 else max(n, 1)


def _check_out_torch(out, x) -> int:
    import torch
    if not _is_torch(out) or out.dtype != torch.float32 or out.dim() != 2 or tuple(out.shape) != tuple(x.shape):
        raise GpuError.shape_mismatch(f"out must be a float32 tensor of shape {tuple(x.shape)}")
    if not out.is_cuda or out.device != x.device:
        raise GpuError.shape_mismatch(f"out must live on {x.device}, not {out.device}")
    b, n = x.shape
    if n > 1 and out.stride(1) != 1:
        raise GpuError.shape_mismatch("out rows must be contiguous (stride 1 along the samples)")
    if b > 1 and out.stride(0) < n:
        raise GpuError.shape_mismatch("out row pitch is shorter than a row")
    return out.stride(0) if b > 1 else max(n, 1)


def _torch_matrix(x):
    import torch
    if x.dtype != torch.float32 or x.dim() != 2 or not x.is_cuda:
        raise GpuError.shape_mismatch("device input must be a 2-D float32 CUDA tensor")
    if x.shape[1] > 1 and x.stride(1) != 1:
        x = x.contiguous()
    ld = x.stride(0) if x.shape[0] > 1 else max(x.shape[1], 1)
    return x, ld


def fir1d_batched_f32_cuda(x, taps, *, ctx: Context | None = None, out=None,
                           tap_order: int = L.TAPS_SCIR):
    """lib.rs:1036-1113.  y[b,i] = sum_t taps[k-1-t] * x[b,i-t]; returns an array shaped like x
    or raises GpuError (the Rust fn returns Result<Array2<f32>, GpuError>)."""
    lib = _load()
    t = _taps_f32(taps)
    if _is_torch(x):
        import torch
        xt, ldx = _torch_matrix(x)
        c = ctx or torch_context(xt)
        b, n = xt.shape
        if out is None:
            y = torch.empty_like(xt, memory_format=torch.contiguous_format)
            ldy = max(n, 1)
        else:
            y, ldy = out, _check_out_torch(out, xt)
        _check(lib.scir_b200_fir1d_batched_f32(c.handle, xt.data_ptr(), ldx, _ptr(t), t.size, tap_order,
                                               y.data_ptr(), ldy, b, n))
        return y
    a = _host_matrix(x)
    c = ctx or default_context()
    b, n = a.shape
    if out is None:
        y, ldy = np.empty_like(a), max(n, 1)
    else:
        y, ldy = out, _check_out_host(out, a.shape)
    _check(lib.scir_b200_fir1d_batched_f32_host(c.handle, _ptr(a), max(n, 1), _ptr(t), t.size, tap_order,
                                                _ptr(y), ldy, b, n))
    return y


def fir1d_batched_f64_cuda(x, taps, *, ctx: Context | None = None, tap_order: int = L.TAPS_SCIR):
    """Device twin of fir1d_batched_f64 (lib.rs:1166-1184): x is a 2-D float64 CUDA tensor (result: same) or a
    float64 numpy array (uploaded, filtered, downloaded)."""
    lib = _load()
    t = np.ascontiguousarray(taps, dtype=np.float64)
    if t.ndim != 1 or t.size < 1:
        raise GpuError.shape_mismatch("taps must be 1-D with at least one element")
    if _is_torch(x):
        import torch
        if x.dtype != torch.float64 or x.dim() != 2 or not x.is_cuda:
            raise GpuError.shape_mismatch("device input must be a 2-D float64 CUDA tensor")
        xt = x.contiguous()
        c = ctx or torch_context(xt)
        y = torch.empty_like(xt)
        b, n = xt.shape
        _check(lib.scir_b200_fir1d_batched_f64(c.handle, xt.data_ptr(), max(n, 1), _ptr(t), t.size, tap_order,
                                               y.data_ptr(), max(n, 1), b, n))
        return y
    a = np.ascontiguousarray(x, dtype=np.float64)
    if a.ndim != 2:
        raise GpuError.shape_mismatch("x must be 2-D (batch, n)")
    c = ctx or default_context()
    b, n = a.shape
    y = np.empty_like(a)
    px, py = C.c_void_p(), C.c_void_p()
    _check(lib.scir_b200_malloc(c.handle, max(a.nbytes, 8), C.byref(px)))
    _check(lib.scir_b200_malloc(c.handle, max(a.nbytes, 8), C.byref(py)))
    try:
        _check(lib.scir_b200_memcpy_h2d(c.handle, px, _ptr(a), a.nbytes))
        _check(lib.scir_b200_fir1d_batched_f64(c.handle, px, max(n, 1), _ptr(t), t.size, tap_order, py, max(n, 1), b, n))
        _check(lib.scir_b200_memcpy_d2h(c.handle, _ptr(y), py, y.nbytes))
    finally:
        lib.scir_b200_free(c.handle, px)
        lib.scir_b200_free(c.handle, py)
    return y


def fir1d_batched_f32(x, taps) -> np.ndarray:
    """The crate's CPU function (lib.rs:1134-1152), kept because `fir1d_batched_f32_auto(.., Device::Cpu)` and the
    crate's public surface name it: y[b,i] = sum_{t=0}^{min(i,k-1)} taps[k-1-t] * x[b,i-t], f32 multiply then f32
    add, newest sample first.  It is reached ONLY when the caller asks for Device.Cpu explicitly -- never from the
    Device.Cuda arm, which raises when the GPU path cannot run.  Vectorised over the samples with the reference's
    per-element operation order, so the values are bit-identical to its loop."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    t = np.ascontiguousarray(taps, dtype=np.float32)
    if a.ndim != 2 or t.ndim != 1:
        raise GpuError.shape_mismatch("x must be 2-D (batch, n) and taps 1-D")
    n, k = a.shape[1], t.size
    y = np.zeros_like(a)
    for d in range(min(k, n)):                              # d = t in lib.rs:1141: newest sample first
        y[:, d:] += t[k - 1 - d] * a[:, : n - d]            # f32 product rounded, then f32 sum rounded (no FMA)
    return y


def fir1d_batched_f32_auto(x, taps, device: Device, **kw):
    """lib.rs:515-531, minus the silent fallback: Device.Cuda runs on the B200 or raises; Device.Cpu is the crate's
    own CPU function, as in the reference (lib.rs:517-518)."""
    if device == Device.Cuda:
        return fir1d_batched_f32_cuda(x, taps, **kw)
    if device == Device.Cpu:
        if _is_torch(x):
            raise GpuError.shape_mismatch("Device.Cpu takes host arrays")
        return fir1d_batched_f32(x, taps)
    raise GpuError.backend_unavailable(f"unknown device {device!r}")


class DeviceArray:
    """lib.rs:77-190 with real device storage.  f32 only on the device (the FIR path's dtype)."""

    def __init__(self, shape, dtype: DType, host: np.ndarray):
        self._shape, self._dtype, self._device = list(shape), dtype, Device.Cpu
        self._host, self._dptr, self._ctx = host, None, None

    @classmethod
    def from_cpu_slice(cls, shape, dtype: DType, data):
        npdt = np.float32 if dtype == DType.F32 else np.float64
        host = np.array(data, dtype=npdt).reshape(-1)
        if int(np.prod(shape)) != host.size:                      # assert_eq! at lib.rs:96
            raise AssertionError("shape product != data length")
        return cls(shape, dtype, host)

    def shape(self):
        return list(self._shape)

    def dtype(self):
        return self._dtype

    def device(self):
        return self._device

    def to_cpu_vec(self):
        if self._device == Device.Cuda:
            out = np.empty(self._host.size, dtype=np.float32)
            _check(_load().scir_b200_memcpy_d2h(self._ctx.handle, _ptr(out), self._dptr, out.nbytes))
            return out.tolist()
        return self._host.tolist()

    def to_device(self, device: Device, ctx: Context | None = None):
        if device == self._device:
            return
        lib = _load()
        if device == Device.Cuda:
            if self._dtype != DType.F32:
                raise GpuError.backend_unavailable("only f32 arrays live on the device")
            self._ctx = ctx or default_context()
            p = C.c_void_p()
            _check(lib.scir_b200_malloc(self._ctx.handle, max(self._host.nbytes, 4), C.byref(p)))
            _check(lib.scir_b200_memcpy_h2d(self._ctx.handle, p, _ptr(self._host), self._host.nbytes))
            self._dptr, self._device = p, Device.Cuda
            self._fin = weakref.finalize(self, lib.scir_b200_free, self._ctx.handle, p)
        else:
            self._host = np.asarray(self.to_cpu_vec(), dtype=np.float32)
            self._fin()
            self._dptr, self._device = None, Device.Cpu

    # ---- elementwise ops with device dispatch (lib.rs:268-377) -------------------------------------------
    def _new_like(self) -> "DeviceArray":
        lib = _load()
        out = DeviceArray(self._shape, self._dtype, np.empty(0, np.float32))
        p = C.c_void_p()
        _check(lib.scir_b200_malloc(self._ctx.handle, max(self._host.size * 4, 4), C.byref(p)))
        out._dptr, out._device, out._ctx = p, Device.Cuda, self._ctx
        out._host = np.empty(self._host.size, np.float32)
        out._fin = weakref.finalize(out, lib.scir_b200_free, self._ctx.handle, p)
        return out

    def _need_cuda(self, what):
        if self._device != Device.Cuda:
            raise GpuError.backend_unavailable(
                f"{what}: operands live on different devices; move them with to_device() first")

    def _host_result(self, values) -> "DeviceArray":
        return DeviceArray(self._shape, self._dtype, np.asarray(values, dtype=self._host.dtype))

    def add_scalar_auto(self, alpha: float) -> "DeviceArray":
        """lib.rs:268-301: Device::Cuda arm (add_scalar_f32_cuda :912-972) on device-resident data; a Device.Cpu
        array runs the crate's own loop (:206-221).  The array's device decides; neither arm falls back to the other."""
        if self._device == Device.Cpu:
            return self._host_result(self._host + self._host.dtype.type(alpha))
        self._need_cuda("add_scalar_auto")
        out = self._new_like()
        _check(_load().scir_b200_add_scalar_f32(self._ctx.handle, self._dptr, float(alpha), out._dptr, self._host.size))
        return out

    def mul_scalar_auto(self, alpha: float) -> "DeviceArray":
        """lib.rs:355-388, Device::Cuda arm (mul_scalar_f32_cuda :974-1034); Device.Cpu: the loop :240-255."""
        if self._device == Device.Cpu:
            return self._host_result(self._host * self._host.dtype.type(alpha))
        self._need_cuda("mul_scalar_auto")
        out = self._new_like()
        _check(_load().scir_b200_mul_scalar_f32(self._ctx.handle, self._dptr, float(alpha), out._dptr, self._host.size))
        return out

    def add_auto(self, other: "DeviceArray") -> "DeviceArray":
        """lib.rs:303-353: shapes must match (ShapeMismatch :304-306), both arrays on the same device."""
        if self._shape != other._shape:
            raise GpuError.shape_mismatch("add_auto: shapes differ")
        if self._device == Device.Cpu and other._device == Device.Cpu:
            return self._host_result(self._host + other._host)
        self._need_cuda("add_auto")
        other._need_cuda("add_auto")
        out = self._new_like()
        _check(_load().scir_b200_add_f32(self._ctx.handle, self._dptr, other._dptr, out._dptr, self._host.size))
        return out

    def fir1d_batched(self, taps, tap_order: int = L.TAPS_SCIR) -> "DeviceArray":
        """Device-resident FIR: no PCIe traffic when chaining."""
        if self._device != Device.Cuda or len(self._shape) != 2:
            raise GpuError.shape_mismatch("need a 2-D array on Device.Cuda")
        lib, t = _load(), _taps_f32(taps)
        b, n = self._shape
        out = DeviceArray(self._shape, self._dtype, np.empty(0, np.float32))
        p = C.c_void_p()
        _check(lib.scir_b200_malloc(self._ctx.handle, max(b * n * 4, 4), C.byref(p)))
        out._dptr, out._device, out._ctx = p, Device.Cuda, self._ctx
        out._host = np.empty(b * n, np.float32)
        out._fin = weakref.finalize(out, lib.scir_b200_free, self._ctx.handle, p)
        _check(lib.scir_b200_fir1d_batched_f32(self._ctx.handle, self._dptr, max(n, 1), _ptr(t), t.size,
                                               tap_order, p, max(n, 1), b, n))
        return out
