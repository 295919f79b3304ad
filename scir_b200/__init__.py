"""scir_b200 -- B200-native (sm_100a) batched real-f32 FIR for SciR.

One hot path, nothing else: `scir-gpu`'s batched FIR entry point and the `scir-signal` FIR routes
onto it (lfilter with a=[1], upfirdn / resample_poly, filtfilt with an FIR numerator), as
hand-written CUDA behind a C ABI (include/scir_b200.h -> scir_b200/lib/libscir_b200.so).

  scir_b200.gpu      mirror of crate `scir-gpu`     (Device, DType, GpuError, DeviceArray, fir1d_batched_f32_*)
  scir_b200.signal   mirror of crate `scir-signal`  (gpu.fir1d_batched_f32, lfilter, upfirdn, resample_poly, filtfilt)
  scir_b200.dist     one-process-per-GPU row sharding over torch.distributed
  scir_b200._lib     the ctypes binding of the C ABI

There is no CPU fallback: importing works anywhere, but every compute call raises GpuError when the
CUDA library or a B200 is missing.
"""
from . import _lib, gpu, signal  # noqa: F401
from .gpu import Device, DType, DeviceArray, GpuError  # noqa: F401

__version__ = "0.1.0"
