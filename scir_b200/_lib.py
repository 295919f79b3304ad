"""ctypes binding of include/scir_b200.h.  Loads scir_b200/lib/libscir_b200.so (built in-tree by
scir_b200/build.py); raises loudly if it is missing -- there is no other implementation to fall
back to."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libscir_b200.so")

OK = 0
ERR_INVALID_ARG, ERR_NO_DEVICE, ERR_OOM, ERR_LAUNCH, ERR_SHAPE, ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6
MAX_TAPS = 7936
TAPS_SCIR, TAPS_LFILTER = 0, 1
PAD_ZERO_STATE, PAD_ODD, PAD_EVEN, PAD_CONSTANT, PAD_SCIPY_NONE = 0, 1, 2, 3, 4
(EXT_CONSTANT, EXT_SYMMETRIC, EXT_EDGE, EXT_SMOOTH, EXT_PERIODIC, EXT_REFLECT, EXT_ANTISYMMETRIC, EXT_ANTIREFLECT,
 EXT_LINE) = range(9)                                       # SciPy's MODE enum, _upfirdn_apply.pyx:77-86
PAD_STAT_MEAN, PAD_STAT_MEDIAN, PAD_STAT_MINIMUM, PAD_STAT_MAXIMUM = 16, 17, 18, 19

i64, vp, fp = C.c_int64, C.c_void_p, C.c_void_p   # float* passed as raw addresses


class ResamplePlan(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "up", "down", "n_out", "half_len", "n_pre_pad", "n_post_pad", "n_pre_remove",
        "len_h_padded", "upfirdn_len")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


# name -> (restype, argtypes); must list EVERY symbol include/scir_b200.h declares
# (tests/test_abi.py checks this table against the header and against the .so)
SIGNATURES = {
    "scir_b200_version": (C.c_char_p, []),
    "scir_b200_last_error": (C.c_char_p, []),
    "scir_b200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "scir_b200_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "scir_b200_ctx_create_on_stream": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    "scir_b200_ctx_destroy": (C.c_int, [vp]),
    "scir_b200_ctx_sync": (C.c_int, [vp]),
    "scir_b200_ctx_device": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "scir_b200_ctx_stream": (C.c_int, [vp, C.POINTER(vp)]),
    "scir_b200_ctx_set_option": (C.c_int, [vp, C.c_char_p, i64]),
    "scir_b200_ctx_get_option": (C.c_int, [vp, C.c_char_p, C.POINTER(i64)]),
    "scir_b200_ctx_launch_count": (C.c_int, [vp, C.POINTER(C.c_uint64)]),
    "scir_b200_malloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "scir_b200_free": (C.c_int, [vp, vp]),
    "scir_b200_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "scir_b200_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "scir_b200_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "scir_b200_host_free": (C.c_int, [vp]),
    "scir_b200_host_register": (C.c_int, [vp, C.c_size_t]),
    "scir_b200_host_unregister": (C.c_int, [vp]),
    "scir_b200_host_is_pinned": (C.c_int, [vp, C.c_size_t, C.POINTER(C.c_int)]),
    "scir_b200_current_device": (C.c_int, [C.POINTER(C.c_int)]),
    "scir_b200_fir1d_batched_f32": (C.c_int, [vp, fp, i64, fp, i64, C.c_int, fp, i64, i64, i64]),
    "scir_b200_fir1d_batched_f32_host": (C.c_int, [vp, fp, i64, fp, i64, C.c_int, fp, i64, i64, i64]),
    "scir_b200_fir1d_batched_f64": (C.c_int, [vp, vp, i64, vp, i64, C.c_int, vp, i64, i64, i64]),
    "scir_b200_lfilter_fir_f32": (C.c_int, [vp, fp, i64, C.c_float, fp, i64, fp, fp, fp, i64, i64, i64]),
    "scir_b200_upfirdn_out_len": (i64, [i64, i64, i64, i64]),
    "scir_b200_upfirdn_f32": (C.c_int, [vp, fp, i64, i64, i64, fp, i64, i64, i64, fp, i64, i64, i64]),
    "scir_b200_upfirdn_mode_f32": (C.c_int, [vp, fp, i64, i64, i64, C.c_int, C.c_float, fp, i64, i64, i64, fp, i64, i64, i64]),
    "scir_b200_resample_poly_pad_f32": (C.c_int, [vp, fp, i64, i64, i64, C.c_int, C.c_float, fp, i64, i64, i64, fp, i64]),
    "scir_b200_resample_poly_plan": (C.c_int, [i64, i64, i64, i64, C.POINTER(ResamplePlan)]),
    "scir_b200_resample_poly_f32": (C.c_int, [vp, fp, i64, i64, i64, fp, i64, i64, i64, fp, i64]),
    "scir_b200_resample_poly_f32_host": (C.c_int, [vp, fp, i64, i64, i64, fp, i64, i64, i64, fp, i64]),
    "scir_b200_filtfilt_fir_f32": (C.c_int, [vp, fp, i64, C.c_int, i64, fp, i64, fp, i64, i64, i64]),
    "scir_b200_filtfilt_fir_f32_host": (C.c_int, [vp, fp, i64, C.c_int, i64, fp, i64, fp, i64, i64, i64]),
    "scir_b200_upfirdn_mode_f64": (C.c_int, [vp, vp, i64, i64, i64, C.c_int, C.c_double, vp, i64, i64, i64, vp, i64, i64, i64]),
    "scir_b200_resample_poly_pad_f64": (C.c_int, [vp, vp, i64, i64, i64, C.c_int, C.c_double, vp, i64, i64, i64, vp, i64]),
    "scir_b200_filtfilt_fir_f64": (C.c_int, [vp, vp, i64, C.c_int, i64, vp, i64, vp, i64, i64, i64]),
    "scir_b200_add_scalar_f32": (C.c_int, [vp, fp, C.c_float, fp, i64]),
    "scir_b200_mul_scalar_f32": (C.c_int, [vp, fp, C.c_float, fp, i64]),
    "scir_b200_add_f32": (C.c_int, [vp, fp, fp, fp, i64]),
    "scir_b200_mg_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]),
    "scir_b200_mg_destroy": (C.c_int, [vp]),
    "scir_b200_mg_device_count": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "scir_b200_mg_ctx": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
    "scir_b200_mg_sync": (C.c_int, [vp]),
    "scir_b200_mg_fir1d_batched_f32": (C.c_int, [vp, C.POINTER(vp), C.POINTER(i64), fp, i64, C.c_int, C.POINTER(vp), C.POINTER(i64), i64, i64]),
    "scir_b200_mg_gather_rows_f32": (C.c_int, [vp, C.POINTER(vp), C.POINTER(i64), C.c_int, fp, i64, i64, i64]),
    "scir_b200_shard_rows": (C.c_int, [i64, C.c_int, C.c_int, C.POINTER(i64), C.POINTER(i64)]),
    "scir_b200_mg_fir1d_batched_f32_host": (C.c_int, [vp, fp, i64, fp, i64, C.c_int, fp, i64, i64, i64]),
    "scir_b200_mg_resample_poly_f32_host": (C.c_int, [vp, fp, i64, i64, i64, fp, i64, i64, i64, fp, i64]),
    "scir_b200_mg_filtfilt_fir_f32_host": (C.c_int, [vp, fp, i64, C.c_int, i64, fp, i64, fp, i64, i64, i64]),
    "scir_b200_microbench_ffma": (C.c_int, [vp, C.c_int, C.POINTER(C.c_double)]),
    "scir_b200_microbench_ffma2": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "scir_b200_microbench_copy": (C.c_int, [vp, C.c_size_t, C.c_int, C.POINTER(C.c_double)]),
    "scir_b200_microbench_pcie": (C.c_int, [vp, C.c_size_t, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    """The loaded CUDA library.  Never substitutes anything else."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} not found: build it with `python -m scir_b200.build` "
                "(nvcc, sm_100a). scir_b200 has no CPU or PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def last_error() -> str:
    return lib().scir_b200_last_error().decode("utf-8", "replace")
