"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports exactly the symbols
include/scir_b200.h declares, the ctypes table covers them all, the pure-integer entry points are
bit-exact with the oracle / SciPy, and compute entry points fail LOUDLY without a GPU (no fallback).
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from scir_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "scir_b200.h")).read()
    return sorted(set(re.findall(r"SCIR_B200_API[^;(]*?\b(scir_b200_\w+)\s*\(", src)))


def test_library_builds_and_loads():
    from scir_b200 import build
    path = build.build()
    assert os.path.exists(path)
    assert L.lib().scir_b200_version().startswith(b"scir-b200")


def test_exports_match_header():
    syms = header_symbols()
    assert len(syms) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (scir_b200_\w+)", out)))
    assert exported == syms
    assert sorted(L.SIGNATURES) == syms
    lib = L.lib()
    for s in syms:
        assert getattr(lib, s) is not None


def test_only_sm100a_code_in_library():
    out = subprocess.run(["cuobjdump", "-lelf", L.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_upfirdn_out_len_bit_exact():
    lib = L.lib()
    assert lib.scir_b200_upfirdn_out_len(1001, 10**8, 320, 441) == 72562360     # test_upfirdn.py:311-322
    rng = np.random.RandomState(0)
    for _ in range(500):
        a = [int(v) for v in (rng.randint(1, 5000), rng.randint(1, 10**7), rng.randint(1, 400), rng.randint(1, 400))]
        assert lib.scir_b200_upfirdn_out_len(*a) == O.upfirdn_out_len(*a)


def test_resample_plan_bit_exact():
    lib = L.lib()
    rng = np.random.RandomState(1)
    cases = [(2**20, 96, 3, 2), (32, 31, 2, 3), (10**7, 2001, 160, 147), (5, 1, 1, 1), (7, 3, 4, 4)]
    cases += [tuple(int(v) for v in (rng.randint(1, 10**5), rng.randint(1, 800), rng.randint(1, 60), rng.randint(1, 60)))
              for _ in range(500)]
    for n_in, len_h, up, down in cases:
        p = L.ResamplePlan()
        assert lib.scir_b200_resample_poly_plan(n_in, len_h, up, down, C.byref(p)) == 0
        assert p.as_dict() == O.resample_poly_plan(n_in, len_h, up, down)
    p = L.ResamplePlan()
    assert lib.scir_b200_resample_poly_plan(10, 5, 0, 1, C.byref(p)) == L.ERR_INVALID_ARG
    assert b"up and down" in lib.scir_b200_last_error()


def test_shard_rows_partition():
    lib = L.lib()
    for batch in (0, 1, 7, 8, 1024, 1025, 8191):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for rank in range(world):
                a, b = C.c_int64(), C.c_int64()
                assert lib.scir_b200_shard_rows(batch, world, rank, C.byref(a), C.byref(b)) == 0
                assert 0 <= b.value - a.value <= batch // world + 1
                covered += list(range(a.value, b.value))
            assert covered == list(range(batch))
    a, b = C.c_int64(), C.c_int64()
    assert lib.scir_b200_shard_rows(8, 2, 2, C.byref(a), C.byref(b)) == L.ERR_INVALID_ARG


def _no_gpu():
    n = C.c_int(0)
    rc = L.lib().scir_b200_device_count(C.byref(n))
    return rc != 0 or n.value == 0


@pytest.mark.skipif(not _no_gpu(), reason="a GPU is present; this checks the no-GPU behaviour")
def test_fails_loudly_without_gpu():
    """north_star: 'Device::Cuda errors if no GPU is present' -- never a CPU result."""
    from scir_b200 import gpu, signal
    x = np.ones((2, 8), np.float32)
    taps = np.ones(3, np.float32)
    with pytest.raises(gpu.GpuError) as e:
        gpu.fir1d_batched_f32_cuda(x, taps)
    assert e.value.kind == "BackendUnavailable"
    with pytest.raises(gpu.GpuError):
        gpu.fir1d_batched_f32_auto(x, taps, gpu.Device.Cuda)
    with pytest.raises(gpu.GpuError):
        signal.gpu.fir1d_batched_f32(x, taps, gpu.Device.Cuda)
    for fn in (lambda: signal.lfilter(taps, [1.0], x), lambda: signal.upfirdn(taps, x, 3, 2),
               lambda: signal.resample_poly(x, 3, 2, taps), lambda: signal.filtfilt_zero_state(taps, x)):
        with pytest.raises(gpu.GpuError):
            fn()
    h = C.c_void_p()
    assert L.lib().scir_b200_ctx_create(0, C.byref(h)) == L.ERR_NO_DEVICE
    assert L.last_error() != ""


def test_device_cpu_arm_is_the_crates_own_loop():
    """lib.rs:517-518: _auto(.., Device::Cpu) is the crate's CPU function.  Asked for explicitly it runs (bit-exact
    with the reference's loop, pinned here on its known-answer vector lib.rs:1251-1260 and on the oracle); it is
    never reached from Device.Cuda, which raises without a GPU (test above)."""
    from scir_b200 import gpu
    x = np.array([[1, 2, 3, 4], [0.5, 0, -0.5, -1]], np.float32)
    y = gpu.fir1d_batched_f32_auto(x, np.array([0.25, 0.5, 0.25], np.float32), gpu.Device.Cpu)
    np.testing.assert_allclose(y, [[0.25, 1, 2, 3], [0.125, 0.25, 0, -0.5]], atol=1e-7)
    rng = np.random.RandomState(3)
    for n, k in ((1, 1), (5, 9), (300, 31), (1000, 63)):
        xr = rng.randn(3, n).astype(np.float32)
        t = rng.randn(k).astype(np.float32)
        assert np.array_equal(gpu.fir1d_batched_f32(xr, t), O.fir1d_batched_f32(xr, t))


def test_out_argument_is_validated():
    """a wrong-dtype / wrong-shape / strided `out` must be rejected before any pointer is taken."""
    from scir_b200 import gpu
    for bad in (np.empty((2, 8), np.float64), np.empty((2, 7), np.float32), np.empty((8, 2), np.float32).T,
                np.empty((2, 16), np.float32)[:, ::2], [0.0] * 16):
        with pytest.raises(gpu.GpuError) as e:
            gpu._check_out_host(bad, (2, 8))
        assert e.value.kind == "ShapeMismatch"
    ro = np.empty((2, 8), np.float32)
    ro.flags.writeable = False
    with pytest.raises(gpu.GpuError):
        gpu._check_out_host(ro, (2, 8))
    assert gpu._check_out_host(np.empty((2, 8), np.float32), (2, 8)) == 8
    assert gpu._check_out_host(np.empty((2, 12), np.float32)[:, :8], (2, 8)) == 12


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under scir_b200/ may reference it."""
    pkg = os.path.join(ROOT, "scir_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in txt.lower(), os.path.join(dp, f)


def test_kaiser_lowpass_matches_scipy_firwin():
    from scipy import signal as ss
    from scir_b200 import signal
    for up, down in ((3, 2), (2, 3), (160, 147), (4, 2), (7, 1)):
        g = np.gcd(up, down)
        mr = max(up // g, down // g)
        np.testing.assert_allclose(signal.kaiser_lowpass(up, down),
                                   ss.firwin(2 * 10 * mr + 1, 1.0 / mr, window=("kaiser", 5.0)), atol=1e-15)
