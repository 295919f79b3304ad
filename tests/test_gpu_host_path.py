"""GPU tests of the *_host entry points' streaming ring (scir_b200/csrc/api.cu: host_pipeline): pageable caller arrays
through the pinned ring + copy threads, pinned / registered arrays by direct DMA, strided rows, more blocks than ring
slots (every slot, event and ticket is reused several times), every `host_stage` route -- all against the device-pointer
path on the same data, which the parity tests pin to the oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]

torch = pytest.importorskip("torch")

from oracle import oracle as O                      # noqa: E402  (checker only)
from scir_b200 import _lib as L                     # noqa: E402
from scir_b200 import gpu, signal                   # noqa: E402


def _fir_host(ctx, x, taps, y):
    lib = L.lib()
    b, n = x.shape
    rc = lib.scir_b200_fir1d_batched_f32_host(ctx.handle, x.ctypes.data, x.strides[0] // 4, taps.ctypes.data, taps.size, L.TAPS_SCIR,
                                              y.ctypes.data, y.strides[0] // 4, b, n)
    assert rc == 0, L.last_error()


@pytest.mark.parametrize("stage,block_rows", [(1, 3), (1, 1), (1, 0), (0, 3), (2, 3)])
def test_pageable_rows_through_every_route(stage, block_rows):
    rng = np.random.RandomState(stage * 10 + block_rows)
    rows, n = 37, 30001
    xbuf = (rng.rand(rows, n + 15).astype(np.float32) * 2 - 1)       # row pitch > n: strided 2-D copies
    x = xbuf[:, :n]
    ybuf = np.full((rows, n + 7), np.float32(-7.0))
    y = ybuf[:, :n]
    taps = rng.randn(63).astype(np.float32)
    ctx = gpu.Context(0)
    ctx.set_option("host_stage", stage)
    ctx.set_option("host_block_rows", block_rows)
    ctx.set_option("host_copy_threads", 3)
    pin = C.c_int(-1)
    assert L.lib().scir_b200_host_is_pinned(x.ctypes.data, 4 * n, C.byref(pin)) == 0 and pin.value == 0
    for _ in range(2):                                               # twice: the ring's buffers and events are reused
        _fir_host(ctx, x, taps, y)
    want = gpu.fir1d_batched_f32_cuda(torch.from_numpy(np.ascontiguousarray(x)).cuda(), taps).cpu().numpy()
    assert np.array_equal(y, want)                                   # same kernel family per row block: bit-equal
    assert np.all(ybuf[:, n:] == -7.0)                               # the pitch gap of the output is never written
    assert ctx.get_option("host_staged_calls") == (2 if stage == 1 else 0)
    assert ctx.get_option("host_registered_calls") == (2 if stage == 2 else 0)
    assert np.abs(y - O.fir1d_batched_f32_acc64(np.ascontiguousarray(x), taps)).max() <= 1e-5 * np.abs(taps).sum()


def test_registered_and_library_pinned_arrays_are_dma_targets():
    lib = L.lib()
    rng = np.random.RandomState(4)
    rows, n = 64, 65536
    taps = rng.randn(31).astype(np.float32)
    x = (rng.rand(rows, n).astype(np.float32) * 2 - 1)
    y = np.empty_like(x)
    ctx = gpu.Context(0)
    ctx.set_option("host_block_rows", 5)
    assert lib.scir_b200_host_register(x.ctypes.data, x.nbytes) == 0, L.last_error()
    assert lib.scir_b200_host_register(y.ctypes.data, y.nbytes) == 0, L.last_error()
    try:
        pin = C.c_int(0)
        assert lib.scir_b200_host_is_pinned(x.ctypes.data, x.nbytes, C.byref(pin)) == 0 and pin.value == 1
        _fir_host(ctx, x, taps, y)
        assert ctx.get_option("host_staged_calls") == 0              # pinned both ways: no staging
        want = gpu.fir1d_batched_f32_cuda(torch.from_numpy(x).cuda(), taps).cpu().numpy()
        assert np.array_equal(y, want)
        # pinned input, pageable output: only the output side is staged
        y2 = np.empty_like(x)
        _fir_host(ctx, x, taps, y2)
        assert ctx.get_option("host_staged_calls") == 1 and np.array_equal(y2, want)
    finally:
        assert lib.scir_b200_host_unregister(x.ctypes.data) == 0
        assert lib.scir_b200_host_unregister(y.ctypes.data) == 0
    hp = C.c_void_p()
    assert lib.scir_b200_host_alloc(x.nbytes, C.byref(hp)) == 0
    try:
        ax = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_float)), shape=(rows, n))
        ax[:] = x
        y3 = np.empty_like(x)
        _fir_host(ctx, ax, taps, y3)
        assert np.array_equal(y3, want)
    finally:
        lib.scir_b200_host_free(hp)


def test_resample_and_filtfilt_host_routes_on_pageable_arrays():
    """Different input and output row lengths (resample_poly) and the scratch-carrying route (filtfilt) through the ring."""
    from scipy.signal import firwin
    rng = np.random.RandomState(6)
    x = (rng.rand(29, 50001).astype(np.float32) * 2 - 1)
    h = firwin(96, 1.0 / 3.0, window=("kaiser", 5.0)).astype(np.float32)
    b = firwin(63, 0.2).astype(np.float32)
    ctx = gpu.Context(0)
    ctx.set_option("host_block_rows", 4)
    yr = signal.resample_poly(x, 3, 2, h, ctx=ctx)
    yf = signal.filtfilt(b, [1.0], x, ctx=ctx)
    assert ctx.get_option("host_staged_calls") == 2
    xd = torch.from_numpy(x).cuda()
    assert np.array_equal(yr, signal.resample_poly(xd, 3, 2, h).cpu().numpy())
    assert np.array_equal(yf, signal.filtfilt(b, [1.0], xd).cpu().numpy())


def test_host_errors_leave_no_work_behind():
    """A failing block (filtfilt padlen > n: ERR_SHAPE from the first block) must return the error with every queued copy job
    finished -- the caller may free its arrays right after."""
    lib = L.lib()
    x = np.ones((8, 100), np.float32)
    y = np.empty_like(x)
    b = np.ones(50, np.float32)
    ctx = gpu.Context(0)
    rc = lib.scir_b200_filtfilt_fir_f32_host(ctx.handle, b.ctypes.data, b.size, L.PAD_ODD, -1, x.ctypes.data, 100, y.ctypes.data, 100, 8, 100)
    assert rc == L.ERR_SHAPE and "padlen" in L.last_error()
    del x, y
    z = np.ones((4, 1000), np.float32)
    out = gpu.fir1d_batched_f32_cuda(z, np.ones(2, np.float32), ctx=ctx)   # the ctx still works
    assert out[0, 5] == 2.0
