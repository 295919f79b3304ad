"""Times the f64 device twin (scir_b200_fir1d_batched_f64) at 256 x 2^20 for a few tap counts.
Usage (GPU box): python tools/time_f64.py"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from scir_b200 import gpu
x = torch.rand((256, 1 << 20), device="cuda", dtype=torch.float64) * 2 - 1
for k in (31, 63, 255):
    taps = np.random.RandomState(k).randn(k)
    ctx = gpu.Context(0)
    for _ in range(2):
        gpu.fir1d_batched_f64_cuda(x, taps, ctx=ctx)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(5):
        gpu.fir1d_batched_f64_cuda(x, taps, ctx=ctx)
    ctx.sync()
    dt = (time.perf_counter() - t0) / 5
    outs = x.numel()
    print(f"f64 256x2^20 K={k}: {dt*1e3:.2f} ms  {outs/dt/1e9:.1f} Gsamples/s  {outs*2*k/dt/1e12:.1f} TFLOP/s f64  {outs*16/dt/1e12:.2f} TB/s")
