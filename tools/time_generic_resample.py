"""Times resample_poly for rates outside the template grid (generic polyphase kernel) next to 3/2 (template kernel).
Usage (GPU box): python tools/time_generic_resample.py"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from scir_b200 import gpu, signal
x = torch.rand((256, 1 << 20), device="cuda") * 2 - 1
for up, down in ((3, 2), (160, 147), (147, 160), (5, 4), (7, 3), (2, 5)):
    ctx = gpu.Context(0)
    w = signal.kaiser_lowpass(up, down)
    for _ in range(2):
        y = signal.resample_poly(x, up, down, w, ctx=ctx)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(5):
        y = signal.resample_poly(x, up, down, w, ctx=ctx)
    ctx.sync()
    dt = (time.perf_counter() - t0) / 5
    outs = y.numel()
    gb = (x.numel() + outs) * 4 / 1e9
    print(f"resample {up}/{down}: {w.size} taps, {dt*1e3:.2f} ms, {outs/dt/1e9:.1f} Gsamples/s out, {gb/dt/1e3:.2f} TB/s algorithmic, tile kernel launches {ctx.get_option('poly_launches')}")
