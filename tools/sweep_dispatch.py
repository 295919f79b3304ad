#!/usr/bin/env python3
"""Times both FIR kernel families over a sweep of (rows, n, K) and prints which one auto dispatch picks:
validates the cost model in api.cu (prefer_toeplitz).  Usage (GPU box): python tools/sweep_dispatch.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from scir_b200 import gpu


def timeit(ctx, x, taps, reps=20):
    for _ in range(3):
        gpu.fir1d_batched_f32_cuda(x, taps, ctx=ctx)
    ctx.sync()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the ctx owns its own stream: time with wall clock around a sync instead of torch events
    import time
    t0 = time.perf_counter()
    for _ in range(reps):
        gpu.fir1d_batched_f32_cuda(x, taps, ctx=ctx)
    ctx.sync()
    return (time.perf_counter() - t0) / reps * 1e6


def main():
    shapes = [(64, 1 << 14), (16, 1 << 18), (64, 1 << 18), (256, 1 << 16), (256, 1 << 18), (1024, 1 << 16), (1024, 1 << 18),
              (32, 1 << 20), (128, 1 << 20)]
    print(f"{'rows':>5} {'n':>8} {'K':>5} {'tiles':>6} {'direct us':>10} {'(tile us)':>10} {'toeplitz us':>12} {'auto us':>9}  auto picks   best")
    for rows, n in shapes:
        x = torch.rand((rows, n), device="cuda") * 2 - 1
        for k in (8, 31, 63, 127, 255, 511):
            taps = np.random.RandomState(k).randn(k).astype(np.float32)
            d, t, a, v3 = gpu.Context(0), gpu.Context(0), gpu.Context(0), gpu.Context(0)
            d.set_option("long_tap_path", 1)
            v3.set_option("variant", 3)                    # one tile per CTA (A/B arm of the direct family)
            tv = timeit(v3, x, taps)
            t.set_option("long_tap_path", 2)
            td, tt = timeit(d, x, taps), timeit(t, x, taps)
            l0 = a.get_option("toeplitz_launches")
            ta = timeit(a, x, taps)
            picked = "toeplitz" if a.get_option("toeplitz_launches") > l0 else "direct"
            best = "toeplitz" if tt < td else "direct"
            flag = "" if picked == best or abs(tt - td) / min(tt, td) < 0.15 else "   <-- model wrong"
            print(f"{rows:5d} {n:8d} {k:5d} {rows * ((n + 16383) // 16384):6d} {td:10.1f} {tv:10.1f} {tt:12.1f} {ta:9.1f}  {picked:9s} {best}{flag}")
            for c in (d, t, a, v3):
                c.close()


if __name__ == "__main__":
    main()
