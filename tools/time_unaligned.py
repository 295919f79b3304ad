import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from scir_b200 import gpu
big = torch.rand((256, (1 << 18) + 8), device="cuda") * 2 - 1
for name, view in (("aligned", big[:, :1 << 18]), ("off by one sample", big[:, 1:(1 << 18) + 1])):
    for k in (63, 255):
        taps = np.random.RandomState(k).randn(k).astype(np.float32)
        res = []
        for mode in (1, 2, 0):
            ctx = gpu.Context(0); ctx.set_option("long_tap_path", mode)
            for _ in range(2): gpu.fir1d_batched_f32_cuda(view, taps, ctx=ctx)
            ctx.sync(); t0 = time.perf_counter()
            for _ in range(5): gpu.fir1d_batched_f32_cuda(view, taps, ctx=ctx)
            ctx.sync(); res.append((time.perf_counter() - t0) / 5 * 1e6)
        print(f"{name:18s} K={k:3d}: direct {res[0]:8.1f} us  toeplitz {res[1]:8.1f} us  auto {res[2]:8.1f} us")
