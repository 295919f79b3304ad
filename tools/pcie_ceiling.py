"""PCIe ceilings of the GPU box with pinned buffers: H2D alone, D2H alone, both at once (what the *_host e2e path
needs: 4.29 GB in + 4.29 GB out per config-2 step).  Usage (GPU box): python tools/pcie_ceiling.py"""
import torch, time
n = 1 << 30  # 4 GiB of float32
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, dtype=torch.float32, device="cuda")
d_b = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
gb = n * 4 / 1e9
print("H2D alone  %.1f GB/s" % (gb / t(h2d)))
print("D2H alone  %.1f GB/s" % (gb / t(d2h)))
dt = t(both)
print("both       %.1f GB/s each way (%.1f ms for 4.29 GB in + 4.29 GB out)" % (gb / dt, dt * 1e3))
# chunked (32 MB pieces) both ways, like the host pipeline
def chunked():
    c = 8 << 20
    for i in range(0, n, c):
        with torch.cuda.stream(s1): d_a[i:i + c].copy_(h_in[i:i + c], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i:i + c].copy_(d_b[i:i + c], non_blocking=True)
dt = t(chunked)
print("both, 32 MB chunks %.1f GB/s each way (%.1f ms)" % (gb / dt, dt * 1e3))
