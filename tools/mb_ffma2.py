"""FP32 issue microbenchmarks on the GPU box: scalar FFMA vs packed FFMA2 (+ interleaved integer work).
Usage: python tools/mb_ffma2.py   (profiles/README.md quotes the numbers)"""
import ctypes as C, sys
sys.path.insert(0, '.')
from scir_b200 import _lib as L, gpu
lib = L.lib(); ctx = gpu.Context(0)
v = C.c_double()
lib.scir_b200_microbench_ffma(ctx.handle, 2000, C.byref(v)); print("FFMA  ", round(v.value, 2), "TFLOP/s")
for mix in (0, 1, 2, 4, 8):
    lib.scir_b200_microbench_ffma2(ctx.handle, 2000, mix, C.byref(v)); print("FFMA2 mix", mix, round(v.value, 2), "TFLOP/s")
