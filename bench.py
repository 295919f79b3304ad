#!/usr/bin/env python3
"""bench.py -- batched FIR output Gsamples/s on N B200s (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1..c5]

A step = one pass of the hot path over one batch of synthetic input.  Default workload is
BASELINE.json configs[1] ("c2"): lfilter(b, [1], x), 1024 channels x 2^20 samples, 63 taps, f32.
N > 1: one process per GPU (torchrun), channels sharded across ranks, no data-path collective
(rows are independent); per-GPU work is fixed => weak scaling, value = all ranks' samples / max time.

One JSON line is printed by rank 0.  Keys beyond the base contract:
  roofline      the dominant kernel against the roofline that bounds it.  HBM-bound launches (config 2 on
                the tcgen05 Toeplitz kernel, config 4): algorithmic bytes / CUDA-event launch time against
                the measured HBM peak (MEASURED_PEAKS.json, else the 6.65 TB/s fallback).  Tensor-bound
                launches (configs 3, 5): executed tensor-pipe TFLOP/s against the measured dense-bf16 peak.
                `tensor`, `fp32` and `shape_roofline` (BASELINE.md's max(bytes/HBM, 2K flops/FP32)) explain it.
  cpu_baseline  the CPU restatement of the reference loop (oracle/, kind "port") timed here on the
                host cores over a bounded row sample.
  e2e           the same metric through the host-array C-ABI call (pinned host buffers, H2D+kernel+D2H
                inside the timed region).
`--impl reference` times the reference's own CPU implementation (its Rust cannot be built here: no
cargo/rustc, so the oracle port, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "batched FIR output Gsamples/s"
UNIT = "Gsamples/s"
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # 74.45, BASELINE.md section 2


def firwin_np(ntaps, cutoff, window="hamming", beta=5.0):
    """scipy.signal.firwin(ntaps, cutoff, window=...) for a lowpass, restated with numpy."""
    m = np.arange(ntaps) - (ntaps - 1) / 2.0
    h = cutoff * np.sinc(cutoff * m)
    if window == "hamming":
        h = h * np.hamming(ntaps)
    else:
        h = h * np.kaiser(ntaps, beta)
    return (h / h.sum()).astype(np.float32)


CONFIGS = {
    # name: rows per GPU, n, description, algorithmic bytes/out, flop/out
    "c1": dict(rows=64, n=1 << 14, k=31, op="fir", desc="fir_bench default 64x16384 k=31"),
    "c2": dict(rows=1024, n=1 << 20, k=63, op="lfilter", desc="lfilter a=[1], 1024x2^20, 63 taps"),
    "c3": dict(rows=256, n=1 << 22, k=4097, op="lfilter", desc="long-tap FIR 256x2^22, 4097 taps"),
    "c4": dict(rows=2048, n=1 << 20, k=96, op="resample", up=3, down=2, desc="resample_poly 3/2, 2048x2^20, Kaiser 96"),
    "c5": dict(rows=8192, n=1 << 18, k=255, op="filtfilt", desc="filtfilt FIR 255 taps, 8192x2^18, odd pad"),
    # SURVEY 8(f).1 (not a BASELINE config): DeviceArray add_scalar_auto on device-resident data, 2^30 elements
    "e1": dict(rows=1024, n=1 << 20, k=1, op="ew", desc="DeviceArray add_scalar_auto, 2^30 f32 elements (8 B/element)"),
}


def make_taps(cfg):
    if cfg["op"] == "ew":
        return np.asarray([1.5], np.float32)
    if cfg["op"] == "fir":
        return (1.0 / (np.arange(cfg["k"], dtype=np.float32) + 1.0)).astype(np.float32)      # fir_bench.rs:14-18
    if cfg["op"] == "lfilter":
        return firwin_np(cfg["k"], 0.25 if cfg["k"] < 1000 else 0.01)
    if cfg["op"] == "resample":
        return firwin_np(cfg["k"], 1.0 / 3.0, window="kaiser")
    return firwin_np(cfg["k"], 0.2)


def algorithmic(cfg, rows):
    """(out samples, bytes, flops) per step per GPU -- SURVEY.md 8(d) per-output figures."""
    n, k = cfg["n"], cfg["k"]
    if cfg["op"] == "resample":
        n_out = -(-n * cfg["up"] // cfg["down"])
        outs = rows * n_out
        return outs, rows * (n * 4 + n_out * 4), outs * 2.0 * k / cfg["up"]
    outs = rows * n
    if cfg["op"] == "filtfilt":
        return outs, outs * 8.0, outs * 2.0 * 2.0 * k
    return outs, outs * 8.0, outs * 2.0 * k


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        if not self.nv:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                     "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80, "applications_clocks_setting": 0x2,
                     "sync_boost": 0x10, "display_clock_setting": 0x100}
            for nme, bit in names.items():
                if r & bit:
                    self.reasons.add(nme)
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


def run_step(cfg, signal_mod, gpu_mod, x, taps, out):
    op = cfg["op"]
    if op == "ew":
        from scir_b200 import _lib as L
        ctx = gpu_mod.torch_context(x)
        rc = L.lib().scir_b200_add_scalar_f32(ctx.handle, C.c_void_p(x.data_ptr()), float(taps[0]), C.c_void_p(out.data_ptr()),
                                              x.numel())
        assert rc == 0, L.last_error()
        return out
    if op == "fir":
        return gpu_mod.fir1d_batched_f32_cuda(x, taps, out=out)
    if op == "lfilter":
        from scir_b200 import _lib as L
        return gpu_mod.fir1d_batched_f32_cuda(x, taps, out=out, tap_order=L.TAPS_LFILTER)
    if op == "resample":
        return signal_mod.resample_poly(x, cfg["up"], cfg["down"], taps)
    return signal_mod.filtfilt(taps, [1.0], x)


def cpu_reference_rate(cfg, taps, threads, budget_s):
    """Times the oracle port of the reference loop (gpu/lib.rs:1134-1152) over a bounded row sample.
    Returns (Gsamples/s, rows, seconds)."""
    from oracle import oracle as O
    n, k = cfg["n"], cfg["k"]
    rng = np.random.RandomState(42)
    # calibrate on one row per thread, then size the sample for ~budget_s
    rows = max(1, threads)
    x = (rng.rand(rows, n).astype(np.float32) * 2 - 1)
    kern_taps = taps if cfg["op"] == "fir" else taps[::-1].copy()    # kernel order = reversed lfilter order
    t0 = time.perf_counter()
    O.fir1d_batched_f32_mt(x, kern_taps, threads)
    dt = time.perf_counter() - t0
    reps = 1
    if cfg["op"] == "filtfilt":
        reps = 2
    want_rows = int(max(rows, min(rows * budget_s / max(dt * reps, 1e-6), 4096)))
    want_rows = max(threads, (want_rows // threads) * threads)
    if want_rows != rows:
        x = (rng.rand(want_rows, n).astype(np.float32) * 2 - 1)
    t0 = time.perf_counter()
    for _ in range(reps):
        O.fir1d_batched_f32_mt(x, kern_taps, threads)
    dt = time.perf_counter() - t0
    return want_rows * n / dt / 1e9, want_rows, dt


def reference_arm(args, cfg, rank, world):
    """--impl reference: the reference's CPU path (oracle port; Rust is not buildable here)."""
    if rank != 0:
        return
    taps = make_taps(cfg)
    threads = os.cpu_count() or 1
    if cfg["op"] == "resample":
        # the reference has no general resampler; its CPU path for this route is the same MAC loop
        # per output over the polyphase taps: time the FIR port at the per-output tap count
        eff = dict(cfg, k=cfg["k"] // cfg["up"], op="lfilter")
        taps = taps[: eff["k"]]
    else:
        eff = cfg
    vals = []
    per_step_budget = 2.0
    for _ in range(args.warmup):
        cpu_reference_rate(eff, taps, threads, 0.3)
    t_all = time.perf_counter()
    sample = None
    for _ in range(args.steps):
        g, rows, dt = cpu_reference_rate(eff, taps, threads, per_step_budget)
        vals.append(g)
        sample = f"{rows} rows x {cfg['n']} samples per step ({dt:.2f} s), extrapolated linearly over rows"
        if time.perf_counter() - t_all > 150:
            break
    v = float(np.median(vals))
    outs, _, _ = algorithmic(cfg, cfg["rows"])
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": args.warmup, "ms_per_step": outs / (v * 1e9) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "config": args.config, "note": "CPU port of gpu/lib.rs:1134-1152, row-parallel"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU (debug)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="ctx option for A/B runs, e.g. long_tap_path=2 toeplitz_terms=3")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.rows:
        cfg["rows"] = args.rows
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_arm(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    from scir_b200 import _lib as L
    from scir_b200 import dist as sdist
    from scir_b200 import gpu, signal

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    rows, n = cfg["rows"], cfg["n"]
    taps = make_taps(cfg)
    # weak scaling: every rank owns `rows` channels of a (rows*world)-channel problem
    r0, r1 = sdist.shard_rows(rows * world, world, rank)
    assert r1 - r0 == rows
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    x = torch.rand((rows, n), device=dev, generator=g) * 2 - 1
    out = torch.empty_like(x) if cfg["op"] in ("fir", "lfilter", "ew") else None
    ctx = gpu.torch_context(x)
    if args.variant:
        ctx.set_option("variant", args.variant)
    opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.opt}
    for key, val in opts.items():
        ctx.set_option(key, val)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        y = run_step(cfg, signal, gpu, x, taps, out)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.sample()
    sampler.start()
    l0 = ctx.launch_count()
    tc0 = ctx.get_option("toeplitz_launches")
    fx0 = ctx.get_option("fixup_launches")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        y = run_step(cfg, signal, gpu, x, taps, out)
    e1.record()
    torch.cuda.synchronize()
    launches = ctx.launch_count() - l0
    ms = e0.elapsed_time(e1) / args.steps
    barrier()
    sampler.sample()
    clocks = sampler.result()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    outs, abytes, aflops = algorithmic(cfg, rows)
    value = outs * world / (ms_max * 1e-3) / 1e9

    # ---- per-kernel launch time for the roofline (events on the launching stream) -------------------
    hbm_peak, peak_src, peaks = measured_peaks()
    ffma = C.c_double(0.0)
    L.lib().scir_b200_microbench_ffma(ctx.handle, 2000, C.byref(ffma))
    tc_launches = ctx.get_option("toeplitz_launches") - tc0
    tensor_path = tc_launches > 0
    # dominant-kernel launches per step (filtfilt: two).  Every FIR launch is followed by its non-finite
    # fix-up kernel (~3 us with finite data), counted in gpu_launches but not a "dominant kernel".
    n_kern = max((launches - (ctx.get_option("fixup_launches") - fx0)) / args.steps, 1)
    kern_ms = ms / n_kern
    achieved_gbs = abytes / (ms * 1e-3) / 1e9
    achieved_tf = aflops / (ms * 1e-3) / 1e12
    t_hbm = abytes / (hbm_peak * 1e9)
    t_fp32 = aflops / (FP32_NOMINAL_TFLOPS * 1e12)
    t_roof = max(t_hbm, t_fp32)
    tensor = None
    bound = "hbm"
    if tensor_path:
        # what the tcgen05 Toeplitz kernel executes: per 128 x 128 output tile, `terms` M128 N128 K16 MMAs for
        # every K step of every Toeplitz block that meets a non-zero tap (fir_toeplitz.cu: issue_tile)
        terms = opts.get("toeplitz_terms", 3)
        k = cfg["k"]
        passes = 1
        if cfg["op"] == "filtfilt":
            if tc_launches // args.steps == 1:             # padded filtfilt fused into one zero-phase pass of 2K-1 taps
                k = 2 * k - 1
            else:
                passes = 2
        pmax = (k - 1 + 127) // 128
        ksteps = sum(8 - (max(0, 128 * pb - (k - 1)) >> 4) for pb in range(pmax + 1))
        exec_flop_per_out = terms * ksteps * (2.0 * 128 * 128 * 16) / (128 * 128)
        exec_tf = outs * passes * exec_flop_per_out / (ms * 1e-3) / 1e12
        tc_peak = float(peaks.get("bf16_tflops", 2250.0))
        t_tensor = outs * passes * exec_flop_per_out / (tc_peak * 1e12)
        tensor = {"split": "fp16x%d block-scaled, fp32 accumulate in TMEM" % terms if not opts.get("toeplitz_split") else
                           "bf16x%d, fp32 accumulate in TMEM" % terms,
                  "executed_tflops": exec_tf, "peak_tflops": tc_peak,
                  "peak_source": "MEASURED_PEAKS.json bf16_tflops (cuBLAS burst)" if "bf16_tflops" in peaks else "nominal 2250",
                  "peak_sustained_tflops": peaks.get("bf16_tflops_sustained"),
                  "frac_executed": exec_tf / tc_peak,
                  "algorithmic_tflops": achieved_tf, "frac_algorithmic_x_terms": achieved_tf * terms / tc_peak,
                  "mma_per_tile": terms * ksteps}
        if t_tensor > t_hbm:
            bound = "tensor"
    if bound == "tensor":
        roofline = {"bound": "tensor", "achieved": tensor["executed_tflops"], "peak": tensor["peak_tflops"], "unit": "TFLOP/s",
                    "frac": tensor["frac_executed"], "traffic": TRAFFIC_PER_LAUNCH.get(args.config),
                    "peak_source": tensor["peak_source"], "kernel_ms": kern_ms,
                    "note": "achieved = tensor-pipe flops the kernel executes (split terms x K steps incl. block padding); "
                            "algorithmic 2*K flop/output figures are in `tensor` and `fp32`",
                    "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "frac": achieved_gbs / hbm_peak}}
    else:
        roofline = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                    "traffic": TRAFFIC_PER_LAUNCH.get(args.config), "peak_source": peak_src, "kernel_ms": kern_ms}
    roofline["tensor"] = tensor
    roofline["fp32"] = {"achieved_tflops": achieved_tf, "peak_nominal_tflops": FP32_NOMINAL_TFLOPS,
                        "frac_nominal": achieved_tf / FP32_NOMINAL_TFLOPS, "peak_ffma_microbench_tflops": ffma.value,
                        "frac_of_microbench": achieved_tf / ffma.value if ffma.value else None,
                        "note": "algorithmic 2*K flop/output against the CUDA-core FP32 peak; > 1 means the tensor path "
                                "beat the FP32 roofline" if tensor_path else "direct-form FFMA kernel"}
    roofline["shape_roofline"] = {"t_ms": t_roof * 1e3, "gsamples": outs / t_roof / 1e9, "frac": (t_roof * 1e3) / ms,
                                  "note": "BASELINE.md per-shape roofline: max(bytes/measured HBM, 2*K flops/nominal FP32)"}

    # ---- e2e: host arrays through the C ABI (pinned buffers, H2D + kernel + D2H timed) ------------------
    e2e = None
    if not args.no_e2e and cfg["op"] != "ew":
        lib = L.lib()
        if cfg["op"] == "resample":
            n_out = -(-n * cfg["up"] // cfg["down"])
        else:
            n_out = n
        in_bytes, out_bytes = rows * n * 4, rows * n_out * 4
        hx, hy = C.c_void_p(), C.c_void_p()
        rc1, rc2 = lib.scir_b200_host_alloc(in_bytes, C.byref(hx)), lib.scir_b200_host_alloc(out_bytes, C.byref(hy))
        alloc_ok = torch.tensor([1 if (rc1 == 0 and rc2 == 0) else 0], device=dev, dtype=torch.int32)
        if world > 1:                                      # every rank takes the same branch (barriers inside)
            dist.all_reduce(alloc_ok, op=dist.ReduceOp.MIN)
        if int(alloc_ok.item()) == 1:
            ax = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_float)), shape=(rows, n))
            ay = np.ctypeslib.as_array(C.cast(hy, C.POINTER(C.c_float)), shape=(rows, n_out))
            ax[:] = x.cpu().numpy()
            hctx = gpu.Context(local_rank)
            for key, val in opts.items():
                hctx.set_option(key, val)
            fpp = C.POINTER(C.c_float)
            tp = taps.ctypes.data_as(fpp)
            xp, yp = C.cast(hx, fpp), C.cast(hy, fpp)
            if cfg["op"] in ("fir", "lfilter"):
                order = L.TAPS_SCIR if cfg["op"] == "fir" else L.TAPS_LFILTER
                api = "scir_b200_fir1d_batched_f32_host"
                call = lambda: lib.scir_b200_fir1d_batched_f32_host(hctx.handle, xp, n, tp, taps.size, order, yp, n_out, rows, n)
            elif cfg["op"] == "resample":
                api = "scir_b200_resample_poly_f32_host"
                call = lambda: lib.scir_b200_resample_poly_f32_host(hctx.handle, tp, taps.size, cfg["up"], cfg["down"], xp, n,
                                                                    rows, n, yp, n_out)
            else:
                api = "scir_b200_filtfilt_fir_f32_host"
                call = lambda: lib.scir_b200_filtfilt_fir_f32_host(hctx.handle, tp, taps.size, L.PAD_ODD, -1, xp, n, yp, n_out,
                                                                   rows, n)
            e_steps = max(3, min(args.steps, 10))
            rc = 0
            for _ in range(2):
                rc |= call()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                rc |= call()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / e_steps
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            yd = y if not isinstance(y, np.ndarray) else torch.from_numpy(y)
            ok = bool(rc == 0 and np.allclose(ay[:2], yd[:2].cpu().numpy(), atol=1e-6))
            e2e = {"value": outs * world / float(tt.item()) / 1e9, "unit": UNIT, "h2d_bytes_per_step": in_bytes,
                   "d2h_bytes_per_step": out_bytes, "ms_per_step": float(tt.item()) * 1e3, "steps": e_steps,
                   "matches_device_path": ok, "api": api + " (pinned host buffers; H2D, kernels and D2H inside the timed region)"}
            if rc != 0:
                e2e["error"] = L.last_error()
        else:
            e2e = {"value": None, "unit": UNIT, "error": L.last_error() or "pinned host allocation failed on another rank"}
        if hx:
            lib.scir_b200_host_free(hx)
        if hy:
            lib.scir_b200_host_free(hy)

    # ---- CPU baseline beside it (rank 0, N=1 only) ----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and cfg["op"] != "ew":
        eff, ctaps = cfg, taps
        if cfg["op"] == "resample":
            eff = dict(cfg, k=cfg["k"] // cfg["up"], op="lfilter")
            ctaps = taps[: eff["k"]]
        v1, rows1, dt1 = cpu_reference_rate(eff, ctaps, 1, 8.0)
        threads = os.cpu_count() or 1
        vn, rowsn, dtn = cpu_reference_rate(eff, ctaps, threads, 8.0)
        cpu = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{rows1} of {rows} rows x {n} samples ({dt1:.1f} s), rows independent: extrapolated linearly",
               "all_cores": {"value": vn, "cores": threads, "sample": f"{rowsn} rows ({dtn:.1f} s)"}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "config": args.config, "rows_per_gpu": rows, "n": n, "taps": cfg["k"],
                       "global_rows": rows * world, "sharding": f"rows x{world}, no collective",
                       "l2": "inputs larger than L2 (per-step working set >> 126 MB)" if abytes > 3e8 else
                             "working set fits L2 (latency/plumbing config)",
                       "variant": args.variant, "options": opts,
                       "tensor_core_launches": int(tc_launches),
                       "arithmetic": ("f32 in/out; " + tensor["split"]) if tensor else "f32 FFMA (CUDA cores)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
# `ncu --set full` capture of this command (profiles/); None until a capture exists for the config.
TRAFFIC_PER_LAUNCH = {
    "c2": 8.576e9,      # profiles/r01_c2_fir_toeplitz_kernel_final.ncu.txt: 4.314 GB read + 4.262 GB written (algorithmic 8.590e9)
    "c3": 8.543e9,      # profiles/r01_c3_fir_toeplitz_kernel_final.ncu.txt: 4.297 + 4.246           (algorithmic 8.590e9)
    "c4": 21.450e9,     # profiles/r01_c4_upfirdn_stream_kernel_ffma2.ncu.txt: 8.612 + 12.838        (algorithmic 21.475e9)
    "c5": 17.169e9,     # profiles/r01_c5_fir_toeplitz_kernel_fused.ncu.txt (one fused pass): 8.599 + 8.570 (algorithmic 17.180e9)
}

if __name__ == "__main__":
    main()
