#!/usr/bin/env python3
"""bench.py -- batched FIR output Gsamples/s on N B200s (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1..c5] [--no-sweep]

A step = one pass of the hot path over one batch of synthetic input.  The headline workload is BASELINE.json
configs[1] ("c2"): lfilter(b, [1], x), 1024 channels x 2^20 samples, 63 taps, f32.  N > 1: one process per GPU
(torchrun), channels sharded across ranks, no data-path collective (rows are independent); per-GPU work is fixed
=> weak scaling, value = all ranks' samples / max-over-ranks time.

ONE JSON line is printed by rank 0.  Keys beyond the base contract:
  roofline      the dominant kernel against the roofline that bounds it.  HBM-bound launches (config 2 on the
                tcgen05 Toeplitz kernel, config 4): algorithmic bytes / CUDA-event launch time against the measured
                HBM peak (MEASURED_PEAKS.json, else the 6.65 TB/s fallback).  Tensor-bound launches (configs 3, 5):
                executed tensor-pipe TFLOP/s against the measured dense-bf16 peak.  `tensor`, `fp32` and
                `shape_roofline` (BASELINE.md's max(bytes/HBM, 2K flops/FP32)) explain it.
  parity        GPU result of the LAST timed step against the f64-accumulating oracle on <= 6 rows of the same
                buffers: max|err|, the tolerance 1e-5 * sum|h| * max|x| (north_star), and their ratio.
  configs       the same record (ms_per_step, value, roofline, clocks with throttle reasons, parity) for EVERY
                BASELINE config c1..c5, measured in this run after the headline (skipped with --no-sweep).
  strong        N > 1 only: BASELINE configs[1] and configs[4] as FIXED problems split over the N GPUs
                (rows/N per GPU), with efficiency against the one-GPU time of the whole problem measured in this run.
  cpu_baseline  the CPU restatement of the reference loop (oracle/, kind "port") timed here on the host cores
                over a bounded row sample (N = 1 only).
  e2e           the same metric through the host-array C-ABI call (pinned host buffers, H2D + kernel + D2H inside
                the timed region), plus `pageable` (the same call on plain malloc'd arrays, what a Rust Vec or numpy
                array is), `ceiling` (what concurrent plain pinned copies of the same volume reach on this box, all
                ranks at once) and, at N > 1, `mg` (rank 0 alone drives all N GPUs through scir_b200_mg_*_host).
`--impl reference` times the reference's own CPU implementation (its Rust cannot be built here: no cargo/rustc, so
the oracle port, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
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
L2_BYTES = 126e6


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
BASELINE_CONFIGS = ("c1", "c2", "c3", "c4", "c5")
STRONG_CONFIGS = ("c2", "c5")          # BASELINE configs[1] "on 1 and 8 B200", configs[4] "scaling sweep 1/2/4/8"


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


def out_len(cfg):
    if cfg["op"] == "resample":
        return -(-cfg["n"] * cfg["up"] // cfg["down"])
    return cfg["n"]


def algorithmic(cfg, rows):
    """(out samples, bytes, flops) per step per GPU -- SURVEY.md 8(d) per-output figures."""
    n, k = cfg["n"], cfg["k"]
    if cfg["op"] == "resample":
        n_out = out_len(cfg)
        outs = rows * n_out
        return outs, rows * (n * 4 + n_out * 4), outs * 2.0 * k / cfg["up"]
    outs = rows * n
    if cfg["op"] == "filtfilt":
        return outs, outs * 8.0, outs * 2.0 * 2.0 * k
    return outs, outs * 8.0, outs * 2.0 * k


def config_dict(name, cfg, rows, world):
    """The workload description both arms print (identical for --impl ours and --impl reference)."""
    _, abytes, _ = algorithmic(cfg, rows)
    return {"workload": cfg["desc"], "config": name, "rows_per_gpu": rows, "n": cfg["n"], "taps": cfg["k"],
            "global_rows": rows * world, "sharding": f"rows x{world}, no collective",
            "l2": "inputs larger than L2 (per-step working set >> 126 MB)" if abytes > 3e8 else
                  "working set fits L2: steps rotate over input/output buffer pairs totalling > 400 MB, so every "
                  "step reads cold data"}


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock, throttle reasons and board power with NVML while a timed region runs."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
               "hw_power_brake_slowdown": 0x80, "applications_clocks_setting": 0x2, "sync_boost": 0x10,
               "display_clock_setting": 0x100}
    REJECT = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        self.power_w, self.mask = [], 0
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
            self.mask |= int(r)
            for nme, bit in self.REASONS.items():
                if r & bit:
                    self.reasons.add(nme)
            self.power_w.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.002)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        med = float(np.median(self.samples))
        reasons = sorted(self.reasons)
        out = {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples),
               "sm_mhz_min": float(min(self.samples)), "reason_mask": hex(self.mask),
               "power_w_max": max(self.power_w) if self.power_w else None}
        # a median well below max MUST carry its reason: NVML reports the power cap only while it is being
        # applied, and a 2 ms poll can miss it -- board power at the limit says the same thing
        if self.max_mhz and med < 0.93 * self.max_mhz and not reasons:
            try:
                lim = self.nv.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1000.0
            except Exception:
                lim = None
            out["power_limit_w"] = lim
            if lim and self.power_w and max(self.power_w) >= 0.9 * lim:
                out["reasons"] = ["sw_power_cap"]
                out["reasons_inferred"] = f"board power {max(self.power_w):.0f} W at the {lim:.0f} W limit"
            else:
                out["reasons"] = ["unexplained_low_clock"]
        return out

    @staticmethod
    def rejected(clocks):
        if not clocks or clocks.get("sm_mhz") is None:
            return False
        return any(r in ClockSampler.REJECT for r in clocks["reasons"]) or "unexplained_low_clock" in clocks["reasons"]


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


# ---- parity of a finished step against the oracle (checker only) ---------------------------
def parity_check(cfg, taps, x_dev, y_dev):
    """max|err| of rows of the GPU result against the f64-accumulating oracle, with the north_star tolerance
    1e-5 * sum|h| * max|x|.  <= 6 rows; config 3 (4097 taps) is checked on two windows per row (the filter is
    causal: a window plus its K-1 history reproduces those outputs exactly)."""
    from oracle import oracle as O
    rows_total, n = x_dev.shape
    rows = sorted(set([0, 1, rows_total // 2, rows_total - 2, rows_total - 1]) & set(range(rows_total)))
    if cfg["op"] == "fir":
        rows = sorted(set(rows + [rows_total // 3]))
    xs = x_dev[rows].cpu().numpy()
    ys = y_dev[rows].cpu().numpy().astype(np.float64)
    xmax = float(np.abs(xs).max())
    op, k = cfg["op"], cfg["k"]
    t64 = taps.astype(np.float64)
    rec = {"rows": len(rows), "oracle": "f64 accumulation of the f32 data (oracle/fir_oracle.c)"}
    if op in ("fir", "lfilter"):
        kt = taps if op == "fir" else taps[::-1].copy()         # oracle takes the reference's tap order
        tol = 1e-5 * float(np.abs(t64).sum()) * xmax
        if k > 1024 and n > (1 << 17):
            L = 1 << 15
            err = 0.0
            for s in (0, n // 2 - 12345, n - L):
                lo = max(0, s - (k - 1))
                want = O.fir1d_batched_f32_acc64(xs[:, lo:s + L], kt)[:, s - lo:]
                err = max(err, float(np.abs(ys[:, s:s + L] - want).max()))
            rec["windows"] = f"3 windows of {L} outputs per row (start, middle, end)"
        else:
            err = float(np.abs(ys - O.fir1d_batched_f32_acc64(xs, kt)).max())
        rec.update(max_err=err, tol=tol)
    elif op == "resample":
        up = cfg["up"]
        h = t64 * up
        tol = 1e-5 * max(float(np.abs(h[t::up]).sum()) for t in range(up)) * xmax      # the taps one output uses
        want = O.resample_poly(xs, cfg["up"], cfg["down"], taps, acc64=True)
        rec.update(max_err=float(np.abs(ys - want).max()), tol=tol,
                   tol_note="sum|h| over the taps of one output phase (h = window * up), the largest phase")
    else:   # filtfilt, odd padding: zero-phase filter hc = b (*) flip(b); the odd extension reaches 3 max|x| at the edges
        hc = np.convolve(t64, t64[::-1])
        tol = 1e-5 * float(np.abs(hc).sum()) * xmax
        want = O.filtfilt_fir(taps, xs, O.PAD_ODD, -1)
        edge = 3 * k + 2 * k
        d = np.abs(ys - want)
        rec.update(max_err=float(d[:, edge:n - edge].max()), tol=tol,
                   edges={"max_err": float(max(d[:, :edge].max(), d[:, n - edge:].max())), "tol": 3.0 * tol,
                          "note": "within padlen + 2K of either end the odd extension 2*x[0] - x[i] reaches 3 max|x|"})
        rec["edges"]["frac"] = rec["edges"]["max_err"] / rec["edges"]["tol"]
    rec["frac"] = rec["max_err"] / rec["tol"] if rec["tol"] > 0 else None
    rec["ok"] = bool(rec["max_err"] <= rec["tol"] and (("edges" not in rec) or rec["edges"]["max_err"] <= rec["edges"]["tol"]))
    return rec


# ---- one config, device-resident: CUDA events over exactly `steps` steps -----------------------------------------------
class Env:
    """Everything a measurement needs (torch, dist, library handles, rank info)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from scir_b200 import _lib as L
        from scir_b200 import dist as sdist
        from scir_b200 import gpu, signal
        self.torch, self.dist, self.L, self.sdist, self.gpu, self.signal = torch, dist, L, sdist, gpu, signal
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            import datetime
            # every collective of this script is tiny and follows at most a few seconds of rank-local work: a rank that
            # waits for minutes is a bug, and a fast abort beats a 10-minute watchdog hang
            dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(seconds=180))
        self.opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.opt}
        self.variant = args.variant
        self.hbm_peak, self.peak_src, self.peaks = measured_peaks()
        self.ffma_tflops = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def any_rank(self, flag) -> bool:
        """True on EVERY rank if `flag` is true on any rank.  Every decision that changes how many collectives a rank
        goes on to call (re-measuring, skipping a leg after a failed allocation) must go through this: a rank-local
        branch around a barrier deadlocks the job (it did: one GPU's clock sample re-measured alone at N = 8)."""
        return self.max_over_ranks(1.0 if flag else 0.0) > 0.0

    def guarded(self, fn):
        """(result, None) or (None, error text) -- the SAME branch on every rank.  A deterministic failure (a bug, an
        out-of-memory every rank hits) is caught on all ranks at the same point and reported in the JSON line instead of
        losing the whole run; a failure on one rank alone, in the middle of another rank's collectives, still ends in the
        process group's 180 s timeout -- an abort, not a hang."""
        out, err = None, None
        try:
            out = fn()
        except Exception as e:                              # noqa: BLE001 -- reported, not swallowed
            err = repr(e)
        if self.any_rank(err is not None):
            return None, err or "failed on another rank"
        return out, None

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(self, v):
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t.item())

    def all_ranks(self, v):
        """[v on rank 0, v on rank 1, ...] (device all-gather of one float64)"""
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        if self.world == 1:
            return [float(t.item())]
        out = self.torch.empty(self.world, device=self.dev, dtype=self.torch.float64)
        self.dist.all_gather_into_tensor(out, t)
        return [float(x) for x in out.tolist()]

    def sum_over_ranks(self, v):
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def executed_mma_flops_per_out(k, terms):
    """What the tcgen05 Toeplitz kernel executes per output: per 128 x 128 tile, `terms` M128 N128 K16 MMAs for every
    K step of every Toeplitz block that meets a non-zero tap (fir_toeplitz.cu: issue_tile)."""
    pmax = (k - 1 + 127) // 128
    ksteps = sum(8 - (max(0, 128 * pb - (k - 1)) >> 4) for pb in range(pmax + 1))
    return terms * ksteps * (2.0 * 128 * 128 * 16) / (128 * 128), terms * ksteps


def measure_config(env, name, cfg, rows, steps, warmup, want_parity=True, keep=False):
    """Times `steps` steps of one config with inputs resident in HBM.  Returns the record (and, with keep=True, the
    tensors of the last step for the e2e cross-check)."""
    torch = env.torch
    n = cfg["n"]
    taps = make_taps(cfg)
    outs, abytes, aflops = algorithmic(cfg, rows)
    nbuf = 1 if abytes > 3e8 else int(math.ceil(4e8 / max(abytes, 1)))
    g = torch.Generator(device=env.dev).manual_seed(42 + env.rank)
    xs = [torch.rand((rows, n), device=env.dev, generator=g) * 2 - 1 for _ in range(nbuf)]
    outs_buf = [torch.empty_like(x) if cfg["op"] in ("fir", "lfilter", "ew") else None for x in xs]
    ctx = env.gpu.torch_context(xs[0])
    if env.variant:
        ctx.set_option("variant", env.variant)
    for key, val in env.opts.items():
        ctx.set_option(key, val)

    step_stats = {}

    def timed(count):
        sampler = ClockSampler(env.local_rank)
        sampler.sample()
        sampler.start()
        l0, tc0, fx0 = ctx.launch_count(), ctx.get_option("toeplitz_launches"), ctx.get_option("fixup_launches")
        os0 = ctx.get_option("os_launches")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(min(count, 64))]      # per-step spread (diagnostic)
        e0.record()
        y = None
        for i in range(count):
            y = run_step(cfg, env.signal, env.gpu, xs[i % nbuf], taps, outs_buf[i % nbuf])
            if i < len(marks):
                marks[i].record()
        e1.record()
        torch.cuda.synchronize()
        sampler.sample()
        clocks = sampler.result()
        prev, per = e0, []
        for mk in marks:
            per.append(prev.elapsed_time(mk))
            prev = mk
        step_stats["median"], step_stats["max"], step_stats["min"] = float(np.median(per)), float(max(per)), float(min(per))
        return (e0.elapsed_time(e1) / count, clocks, ctx.launch_count() - l0, ctx.get_option("toeplitz_launches") - tc0,
                ctx.get_option("fixup_launches") - fx0, y, (count - 1) % nbuf, ctx.get_option("os_launches") - os0)

    for i in range(warmup):
        run_step(cfg, env.signal, env.gpu, xs[i % nbuf], taps, outs_buf[i % nbuf])
    env.barrier()
    ms, clocks, launches, tc_launches, fx_launches, y, last, os_launches = timed(steps)
    remeasured = False
    if env.any_rank(ClockSampler.rejected(clocks)):         # the contract: rejected once, measured again -- by ALL ranks together
        env.barrier()
        first = clocks
        ms, clocks, launches, tc_launches, fx_launches, y, last, os_launches = timed(steps)
        clocks["first_attempt"] = {k: first.get(k) for k in ("sm_mhz", "reasons")}
        remeasured = True
    env.barrier()
    ms_ranks = env.all_ranks(ms)
    ms_max = max(ms_ranks)
    value = outs * env.world / (ms_max * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (events on the launching stream) ------------------------------------
    tensor_path = tc_launches > 0
    n_kern = max((launches - fx_launches) / steps, 1)       # dominant-kernel launches per step (two-pass filtfilt: 2)
    kern_ms = ms / n_kern
    achieved_gbs = abytes / (ms * 1e-3) / 1e9
    achieved_tf = aflops / (ms * 1e-3) / 1e12
    t_hbm = abytes / (env.hbm_peak * 1e9)
    t_fp32 = aflops / (FP32_NOMINAL_TFLOPS * 1e12)
    t_roof = max(t_hbm, t_fp32)
    tensor, bound = None, "hbm"
    if tensor_path:
        terms = env.opts.get("toeplitz_terms", 3)
        k, passes = cfg["k"], 1
        if cfg["op"] == "filtfilt":
            if tc_launches // steps == 1:                   # padded filtfilt fused into one zero-phase pass of 2K-1 taps
                k = 2 * k - 1
            else:
                passes = 2
        per_out, mma_per_tile = executed_mma_flops_per_out(k, terms)
        hh_blocks = 0
        try:                                                # what the last launch really issued (small-tap blocks run hi x hi only)
            mpt = ctx.get_option("toeplitz_mma_per_tile")
            hh_blocks = ctx.get_option("toeplitz_hh_blocks")
            if mpt > 0:
                per_out, mma_per_tile = per_out * mpt / mma_per_tile, mpt
        except Exception:
            pass
        exec_tf = outs * passes * per_out / (ms * 1e-3) / 1e12
        tc_peak = float(env.peaks.get("bf16_tflops", 2250.0))
        t_tensor = outs * passes * per_out / (tc_peak * 1e12)
        tensor = {"split": ("fp16x%d block-scaled, fp32 accumulate in TMEM" % terms) if not env.opts.get("toeplitz_split") else
                           ("bf16x%d, fp32 accumulate in TMEM" % terms),
                  "executed_tflops": exec_tf, "peak_tflops": tc_peak,
                  "peak_source": "MEASURED_PEAKS.json bf16_tflops (cuBLAS burst)" if "bf16_tflops" in env.peaks else "nominal 2250",
                  "peak_sustained_tflops": env.peaks.get("bf16_tflops_sustained"),
                  "frac_executed": exec_tf / tc_peak,
                  "algorithmic_tflops": achieved_tf, "frac_algorithmic_x_terms": achieved_tf * terms / tc_peak,
                  "mma_per_tile": mma_per_tile, "blocks_hi_x_hi_only": hh_blocks}
        if t_tensor > t_hbm:
            bound = "tensor"
    traffic = TRAFFIC_PER_LAUNCH.get(name) if rows == CONFIGS[name]["rows"] else None
    fft = None
    if os_launches > 0:
        # overlap-save FFT kernel (fir_os.cu): block N = L + K - 1, two blocks per complex transform
        k_eff = cfg["k"]
        if cfg["op"] == "filtfilt" and os_launches // steps == 1:
            k_eff = 2 * k_eff - 1
        nfft = 4096 if k_eff <= 640 else 16384
        lblk = nfft - (k_eff - 1)
        passes = 6 if nfft == 4096 else 8                      # shared-memory round trips per transform pair (fwd + inv)
        fft = {"block": nfft, "valid_outputs_per_block": lblk, "taps_per_pass": k_eff,
               "smem_bytes_per_output": passes * 0.5 * 2 * nfft * 8 / (2.0 * lblk),
               "hbm_bytes_per_output_executed": (2.0 * nfft * 4 + 2.0 * lblk * 4) / (2.0 * lblk),
               "note": "FP32 shared-memory FFT, two real blocks per complex transform; executes O(log N) flop per output "
                       "instead of 2K: the binding resources are shared-memory bandwidth and FP32 issue, not HBM or the tensor pipe"}
        traffic = TRAFFIC_PER_LAUNCH.get(name + "_os") if rows == CONFIGS[name]["rows"] else None
    if bound == "tensor":
        roofline = {"bound": "tensor", "achieved": tensor["executed_tflops"], "peak": tensor["peak_tflops"], "unit": "TFLOP/s",
                    "frac": tensor["frac_executed"], "traffic": traffic,
                    "peak_source": tensor["peak_source"], "kernel_ms": kern_ms,
                    "note": "achieved = tensor-pipe flops the kernel executes (split terms x K steps incl. block padding); "
                            "algorithmic 2*K flop/output figures are in `tensor` and `fp32`",
                    "hbm": {"achieved": achieved_gbs, "peak": env.hbm_peak, "frac": achieved_gbs / env.hbm_peak}}
    else:
        roofline = {"bound": "hbm", "achieved": achieved_gbs, "peak": env.hbm_peak, "unit": "GB/s",
                    "frac": achieved_gbs / env.hbm_peak, "traffic": traffic, "peak_source": env.peak_src, "kernel_ms": kern_ms}
    roofline["tensor"] = tensor
    roofline["fft"] = fft
    roofline["fp32"] = {"achieved_tflops": achieved_tf, "peak_nominal_tflops": FP32_NOMINAL_TFLOPS,
                        "frac_nominal": achieved_tf / FP32_NOMINAL_TFLOPS,
                        "peak_ffma_microbench_tflops": env.ffma_tflops,
                        "frac_of_microbench": achieved_tf / env.ffma_tflops if env.ffma_tflops else None,
                        "note": "algorithmic 2*K flop/output against the CUDA-core FP32 peak; > 1 means the tensor path "
                                "beat the FP32 roofline" if tensor_path else "direct-form FFMA kernel"}
    roofline["shape_roofline"] = {"t_ms": t_roof * 1e3, "gsamples": outs / t_roof / 1e9, "frac": (t_roof * 1e3) / ms,
                                  "note": "BASELINE.md per-shape roofline: max(bytes/measured HBM, 2*K flops/nominal FP32)"}
    rec = {"workload": cfg["desc"], "rows_per_gpu": rows, "steps": steps, "ms_per_step": ms_max, "ms_by_rank": ms_ranks,
           "step_ms_rank0": dict(step_stats),
           "value": value, "unit": UNIT,
           "roofline": roofline, "clocks": clocks, "clock_remeasured": remeasured,
           "gpu_launches": int(launches), "tensor_core_launches": int(tc_launches), "fft_launches": int(os_launches),
           "arithmetic": ("f32 in/out; " + tensor["split"]) if tensor else
                         ("f32 overlap-save FFT (CUDA cores)" if os_launches > 0 else "f32 FFMA (CUDA cores)"),
           "l2": config_dict(name, cfg, rows, env.world)["l2"]}
    if want_parity and cfg["op"] != "ew":
        yd = y if not isinstance(y, np.ndarray) else torch.from_numpy(y)
        try:
            rec["parity"] = parity_check(cfg, taps, xs[last], yd)
        except Exception as e:                              # the checker must never take the measurement down
            rec["parity"] = {"error": repr(e)}
    kept = (xs[last], y, taps, ctx) if keep else None
    if not keep:
        del xs, outs_buf, y
        torch.cuda.empty_cache()
    return rec, kept


# ---- e2e: host arrays through the C ABI -------------------------------------------------------------------------------
def host_call(env, cfg, taps, hctx_handle, xp, yp, rows, mg=None):
    """The reference-facing call on host arrays for this config (lib.rs:515-531 and the scir-signal routes)."""
    lib, L = env.L.lib(), env.L
    n, n_out = cfg["n"], out_len(cfg)
    fpp = C.POINTER(C.c_float)
    tp = taps.ctypes.data_as(fpp)
    if cfg["op"] in ("fir", "lfilter"):
        order = L.TAPS_SCIR if cfg["op"] == "fir" else L.TAPS_LFILTER
        if mg is not None:
            return "scir_b200_mg_fir1d_batched_f32_host", lambda: lib.scir_b200_mg_fir1d_batched_f32_host(
                mg, xp, n, tp, taps.size, order, yp, n_out, rows, n)
        return "scir_b200_fir1d_batched_f32_host", lambda: lib.scir_b200_fir1d_batched_f32_host(
            hctx_handle, xp, n, tp, taps.size, order, yp, n_out, rows, n)
    if cfg["op"] == "resample":
        if mg is not None:
            return "scir_b200_mg_resample_poly_f32_host", lambda: lib.scir_b200_mg_resample_poly_f32_host(
                mg, tp, taps.size, cfg["up"], cfg["down"], xp, n, rows, n, yp, n_out)
        return "scir_b200_resample_poly_f32_host", lambda: lib.scir_b200_resample_poly_f32_host(
            hctx_handle, tp, taps.size, cfg["up"], cfg["down"], xp, n, rows, n, yp, n_out)
    if mg is not None:
        return "scir_b200_mg_filtfilt_fir_f32_host", lambda: lib.scir_b200_mg_filtfilt_fir_f32_host(
            mg, tp, taps.size, L.PAD_ODD, -1, xp, n, yp, n_out, rows, n)
    return "scir_b200_filtfilt_fir_f32_host", lambda: lib.scir_b200_filtfilt_fir_f32_host(
        hctx_handle, tp, taps.size, L.PAD_ODD, -1, xp, n, yp, n_out, rows, n)


def measure_e2e(env, cfg, rows, steps, kept):
    torch, lib, L = env.torch, env.L.lib(), env.L
    x_dev, y_dev, taps, _ = kept
    n, n_out = cfg["n"], out_len(cfg)
    outs, _, _ = algorithmic(cfg, rows)
    in_bytes, out_bytes = rows * n * 4, rows * n_out * 4
    hx, hy = C.c_void_p(), C.c_void_p()
    rc1, rc2 = lib.scir_b200_host_alloc(in_bytes, C.byref(hx)), lib.scir_b200_host_alloc(out_bytes, C.byref(hy))
    ok_all = env.min_over_ranks(1.0 if (rc1 == 0 and rc2 == 0) else 0.0)        # every rank takes the same branch
    if ok_all < 1.0:
        e2e = {"value": None, "unit": UNIT, "error": L.last_error() or "pinned host allocation failed on another rank"}
        if hx:
            lib.scir_b200_host_free(hx)
        if hy:
            lib.scir_b200_host_free(hy)
        return e2e
    fpp = C.POINTER(C.c_float)
    ax = np.ctypeslib.as_array(C.cast(hx, fpp), shape=(rows, n))
    ay = np.ctypeslib.as_array(C.cast(hy, fpp), shape=(rows, n_out))
    ax[:] = x_dev.cpu().numpy()
    hctx = env.gpu.Context(env.local_rank)
    for key, val in env.opts.items():
        hctx.set_option(key, val)
    yd = y_dev if not isinstance(y_dev, np.ndarray) else torch.from_numpy(y_dev)
    y_head = yd[:2].cpu().numpy()
    e_steps = max(3, min(steps, 10))

    def timed(call, count, warm=2):
        rc = 0
        for _ in range(warm):
            rc |= call()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(count):
            rc |= call()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / count
        return env.max_over_ranks(dt), rc

    api, call = host_call(env, cfg, taps, hctx.handle, C.cast(hx, fpp), C.cast(hy, fpp), rows)
    dt, rc = timed(call, e_steps)
    ok = bool(rc == 0 and np.allclose(ay[:2], y_head, atol=1e-6))
    e2e = {"value": outs * env.world / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": in_bytes,
           "d2h_bytes_per_step": out_bytes, "ms_per_step": dt * 1e3, "steps": e_steps,
           "matches_device_path": ok, "gbs_each_way_per_gpu": in_bytes / dt / 1e9,
           "api": api + " (pinned host buffers; H2D, kernels and D2H inside the timed region)"}
    if rc != 0:
        e2e["error"] = L.last_error()

    # ---- the same call on PAGEABLE arrays: what a Rust Vec / ndarray::Array2 / numpy array is (lib.rs:1042-1044) ----
    px = py = None
    try:
        px = np.empty((rows, n), np.float32)
        py = np.empty((rows, n_out), np.float32)
        px[:] = ax
        py[:] = 0
    except MemoryError:
        px = py = None
    if env.any_rank(px is None):
        e2e["pageable"] = {"error": "pageable test arrays could not be allocated on every rank"}
    else:
        pin = C.c_int(-1)
        lib.scir_b200_host_is_pinned(px.ctypes.data_as(C.c_void_p), px.nbytes, C.byref(pin))
        _, pcall = host_call(env, cfg, taps, hctx.handle, px.ctypes.data_as(fpp), py.ctypes.data_as(fpp), rows)
        p_steps = max(2, min(steps, 5))
        s0 = hctx.get_option("host_staged_calls")
        dtp, rcp = timed(pcall, p_steps, warm=1)
        # (the pinned and the pageable call may use different row blocks, hence different kernels: compare to tolerance)
        okp = bool(rcp == 0 and np.allclose(py[:2], y_head, atol=1e-5) and np.allclose(py[rows // 2:], ay[rows // 2:], atol=1e-5))
        e2e["pageable"] = {"value": outs * env.world / dtp / 1e9, "unit": UNIT, "ms_per_step": dtp * 1e3, "steps": p_steps,
                           "vs_pinned": dt / dtp, "matches_device_path": okp, "input_is_pinned": int(pin.value),
                           "route": "ctx pinned ring + copy threads" if hctx.get_option("host_staged_calls") > s0 else
                                    ("cudaHostRegister per call" if hctx.get_option("host_registered_calls") > 0 else
                                     "cudaMemcpyAsync on pageable memory"),
                           "copy_threads": hctx.get_option("host_copy_threads") or "auto"}
        if rcp != 0:
            e2e["pageable"]["error"] = L.last_error()
    del px, py

    # ---- ceiling: what plain pinned copies reach on THIS box, every rank at the same time ---------------------------------
    # 4 x 1 GiB each way per rank on two streams, all ranks released together by a barrier (an unsynchronised ceiling --
    # ranks in different phases -- overstates what the host delivers when all GPUs pull at once)
    nbytes, reps = 1 << 30, 4
    hp_in = hp_out = d_in = d_out = None
    try:
        hp_in = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
        hp_out = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
        hp_in.fill_(1.0)
        d_in = torch.empty(nbytes // 4, dtype=torch.float32, device=env.dev)
        d_out = torch.ones(nbytes // 4, dtype=torch.float32, device=env.dev)
    except Exception as e:                                  # rank-local failure: agreed on below, before any collective
        hp_in = None
        e2e["ceiling"] = {"error": repr(e)}
    if env.any_rank(hp_in is None):
        e2e.setdefault("ceiling", {"error": "pinned buffers for the ceiling test could not be allocated on every rank"})
    else:
        s1, s2 = torch.cuda.Stream(device=env.dev), torch.cuda.Stream(device=env.dev)

        def copies(h2d, d2h):
            env.barrier()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            if h2d:
                with torch.cuda.stream(s1):
                    ev[0].record()
                    for _ in range(reps):
                        d_in.copy_(hp_in, non_blocking=True)
                    ev[1].record()
            if d2h:
                with torch.cuda.stream(s2):
                    ev[2].record()
                    for _ in range(reps):
                        hp_out.copy_(d_out, non_blocking=True)
                    ev[3].record()
            torch.cuda.synchronize()
            t_h = ev[0].elapsed_time(ev[1]) * 1e-3 if h2d else None
            t_d = ev[2].elapsed_time(ev[3]) * 1e-3 if d2h else None
            return (reps * nbytes / t_h / 1e9 if h2d else None), (reps * nbytes / t_d / 1e9 if d2h else None)

        copies(True, True)                                                   # warm-up
        h_alone, _ = copies(True, False)
        _, d_alone = copies(False, True)
        h_dup, d_dup = copies(True, True)
        each = min(h_dup, d_dup)
        e2e["ceiling"] = {"h2d_alone_gbs": h_alone, "d2h_alone_gbs": d_alone, "duplex_h2d_gbs": h_dup, "duplex_d2h_gbs": d_dup,
                          "duplex_each_way_gbs_min_over_ranks": env.min_over_ranks(each),
                          "duplex_each_way_gbs_sum_over_ranks": env.sum_over_ranks(each),
                          "h2d_alone_gbs_sum_over_ranks": env.sum_over_ranks(h_alone),
                          "d2h_alone_gbs_sum_over_ranks": env.sum_over_ranks(d_alone),
                          "note": "4 x 1 GiB pinned copies each way per rank on two streams, every rank released by the same barrier "
                                  "(rank 0's figures; min / sum over ranks beside them)"}
        agg = e2e["ceiling"]["duplex_each_way_gbs_sum_over_ranks"]
        e2e["ceiling"]["frac_of_ceiling"] = (in_bytes * env.world / dt / 1e9) / agg if agg else None
        e2e["ceiling_gbs"] = agg
    del hp_in, hp_out, d_in, d_out

    # ---- N > 1: the in-process front end north_star (d) describes -- rank 0 drives all N GPUs, the others wait ------
    if env.world > 1:
        env.barrier()
        mg_rec = None
        if env.rank == 0:
            try:
                devs = (C.c_int * env.world)(*range(env.world))
                mg = C.c_void_p()
                rcm = lib.scir_b200_mg_create(devs, env.world, C.byref(mg))
                if rcm == 0:
                    _, mcall = host_call(env, cfg, taps, None, C.cast(hx, fpp), C.cast(hy, fpp), rows, mg=mg)
                    ay[:] = 0
                    for _ in range(2):
                        rcm |= mcall()
                    t0 = time.perf_counter()
                    for _ in range(e_steps):
                        rcm |= mcall()
                    dtm = (time.perf_counter() - t0) / e_steps
                    okm = bool(rcm == 0 and np.allclose(ay[:2], y_head, atol=1e-6))
                    mg_rec = {"value": outs / dtm / 1e9, "unit": UNIT, "ms_per_step": dtm * 1e3, "steps": e_steps,
                              "devices": env.world, "rows_total": rows, "rows_per_gpu": rows // env.world,
                              "matches_device_path": okm,
                              "api": "scir_b200_mg_*_host: ONE host call, rows of the FIXED problem (this config's rows) sharded "
                                     "over the N devices, one stream + host thread per device (pinned arrays)"}
                    lib.scir_b200_mg_destroy(mg)
                if rcm != 0:
                    mg_rec = {"error": L.last_error()}
            except Exception as e:
                mg_rec = {"error": repr(e)}
        env.barrier()
        e2e["mg"] = mg_rec
    lib.scir_b200_host_free(hx)
    lib.scir_b200_host_free(hy)
    hctx.close()
    return e2e


# ---- CPU reference arm ---------------------------------------------------------------------------------------------
def cpu_cfg(cfg, taps):
    """The reference's CPU path for a config, as the oracle port's FIR loop: (effective cfg, taps, passes)."""
    if cfg["op"] == "resample":
        # the reference has no general resampler; its CPU path for this route is the same MAC loop per output over
        # the polyphase taps: time the FIR port at the per-output tap count
        eff = dict(cfg, k=cfg["k"] // cfg["up"], op="lfilter")
        return eff, taps[: eff["k"]], 1
    return cfg, taps, (2 if cfg["op"] == "filtfilt" else 1)


_CPU_X = {}


def cpu_time_rows(cfg, taps, threads, rows, passes):
    """Seconds for `rows` rows of the oracle port of gpu/lib.rs:1134-1152 on `threads` threads."""
    from oracle import oracle as O
    key = (rows, cfg["n"])
    if key not in _CPU_X:
        _CPU_X.clear()
        _CPU_X[key] = (np.random.RandomState(42).rand(rows, cfg["n"]).astype(np.float32) * 2 - 1)
    x = _CPU_X[key]
    kern_taps = taps if cfg["op"] == "fir" else taps[::-1].copy()    # kernel order = reversed lfilter order
    t0 = time.perf_counter()
    for _ in range(passes):
        O.fir1d_batched_f32_mt(x, kern_taps, threads)
    return time.perf_counter() - t0


def cpu_reference_rate(cfg, taps, threads, budget_s):
    """Times the oracle port over a bounded row sample sized for ~budget_s.  Returns (Gsamples/s, rows, seconds)."""
    eff, ctaps, passes = cpu_cfg(cfg, taps)
    rows = max(1, threads)
    dt = cpu_time_rows(eff, ctaps, threads, rows, passes)
    want = int(max(rows, min(rows * budget_s / max(dt, 1e-6), 4096)))
    want = max(threads, (want // threads) * threads)
    if want != rows:
        rows = want
        dt = cpu_time_rows(eff, ctaps, threads, rows, passes)
    return rows * cfg["n"] / dt / 1e9, rows, dt


def reference_arm(args, name, cfg, rank, world):
    """--impl reference: the reference's CPU path (oracle port; Rust is not buildable here).  Exactly --steps steps,
    each one bounded sample of the workload (the row count is fixed after a calibration pass)."""
    if rank != 0:
        return
    taps = make_taps(cfg)
    threads = os.cpu_count() or 1
    eff, ctaps, passes = cpu_cfg(cfg, taps)
    per_step = min(2.0, 90.0 / max(args.steps + args.warmup, 1))
    dt0 = cpu_time_rows(eff, ctaps, threads, threads, passes)             # calibration (untimed)
    rows = int(max(threads, min(threads * per_step / max(dt0, 1e-6), 4096)))
    rows = max(threads, (rows // threads) * threads)
    for _ in range(args.warmup):
        cpu_time_rows(eff, ctaps, threads, rows, passes)
    ratio = 1.0       # outputs timed = rows * n at the config's per-output tap count (see cpu_cfg)
    vals, dts = [], []
    for _ in range(args.steps):
        dt = cpu_time_rows(eff, ctaps, threads, rows, passes)
        dts.append(dt)
        vals.append(rows * cfg["n"] * ratio / dt / 1e9)
    v = float(np.median(vals))
    # what the reference actually ships is single-threaded (no rayon anywhere): state that figure too
    rows1 = max(1, int(rows / threads / 2) or 1)
    dt1 = cpu_time_rows(eff, ctaps, 1, rows1, passes)
    v1 = rows1 * cfg["n"] * ratio / dt1 / 1e9
    outs, _, _ = algorithmic(cfg, cfg["rows"])
    sample = (f"{rows} of {cfg['rows']} rows x {cfg['n']} samples per step ({float(np.median(dts)):.2f} s), rows independent: "
              "extrapolated linearly")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": outs / (v * 1e9) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(name, cfg, cfg["rows"], world),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "single_thread": {"value": v1, "cores": 1, "sample": f"{rows1} rows ({dt1:.2f} s)",
                                           "note": "the reference loop as shipped is single-threaded (gpu/lib.rs:1134-1152)"},
                         "note": "CPU port of gpu/lib.rs:1134-1152 (oracle/fir_oracle.c), row-parallel over all host threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
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
    ap.add_argument("--no-sweep", action="store_true", help="only the headline config (skip `configs` and `strong`)")
    ap.add_argument("--strong-div", type=int, default=0,
                    help="debug: run the strong-split phase with rows/D per GPU even at N=1 (what one rank of a D-GPU run does)")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="ctx option for A/B runs, e.g. long_tap_path=2 toeplitz_terms=3")
    args = ap.parse_args()
    name = args.config
    cfg = dict(CONFIGS[name])
    if args.rows:
        cfg["rows"] = args.rows
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, name, cfg, rank, world)
        return

    env = Env(args)
    torch = env.torch
    rows = cfg["rows"]
    r0, r1 = env.sdist.shard_rows(rows * world, world, rank)     # weak scaling: every rank owns `rows` channels
    assert r1 - r0 == rows

    # the FP32 denominator measured on this box (for roofline.fp32)
    probe = torch.empty(4, device=env.dev)
    ffma = C.c_double(0.0)
    env.L.lib().scir_b200_microbench_ffma(env.gpu.torch_context(probe).handle, 2000, C.byref(ffma))
    env.ffma_tflops = ffma.value

    want_e2e = not args.no_e2e and cfg["op"] != "ew"
    head, kept = measure_config(env, name, cfg, rows, args.steps, args.warmup, keep=want_e2e)
    # (the e2e leg allocates and frees 16 GiB of pinned / pageable host memory and, at N > 1, opens contexts on every GPU
    # from rank 0: it runs LAST, after every device-resident measurement -- measured: run first, it left a one-off ~110 ms
    # stall in a later timed region)

    # ---- every BASELINE config, same rules (device-resident) --------------------------------------------------
    sweep, strong = None, None
    if not args.no_sweep and not args.rows and name in BASELINE_CONFIGS:
        sweep = {}
        for cn in BASELINE_CONFIGS:
            if cn == name:
                sweep[cn] = {k: v for k, v in head.items()}
                continue
            c = CONFIGS[cn]
            outs_c, abytes_c, _ = algorithmic(c, c["rows"])
            steps_c = args.steps if abytes_c > 3e8 else max(args.steps, 200)     # short configs: enough steps for the clock sampler
            # (failures are agreed across ranks: Env.guarded)
            res, err = env.guarded(lambda: measure_config(env, cn, c, c["rows"], steps_c, args.warmup)[0])
            sweep[cn] = res if err is None else {"error": err}
        # ---- the BASELINE shapes as FIXED problems split over N GPUs (strong scaling) ------------------------------
        div = world if world > 1 else args.strong_div
        if div > 1:
            strong = {}
            for cn in STRONG_CONFIGS:
                c = CONFIGS[cn]
                if c["rows"] % div:
                    continue
                rec, err = env.guarded(lambda: measure_config(env, cn, c, c["rows"] // div, args.steps, args.warmup)[0])
                if err is not None or "ms_per_step" not in sweep.get(cn, {}):
                    strong[cn] = {"error": err or "the one-GPU measurement of this config failed"}
                    continue
                t1 = sweep[cn]["ms_per_step"]                                   # the whole problem on ONE GPU, this run (max over ranks)
                outs_total, _, _ = algorithmic(c, c["rows"])
                strong[cn] = {"workload": c["desc"] + f" split over {div} GPUs", "rows_per_gpu": c["rows"] // div,
                              "ms": rec["ms_per_step"], "value": outs_total / (rec["ms_per_step"] * 1e-3) / 1e9, "unit": UNIT,
                              "one_gpu_ms": t1, "speedup_vs_n1": t1 / rec["ms_per_step"],
                              "efficiency_vs_n1": t1 / rec["ms_per_step"] / div,
                              "roofline": {k: rec["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "kernel_ms")},
                              "ms_by_rank": rec["ms_by_rank"], "step_ms_rank0": rec["step_ms_rank0"],
                              "clocks": rec["clocks"], "parity": rec.get("parity"), "gpu_launches": rec["gpu_launches"]}

    e2e = None
    if want_e2e:
        e2e, err = env.guarded(lambda: measure_e2e(env, cfg, rows, args.steps, kept))
        if err is not None:
            e2e = {"value": None, "unit": UNIT, "error": err}
    del kept
    torch.cuda.empty_cache()

    # ---- CPU baseline beside it (rank 0, N=1 only) ----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and cfg["op"] != "ew":
        taps = make_taps(cfg)
        v1, rows1, dt1 = cpu_reference_rate(cfg, taps, 1, 8.0)
        threads = os.cpu_count() or 1
        vn, rowsn, dtn = cpu_reference_rate(cfg, taps, threads, 8.0)
        cpu = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{rows1} of {rows} rows x {cfg['n']} samples ({dt1:.1f} s), rows independent: extrapolated linearly",
               "all_cores": {"value": vn, "cores": threads, "sample": f"{rowsn} rows ({dtn:.1f} s)"}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(name, cfg, rows, world),
            "gpu": {"variant": args.variant, "options": env.opts, "tensor_core_launches": head["tensor_core_launches"],
                    "fft_launches": head["fft_launches"], "arithmetic": head["arithmetic"]},
            "roofline": head["roofline"], "parity": head.get("parity"), "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": head["gpu_launches"], "clocks": head["clocks"],
            "configs": sweep, "strong": strong,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        env.dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
# `ncu --set full` capture of this command (profiles/); None until a capture exists for the config.
TRAFFIC_PER_LAUNCH = {
    "c2": 8.576e9,      # profiles/r01_c2_fir_toeplitz_kernel_final.ncu.txt: 4.314 GB read + 4.262 GB written (algorithmic 8.590e9)
    "c3": 8.543e9,      # profiles/r01_c3_fir_toeplitz_kernel_final.ncu.txt: 4.297 + 4.246           (algorithmic 8.590e9)
    "c3_os": 8.547e9,   # profiles/r02_c3_fir_os_kernel.ncu.txt (overlap-save FFT, the default): 4.297 + 4.250
    "c5_os": 17.140e9,  # profiles/r02_c5_fir_os_kernel_n4096.ncu.txt (A/B arm): 8.592 + 8.548
    "c4": 21.452e9,     # profiles/r02_c4_upfirdn_ws_kernel.ncu.txt: 8.618 + 12.834                  (algorithmic 21.475e9)
    "c5": 17.169e9,     # profiles/r01_c5_fir_toeplitz_kernel_fused.ncu.txt (one fused pass): 8.599 + 8.570 (algorithmic 17.180e9)
}

if __name__ == "__main__":
    main()
