#!/usr/bin/env python3
"""The OPTIONAL collective of the path on real hardware (north_star (d); SURVEY.md 8(e), section 5: "~4.2 ms for C2 vs
0.23 ms compute"): BASELINE config 2 split over the N ranks (rows / N per GPU), filtered with no communication, then
(a) gathered to ONE rank over NCCL point-to-point (scir_b200.dist.gather_rows(dst=0)) and (b) all-gathered, each timed
on the device outside the hot path, and checked against rank 0 filtering the whole problem itself.

    torchrun --nproc-per-node N tools/nccl_gather_check.py        (prints one JSON line on rank 0)
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import CONFIGS, make_taps          # noqa: E402
from scir_b200 import _lib as L               # noqa: E402
from scir_b200 import dist as sdist           # noqa: E402
from scir_b200 import gpu                     # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS["c2"]
    rows, n = cfg["rows"], cfg["n"]
    taps = make_taps(cfg)
    r0, r1 = sdist.shard_rows(rows, world, rank)
    g = torch.Generator(device=dev).manual_seed(1234)              # same stream on every rank: replicated input
    x_full = torch.rand((rows, n), device=dev, generator=g) * 2 - 1
    x_local = x_full[r0:r1].contiguous()
    if rank != 0:
        del x_full
    y_local = gpu.fir1d_batched_f32_cuda(x_local, taps, tap_order=L.TAPS_LFILTER)
    torch.cuda.synchronize()
    dist.barrier()

    def timed(fn, reps=5):
        fn()                                                        # warm-up (NCCL channels)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), out

    ms_one, y0 = timed(lambda: sdist.gather_rows(y_local, rows, dst=0))
    ok_one = None
    if rank == 0:
        want = gpu.fir1d_batched_f32_cuda(x_full, taps, tap_order=L.TAPS_LFILTER)
        torch.cuda.synchronize()
        ok_one = bool(torch.allclose(y0, want, atol=1e-5))          # shards of different sizes may take different kernels
        del want
    del y0
    torch.cuda.empty_cache()
    ms_all, ya = timed(lambda: sdist.gather_rows(y_local, rows), reps=3)
    ok_all = bool(torch.equal(ya[r0:r1], y_local))
    flags = torch.tensor([1 if ok_all else 0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    # compute time of the sharded filtering itself, for scale
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(10):
        gpu.fir1d_batched_f32_cuda(x_local, taps, tap_order=L.TAPS_LFILTER, out=y_local)
    e1.record()
    torch.cuda.synchronize()
    tc = torch.tensor([e0.elapsed_time(e1) / 10], device=dev, dtype=torch.float64)
    dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    if rank == 0:
        moved = rows * n * 4 * (world - 1) / world
        print(json.dumps({"workload": cfg["desc"] + f" split over {world} GPUs", "n_gpus": world,
                          "filter_ms_max_over_ranks": float(tc.item()),
                          "gather_to_rank0": {"ms": ms_one, "bytes_into_rank0": moved, "gbs": moved / ms_one / 1e6, "matches_single_gpu": ok_one,
                                              "how": "NCCL point-to-point fan-in (dist.isend / irecv), scir_b200.dist.gather_rows(dst=0)"},
                          "all_gather": {"ms": ms_all, "bytes_per_rank": moved, "ok": bool(flags.item()),
                                         "how": "NCCL all_gather_into_tensor, scir_b200.dist.gather_rows()"},
                          "survey_estimate_ms": 4.2}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
