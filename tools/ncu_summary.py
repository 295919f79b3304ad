#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page + per-opcode instruction mix + top stall sites) as text.
Usage: tools/ncu_summary.py <report.ncu-rep> [kernel-index]"""
import csv, subprocess, sys, io
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
ix = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("== kernel:", r[ix["Kernel Name"]][:110], "grid", r[ix.get("Grid Size", 0)], "block", r[ix.get("Block Size", 0)])
    for k in KEYS:
        if k in ix:
            print(f"  {k:90s} {units[ix[k]]:10s} {r[ix[k]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# the source page may hold several kernels: split on header rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
for b in blocks:
    h = {n: i for i, n in enumerate(b["hdr"])}
    data = b["rows"]
    tot = sum(int(r[h["Instructions Executed"]]) for r in data) or 1
    print("== instruction mix:", b["name"][:100], "total warp-instructions", tot)
    c = Counter()
    for r in data:
        t = r[h["Source"]].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        c[op] += int(r[h["Instructions Executed"]])
    print("  " + ", ".join(f"{op} {100*n/tot:.2f}%" for op, n in c.most_common(14)))
    print("  top not-issued stall sites:")
    ni = "Warp Stall Sampling (Not-issued Samples)"
    for r in sorted(data, key=lambda r: -int(r[h[ni]]))[:10]:
        why = {k.replace(" (Not Issued)", ""): r[h[k]] for k in b["hdr"] if k.endswith("(Not Issued)") and r[h[k]] not in ("0", "")}
        print(f"    {r[h['Source']].strip()[:58]:58s} samples={r[h[ni]]:>7s} exec={r[h['Instructions Executed']]:>10s} {why}")
