#!/bin/bash
# Multi-GPU session on ONE box: bash tools/mgpu_session.sh <tag> <N> [stages...]   (stages: bench gather test mgtime)
TAG=$1; N=$2; shift 2
STAGES=${*:-bench gather test}
OUT=gpurun_out/$TAG
mkdir -p $OUT
has() { [[ " $STAGES " == *" $1 "* ]]; }
nvidia-smi topo -m > $OUT/topo.txt 2>&1; nproc >> $OUT/topo.txt; free -g >> $OUT/topo.txt
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { echo BUILD FAILED; tail -30 $OUT/build.log; }
if has bench; then
  SECONDS=0
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench N=$N rc=$? wall ${SECONDS}s"
  tail -3 $OUT/bench_n$N.err
  python - $OUT/bench_n$N.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
except Exception as e:
    print("parse failed", e); sys.exit(0)
print("HEAD value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "n_gpus", d["n_gpus"], "clk", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
for k, v in (d.get("configs") or {}).items():
    print(" ", k, {kk: (round(v[kk], 4) if isinstance(v.get(kk), float) else v.get(kk)) for kk in ("ms_per_step", "value", "error")}, "parity", (v.get("parity") or {}).get("frac"))
for k, v in (d.get("strong") or {}).items():
    print("  strong", k, {kk: v.get(kk) for kk in ("rows_per_gpu", "ms", "value", "one_gpu_ms", "speedup_vs_n1", "efficiency_vs_n1", "error")})
e = d.get("e2e") or {}
print("  e2e", {k: e.get(k) for k in ("value", "ms_per_step", "gbs_each_way_per_gpu", "matches_device_path", "error")})
print("  e2e.pageable", e.get("pageable")); print("  e2e.ceiling", e.get("ceiling")); print("  e2e.mg", e.get("mg"))
PY
fi
if has gather; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
      tools/nccl_gather_check.py > $OUT/gather_n$N.json 2> $OUT/gather_n$N.err; echo "gather rc=$?"; cat $OUT/gather_n$N.json; tail -3 $OUT/gather_n$N.err
fi
if has test; then
  timeout 600 python -m pytest tests -m gpu -q -x -k "mg_ or current_device or gather" > $OUT/pytest_mg.log 2>&1; echo "pytest(mg) rc=$?"; tail -5 $OUT/pytest_mg.log
fi
ls $OUT
