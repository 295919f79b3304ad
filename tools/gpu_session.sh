#!/bin/bash
# One gpurun call: GPU parity tests, bench lines for the configs, ncu launch list + full captures.
# Usage (from the repo root on the GPU box): bash tools/gpu_session.sh <tag> [stages...]
#   stages: test smoke bench ref launches ncu_c2 ncu_c4 ncu_c5 ncu_c3 (default: all but ncu_c5/ncu_c3)
TAG=${1:-r01}; shift
STAGES=${*:-test smoke bench ref launches ncu_c2 ncu_c4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
has() { [[ " $STAGES " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/nvidia_smi.csv 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { echo BUILD FAILED; tail -30 $OUT/build.log; }
if has toep; then
  timeout 600 python -m pytest tests/test_gpu_toeplitz.py -x -q -s > $OUT/pytest_toep.log 2>&1; echo "toeplitz pytest rc=$?" | tee -a $OUT/pytest_toep.log
  tail -40 $OUT/pytest_toep.log
fi
if has test; then
  timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -25 $OUT/pytest_gpu.log
fi
if has smoke; then
  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
fi
if has bench; then
  for cfg in ${BENCH_CFGS:-c2 c1 c4 c5 c3}; do
    steps=20; [ $cfg = c3 ] && steps=5
    extra=""; [ $cfg != c2 ] && extra="--no-cpu"
    timeout 600 python bench.py --config $cfg --steps $steps --warmup 3 $extra > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err; echo "bench $cfg rc=$?"
    python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$cfg.json"))
    r=d["roofline"]
    print("$cfg", "value=%.1f"%d["value"], "ms=%.3f"%d["ms_per_step"], "fp32_frac=%.3f"%r["fp32"]["frac_nominal"], "hbm_frac=%.3f"%r["frac"], "shape_frac=%.3f"%r["shape_roofline"]["frac"], "e2e=", (d.get("e2e") or {}).get("value"), "clk", d["clocks"])
except Exception as e:
    print("$cfg bench parse failed", e); print(open("$OUT/bench_$cfg.err").read()[-2000:])
PY
  done
  for v in ${BENCH_VARIANTS:-3}; do
    timeout 300 python bench.py --config c2 --steps 20 --warmup 3 --no-cpu --no-e2e --variant $v > $OUT/bench_c2_v$v.json 2> $OUT/bench_c2_v$v.err
    python -c "import json; d=json.load(open('$OUT/bench_c2_v$v.json')); print('c2 variant $v', d['value'], d['ms_per_step'])"
  done
fi
if has toepbench; then
  for spec in "c3 long_tap_path=2" "c3 toeplitz_ts=0" \
              "c5 long_tap_path=2" "c5 toeplitz_ts=0" "c5 toeplitz_loader=1" \
              "c2 long_tap_path=2" "c2 toeplitz_ts=0" "c2 toeplitz_loader=1" "c2 toeplitz_split=1" "c2 toeplitz_terms=4" \
              "c2 long_tap_path=1" "c5 long_tap_path=1"; do
    set -- $spec; cfg=$1; shift; o=""; tag=$cfg; for kv in "$@"; do o="$o --opt $kv"; tag="${tag}_${kv%%=*}${kv#*=}"; done
    timeout 300 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu --no-e2e $o > $OUT/benchT_$tag.json 2> $OUT/benchT_$tag.err
    python -c "import json; d=json.load(open('$OUT/benchT_$tag.json')); print('$spec ->', round(d['value'],2), 'Gs/s', round(d['ms_per_step'],3), 'ms', 'tc_launches', d['config'].get('tensor_core_launches'))" || tail -5 $OUT/benchT_$tag.err
  done
fi
if has ncu_toep; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fir_toeplitz -s 3 -c 1 -f -o $OUT/prof_c3_toeplitz \
     python bench.py --config c3 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_c3_toeplitz.log 2>&1; echo "ncu toeplitz c3 rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fir_toeplitz -s 3 -c 1 -f -o $OUT/prof_c2_toeplitz \
     python bench.py --config c2 --opt long_tap_path=2 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_c2_toeplitz.log 2>&1; echo "ncu toeplitz c2 rc=$?"
fi
if has ncu_toep_c5; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fir_toeplitz -s 6 -c 2 -f -o $OUT/prof_c5_toeplitz \
     python bench.py --config c5 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_c5_toeplitz.log 2>&1; echo "ncu toeplitz c5 rc=$?"
fi
if has ncu_toep_c2; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fir_toeplitz -s 3 -c 1 -f -o $OUT/prof_c2_toeplitz \
     python bench.py --config c2 --opt long_tap_path=2 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_c2_toeplitz.log 2>&1; echo "ncu toeplitz c2 rc=$?"
fi
if has ref; then
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
fi
if has launches; then
  # launch list of the default bench command (cold-cache, serialised: shares only)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_c2.csv \
     python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
fi
for cfg in c2 c4 c5 c3; do
  if has ncu_$cfg; then
    pat="regex:fir_|upfirdn_"
    timeout 900 ncu --set full --clock-control none --import-source on -k "$pat" -s 3 -c 1 -f -o $OUT/prof_$cfg \
       python bench.py --config $cfg --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_$cfg.log 2>&1; echo "ncu full $cfg rc=$?"
  fi
done
ls -la $OUT
