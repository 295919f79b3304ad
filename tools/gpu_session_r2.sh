#!/bin/bash
# Round-2 gpurun session: bash tools/gpu_session_r2.sh <tag> [stages...]
#   stages: test smoke bench ref pageable launches ncu_c4 ncu_c5 ncu_c2 ncu_c3 f64 (default: test smoke bench ref)
TAG=${1:-r2a}; shift
STAGES=${*:-test smoke bench ref}
OUT=gpurun_out/$TAG
mkdir -p $OUT
has() { [[ " $STAGES " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,memory.total --format=csv > $OUT/nvidia_smi.csv 2>&1
nproc > $OUT/host.txt; free -g >> $OUT/host.txt; lscpu | head -25 >> $OUT/host.txt; nvidia-smi topo -m >> $OUT/host.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { echo BUILD FAILED; tail -30 $OUT/build.log; }
if has test; then
  timeout 1800 python -m pytest tests -m gpu -q -x --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -40 $OUT/pytest_gpu.log
fi
if has testfast; then
  timeout 900 python -m pytest tests -m gpu -q -x ${PYTEST_K:+-k "$PYTEST_K"} > $OUT/pytest_fast.log 2>&1; echo "pytest(fast) rc=$?" | tee -a $OUT/pytest_fast.log
  tail -30 $OUT/pytest_fast.log
fi
if has smoke; then
  timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
fi
summ() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("parse failed", e); sys.exit(0)
def p(name, r):
    if not r or "error" in r: print(name, r); return
    rf = r["roofline"]; par = r.get("parity") or {}
    print(f"{name}: {r['ms_per_step']:.4f} ms  {r['value']:.1f} Gs/s  {rf['bound']} frac {rf['frac']:.3f}  shape {rf['shape_roofline']['frac']:.3f}  "
          f"parity {par.get('frac')} ok={par.get('ok')} edges={(par.get('edges') or {}).get('frac')}  clk {r['clocks'].get('sm_mhz')} {r['clocks'].get('reasons')}")
print("HEAD", d["config"]["config"], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "n_gpus", d["n_gpus"])
for k, v in (d.get("configs") or {}).items(): p(k, v)
for k, v in (d.get("strong") or {}).items(): print("strong", k, {kk: v.get(kk) for kk in ("rows_per_gpu", "ms", "value", "one_gpu_ms", "efficiency_vs_n1", "error")})
e = d.get("e2e") or {}
print("e2e", {k: e.get(k) for k in ("value", "ms_per_step", "gbs_each_way_per_gpu", "matches_device_path", "ceiling_gbs", "error")})
print("e2e.pageable", e.get("pageable")); print("e2e.ceiling", e.get("ceiling")); print("e2e.mg", e.get("mg"))
print("cpu", d.get("cpu_baseline"))
PY
}
if has bench; then
  SECONDS=0; timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
  echo "bench wall: $SECONDS s"
  summ $OUT/bench.json
fi
if has ref; then
  SECONDS=0; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
  echo "ref wall: $SECONDS s"; cut -c1-600 $OUT/bench_ref.json
fi
if has pageable; then
  for spec in ${PAGEABLE_SPECS:-"host_stage_nt=1" "host_stage_nt=0" "host_copy_threads=4" "host_copy_threads=12" "host_block_rows=4" "host_block_rows=1" "host_stage_wc=1"}; do
    o=""; tag=""; for kv in $spec; do o="$o --opt $kv"; tag="${tag}_${kv%%=*}${kv#*=}"; done
    timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-sweep $o > $OUT/benchP$tag.json 2> $OUT/benchP$tag.err
    python -c "
import json; d=json.loads(open('$OUT/benchP$tag.json').read().strip().splitlines()[-1]); e=d['e2e']
print('$spec -> pinned', round(e['value'],2), 'pageable', e['pageable'])" || tail -5 $OUT/benchP$tag.err
  done
fi
if has abench; then    # A/B lines: BENCH_AB="c4:upfirdn_variant=7 c5:toeplitz_m=64 ..."
  for spec in $BENCH_AB; do
    cfg=${spec%%:*}; kvs=${spec#*:}; o=""; [ "$kvs" != "$cfg" ] && for kv in ${kvs//,/ }; do o="$o --opt $kv"; done
    timeout 300 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu --no-e2e --no-sweep $o > $OUT/benchAB_${spec//[:=,]/_}.json 2> $OUT/benchAB_${spec//[:=,]/_}.err
    python -c "
import json; d=json.loads(open('$OUT/benchAB_${spec//[:=,]/_}.json').read().strip().splitlines()[-1]); r=d['roofline']
print('$spec ->', round(d['ms_per_step'],4), 'ms', round(d['value'],1), 'Gs/s', r['bound'], round(r['frac'],3), 'parity', (d.get('parity') or {}).get('frac'), 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))" || tail -5 $OUT/benchAB_${spec//[:=,]/_}.err
  done
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
     python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-sweep > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches_all_configs.csv \
     python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_launches_all.log 2>&1; echo "ncu launches (all configs) rc=$?"
fi
for cfg in c2 c4 c5 c3; do
  if has ncu_$cfg; then
    pat="regex:fir_toeplitz|fir_tile|fir_stream|upfirdn_|fir_os"
    timeout 900 ncu --set full --clock-control none --import-source on -k "$pat" -s 3 -c 1 -f -o $OUT/prof_$cfg \
       python bench.py --config $cfg --steps 1 --warmup 3 --no-e2e --no-cpu --no-sweep $NCU_OPTS > $OUT/ncu_full_$cfg.log 2>&1; echo "ncu full $cfg rc=$?"
  fi
done
ls -la $OUT | head -50
