#!/bin/bash
# bash tools/e2e_n.sh <tag> <N> "<opt specs separated by ;>"   e.g. "default;host_block_rows=32"
TAG=$1; N=$2; SPECS=$3
OUT=gpurun_out/$TAG; mkdir -p $OUT
IFS=';' read -ra arr <<< "$SPECS"
for spec in "${arr[@]}"; do
  o=""; [ "$spec" != default ] && for kv in $spec; do o="$o --opt $kv"; done
  name=${spec// /_}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-sweep --no-cpu $o > $OUT/e2e_$name.json 2> $OUT/e2e_$name.err
  python - "$OUT/e2e_$name.json" "$spec" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    e = d["e2e"]
    print(sys.argv[2], "| e2e", round(e["value"], 2), "Gs/s", round(e["ms_per_step"], 1), "ms", round(e["gbs_each_way_per_gpu"], 1), "GB/s/GPU",
          "| pageable", round((e.get("pageable") or {}).get("value", 0), 2), "| mg", (e.get("mg") or {}).get("value"), (e.get("mg") or {}).get("ms_per_step"))
    print("   ceiling", e.get("ceiling"))
except Exception as ex:
    print(sys.argv[2], "failed", ex)
PY
done
