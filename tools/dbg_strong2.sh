mkdir -p gpurun_out/r2k
for tag in noe2e full; do
  extra=""; [ $tag = noe2e ] && extra="--no-e2e"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 $extra > gpurun_out/r2k/$tag.json 2> gpurun_out/r2k/$tag.err
  python - $tag <<'PY'
import json, sys
d=json.loads([l for l in open('gpurun_out/r2k/%s.json' % sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
for k,v in d['configs'].items(): print(sys.argv[1], k, [round(x,3) for x in v['ms_by_rank']], v['step_ms_rank0'])
for k,v in d['strong'].items(): print(sys.argv[1], 'strong',k, [round(x,3) for x in v['ms_by_rank']], v['step_ms_rank0'])
PY
done
