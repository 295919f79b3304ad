mkdir -p gpurun_out/r2j
python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --strong-div 2 > gpurun_out/r2j/a.json 2> gpurun_out/r2j/a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j/a.json').read().strip().splitlines()[-1])
for k,v in d['configs'].items(): print(k, v['ms_by_rank'], v['gpu_launches'], v['tensor_core_launches'], v['fft_launches'])
for k,v in d['strong'].items(): print('strong',k, v['ms_by_rank'], v['gpu_launches'])
PY
