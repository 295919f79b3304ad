for r in 8192 4096 2048 1024; do
python bench.py --config c5 --rows $r --steps 10 --warmup 3 --no-cpu --no-e2e --no-sweep > gpurun_out/r2h/c5_$r.json 2> gpurun_out/r2h/c5_$r.err
python -c "
import json; d=json.loads(open('gpurun_out/r2h/c5_$r.json').read().strip().splitlines()[-1]); print($r, round(d['ms_per_step'],3), d['gpu'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['gpu_launches'])"
done
