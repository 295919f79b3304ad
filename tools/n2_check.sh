# 2-GPU check of the torchrun path: configs 2, 5, 4 (rows sharded, no collective) + the in-process mg front end
port=29540
for cfg in c2 c5 c4; do
  port=$((port+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus 2 --config $cfg --steps 10 --warmup 3 --no-cpu > /tmp/n2_$cfg.json 2> /tmp/n2_$cfg.err
  python -c "
import json
try:
    d=json.loads(open('/tmp/n2_$cfg.json').read().strip().splitlines()[-1])
    print(d['config']['config'], 'N=2', round(d['value'],1), 'Gs/s', round(d['ms_per_step'],3), 'ms  e2e', d['e2e'] and round(d['e2e']['value'],2))
except Exception as e:
    print('$cfg failed', e); print(open('/tmp/n2_$cfg.err').read()[-1500:])
"
done
timeout 300 python -m pytest tests -m gpu -q -k "mg_" 2>&1 | tail -2
