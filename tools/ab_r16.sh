mkdir -p gpurun_out/r16
python -c "import __graft_entry__ as g; g.build()"
for rep in 1 2; do
for o in "" "--opt toeplitz_stcs=1"; do
  python bench.py --config c2 --steps 30 --warmup 5 --no-cpu --no-e2e $o | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 [$o]', round(d['value'],1), round(d['ms_per_step'],4))"
done; done
for o in "--opt toeplitz_stcs=0" "--opt toeplitz_stcs=1"; do
  python bench.py --config c5 --steps 10 --warmup 3 --no-cpu --no-e2e $o | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c5 [$o]', round(d['value'],1), round(d['ms_per_step'],4), d['clocks']['sm_mhz'])"
done
for r in 8 32 128; do
  python bench.py --config c2 --steps 5 --warmup 3 --no-cpu --opt host_block_rows=$r | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e rows=$r', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],2))"
done
