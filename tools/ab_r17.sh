python -c "import __graft_entry__ as g; g.build()"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for o in "--opt long_tap_path=1" "--opt long_tap_path=1 --opt ffma2=0" "--variant 3" "--variant 3 --opt ffma2=0"; do
  python bench.py --config c2 --steps 20 --warmup 3 --no-cpu --no-e2e $o | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 [$o]', round(d['value'],1), round(d['ms_per_step'],4))"
done
for o in "--opt long_tap_path=1" "--opt long_tap_path=1 --opt ffma2=0"; do
  python bench.py --config c5 --steps 5 --warmup 3 --no-cpu --no-e2e $o | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c5 [$o]', round(d['value'],1), round(d['ms_per_step'],4))"
done
