python tools/nested_timing.py
python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C5 N=1 value', d['value'], 'ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], d['roofline']['frac'])"
