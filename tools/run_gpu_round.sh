for m in 4 3 2; do
  echo "== M3B_STRIP_MIN_CTAS=$m"
  M3B_STRIP_MIN_CTAS=$m python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'])"
  M3B_STRIP_MIN_CTAS=$m python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c3 value', d['value'], 'ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'])"
done
