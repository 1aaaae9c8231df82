#!/bin/bash
mkdir -p gpurun_out
for w in c4 c4x c5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2k_bench_$w.json 2> gpurun_out/r2k_bench_$w.err; echo "$w rc $?"
  grep -o '"value": [0-9.]*' gpurun_out/r2k_bench_$w.json | head -1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2k_bench_$w.json | head -1; grep -o '"frac": [0-9.]*' gpurun_out/r2k_bench_$w.json | head -1; grep -o '"kernel_ms": [0-9.]*' gpurun_out/r2k_bench_$w.json | head -1; tail -2 gpurun_out/r2k_bench_$w.err
done
