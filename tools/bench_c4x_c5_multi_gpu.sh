#!/bin/bash
# 8 GPUs: the nested workload with load (c4x) and config 5 (16384^2) with checkpoint / diagnostics I/O timed
mkdir -p gpurun_out
N=${1:-8}
run() {  # name, extra args
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline $2 > gpurun_out/bench_$1_$N.json 2> gpurun_out/bench_$1_$N.err; echo "$1 rc $?"
  grep -o '"value": [0-9.]*' gpurun_out/bench_$1_$N.json | head -1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_$1_$N.json | head -1; grep -o '"parity_vs_1gpu": {[^}]*}' gpurun_out/bench_$1_$N.json | cut -c1-110; grep -o '"scaling_reference": {[^}]*}' gpurun_out/bench_$1_$N.json | cut -c1-80; grep -o '"io": {[^}]*}' gpurun_out/bench_$1_$N.json | cut -c1-400; grep -o '"frac": [0-9.]*' gpurun_out/bench_$1_$N.json | head -1; grep -v "^\*\|OMP_NUM" gpurun_out/bench_$1_$N.err | tail -3
}
run c4x "--workload c4x --no-e2e"
run c5 "--workload c5 --io --no-e2e"
df -h /tmp | tail -1
