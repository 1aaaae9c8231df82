#!/bin/bash
# 2-GPU box visit: the multi-GPU tests, the C3 bench line (NVLink keys, parity_vs_1gpu) and the nested c4x workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2s_pytest_mgpu.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r2s_pytest_mgpu.log
TAG=_r2s bash tools/bench_multi_gpu.sh 2
run() {
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline $2 > gpurun_out/r2s_bench_$1_2.json 2> gpurun_out/r2s_bench_$1_2.err; echo "$1 rc $?"
  grep -o '"value": [0-9.]*' gpurun_out/r2s_bench_$1_2.json | head -1; grep -o '"parity_vs_1gpu": {[^}]*}' gpurun_out/r2s_bench_$1_2.json | cut -c1-110; grep -o '"scaling_reference": {[^}]*}' gpurun_out/r2s_bench_$1_2.json | cut -c1-80; grep -v "^\*\|OMP_NUM" gpurun_out/r2s_bench_$1_2.err | tail -3
}
run c4x "--workload c4x --no-e2e"
