#!/bin/bash
# 2 GPUs: bench line with parity_vs_1gpu (every multi-rank command under its own short timeout)
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench2.json 2> gpurun_out/r2e_bench2.err; echo "bench2 rc $?"
grep -o '"value": [0-9.]*' gpurun_out/r2e_bench2.json | head -1; grep -o '"parity_vs_1gpu": {[^}]*}' gpurun_out/r2e_bench2.json; grep -o '"exchange": {[^}]*}' gpurun_out/r2e_bench2.json; grep -o '"scaling_reference": {[^}]*}' gpurun_out/r2e_bench2.json; grep -o '"roofline": {[^}]*}' gpurun_out/r2e_bench2.json; tail -3 gpurun_out/r2e_bench2.err
