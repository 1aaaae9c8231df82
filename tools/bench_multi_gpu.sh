#!/bin/bash
# N GPUs (first argument): bench line with parity_vs_1gpu and the exchange keys
N=${1:-8}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline $EXTRA_ARGS > gpurun_out/bench$N$TAG.json 2> gpurun_out/bench$N$TAG.err; echo "bench$N rc $?"
grep -o '"value": [0-9.]*' gpurun_out/bench$N$TAG.json | head -1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench$N$TAG.json | head -1; grep -o '"parity_vs_1gpu": {[^}]*}' gpurun_out/bench$N$TAG.json | cut -c1-120; grep -o '"exchange": {[^}]*}' gpurun_out/bench$N$TAG.json | cut -c200-; grep -o '"scaling_reference": {[^}]*}' gpurun_out/bench$N$TAG.json | cut -c1-80; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench$N$TAG.json | head -1; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench$N$TAG.json; grep -v "^\*\|OMP_NUM" gpurun_out/bench$N$TAG.err | tail -3
