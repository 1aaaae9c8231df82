#!/bin/bash
# 2 GPUs: per-segment device timeline of small steps (fixed cost per step): depth 4 (C2: 0.5 M cells per rank) and depth 5
mkdir -p gpurun_out
for d in 4 5; do
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/trace_steps.py $d > gpurun_out/r2m_trace_d$d.log 2>&1; echo "trace d$d rc $?"
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2m_trace_d$d.log | tail -16
done
timeout 100 python tools/trace_steps.py 4 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -9
