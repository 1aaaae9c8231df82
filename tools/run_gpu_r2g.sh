#!/bin/bash
# 2 GPUs: quick multi-GPU parity (tools/multi_gpu_check.py, quick cases) then the bench line; everything under short timeouts
mkdir -p gpurun_out
MGC_QUICK=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/r2g_mgc.log 2>&1; echo "mgc rc $?"; grep -v "^\*\|OMP_NUM" gpurun_out/r2g_mgc.log | tail -4 | cut -c1-400
bash tools/run_gpu_r2f.sh 2
