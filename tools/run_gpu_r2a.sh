#!/bin/bash
# Round 2, first box visit: does stage_tma agree with the oracle, and what does it cost against stage_strip?
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
tools/micro/seed_accuracy > gpurun_out/r2a_seed.log 2>&1
timeout 600 python tools/kernel_bench.py tma3 > gpurun_out/r2a_kb_tma3.log 2>&1; echo "tma3 rc $?" >> gpurun_out/r2a_rc.log
M3B_STAGE=strip timeout 600 python tools/kernel_bench.py strip > gpurun_out/r2a_kb_strip.log 2>&1; echo "strip rc $?" >> gpurun_out/r2a_rc.log
M3B_TMA_CTAS=2 timeout 600 python tools/kernel_bench.py tma2 > gpurun_out/r2a_kb_tma2.log 2>&1; echo "tma2 rc $?" >> gpurun_out/r2a_rc.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2a_rc.log
tail -3 gpurun_out/r2a_pytest.log
grep -h "TIMING\|PARITY" gpurun_out/r2a_kb_*.log
cat gpurun_out/r2a_rc.log gpurun_out/r2a_seed.log
