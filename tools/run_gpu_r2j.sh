#!/bin/bash
mkdir -p gpurun_out
M3B_TMA_CTAS=4 timeout 600 python tools/kernel_bench.py tma4 > gpurun_out/r2j_kb_tma4.log 2>&1; echo "tma4 rc $?"
timeout 300 python tools/stage_time.py c3 20 > gpurun_out/r2j_tma3_c3.log 2>&1
timeout 300 python tools/stage_time.py c2 100 > gpurun_out/r2j_tma3_c2.log 2>&1
grep -h "TIMING\|PARITY\|STAGE" gpurun_out/r2j_*.log
