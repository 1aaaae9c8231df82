#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/kernel_bench.py tma3 > gpurun_out/r2c_kb_tma3.log 2>&1; echo "tma3 rc $?"
M3B_TMA_CTAS=2 timeout 300 python tools/stage_time.py c3 20 > gpurun_out/r2c_tma2.log 2>&1
grep -h "TIMING\|PARITY\|STAGE" gpurun_out/r2c_*.log
python tools/stage_time.py c3 6 > gpurun_out/r2c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stage_tma -s 6 -c 2 -o gpurun_out/r2c_tma -f python tools/stage_time.py c3 6 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
