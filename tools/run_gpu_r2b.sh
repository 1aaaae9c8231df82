#!/bin/bash
mkdir -p gpurun_out
python tools/stage_time.py c3 6 > gpurun_out/r2b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stage_tma -s 6 -c 2 -o gpurun_out/r2b_tma -f python tools/stage_time.py c3 6 > gpurun_out/r2b_ncu.log 2>&1
cat gpurun_out/r2b_plain.log; tail -5 gpurun_out/r2b_ncu.log
