#!/bin/bash
# Box visit: timings after the sink-path changes (reach 40, sums by recursive halving) and the rotated warp roles, then the GPU suite.
mkdir -p gpurun_out
{
for w in c4 c4x c3 c2; do python tools/stage_time.py $w 40; done
python tools/nested_timing.py 2>&1 | tail -3
} > gpurun_out/r2r_timing.log 2>&1
cat gpurun_out/r2r_timing.log | cut -c1-220
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r2r_pytest.log
