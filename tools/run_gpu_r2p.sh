#!/bin/bash
# Box visit: timings of the split of the jump blocks' tiles (c4, c4x) and of the uniform workloads, then the GPU suite.
mkdir -p gpurun_out
{
for w in c4 c4x; do
  python tools/stage_time.py $w 40
  M3B_JUMP_AFTER=0 python tools/stage_time.py $w 40
  M3B_SPLIT_JUMP=0 python tools/stage_time.py $w 40
done
python tools/stage_time.py c3 20
python tools/stage_time.py c2 100
} > gpurun_out/r2p_timing.log 2>&1
tail -8 gpurun_out/r2p_timing.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r2p_pytest.log
