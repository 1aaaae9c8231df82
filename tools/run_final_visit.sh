#!/bin/bash
bash tools/run_profile_r02.sh
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in c4 c4x c2; do python tools/stage_time.py $w 40; done 2>&1 | tee gpurun_out/r02f_secondary_timing.log | cut -c1-160
