#!/bin/bash
# Box visit: which of the update inputs go through ld.global.nc (bit 0 own state, bit 1 initial state + buffer rate, bit 2 step-start state)
mkdir -p gpurun_out
{
python tools/stage_time.py c3 20
for k in 1 2 3 4 5 6 7; do M3B_LIBRARY=$PWD/build/variants/nc$k.so python tools/stage_time.py c3 20; done
python tools/stage_time.py c3 20
} > gpurun_out/r2w_variants.log 2>&1
grep -v "^$" gpurun_out/r2w_variants.log | cut -c1-200
