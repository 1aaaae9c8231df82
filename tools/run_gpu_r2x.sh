#!/bin/bash
# Box visit: default build (own state through ld.global.nc) against: late select tied to both faces, streaming stores, shallow pipeline
mkdir -p gpurun_out
{
for rep in 1 2; do
python tools/stage_time.py c3 20
for v in lb cs cslb nodeep; do M3B_LIBRARY=$PWD/build/variants/$v.so python tools/stage_time.py c3 20; done
done
python tools/kernel_bench.py default 2>&1 | grep "PARITY\|TAG\|TIMING"
} > gpurun_out/r2x_variants.log 2>&1
grep -v "^$" gpurun_out/r2x_variants.log | cut -c1-200
