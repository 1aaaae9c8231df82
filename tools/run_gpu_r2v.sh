#!/bin/bash
# Box visit: update-input load variants of stage_tma (ld.global.nc, late select tied to both faces) on C3
mkdir -p gpurun_out
{
for rep in 1 2; do
python tools/stage_time.py c3 20
for v in nc lb nclb; do M3B_LIBRARY=$PWD/build/variants/$v.so python tools/stage_time.py c3 20; done
done
for v in nc lb nclb; do M3B_LIBRARY=$PWD/build/variants/$v.so python tools/kernel_bench.py $v 2>&1 | grep "PARITY\|TAG"; done
} > gpurun_out/r2v_variants.log 2>&1
grep -v "^$" gpurun_out/r2v_variants.log | cut -c1-200
