#!/bin/bash
# Box visit: the three-barrier variant of stage_tma (parity + timing) beside the default build; 2 CTAs per SM for insight
mkdir -p gpurun_out
{
M3B_LIBRARY=$PWD/build/variants/three.so python tools/kernel_bench.py three
for rep in 1 2; do
python tools/stage_time.py c3 20
M3B_LIBRARY=$PWD/build/variants/three.so python tools/stage_time.py c3 20
done
M3B_TMA_CTAS=2 python tools/stage_time.py c3 20
python tools/stage_time.py c2 100
M3B_LIBRARY=$PWD/build/variants/three.so python tools/stage_time.py c2 100
} > gpurun_out/r2u_variants.log 2>&1
grep -v "^$" gpurun_out/r2u_variants.log | cut -c1-200
