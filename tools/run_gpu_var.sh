#!/bin/bash
# Box visit: build variants of stage_tma.cu (tools/build_variant.sh) beside the default, C3, 20 steps each.  usage: run_gpu_var.sh name ...
mkdir -p gpurun_out
{
python tools/stage_time.py c3 20
for v in "$@"; do M3B_LIBRARY=$PWD/build/variants/$v.so python tools/stage_time.py c3 20; done
python tools/stage_time.py c3 20
} > gpurun_out/variants.log 2>&1
grep -v "^$" gpurun_out/variants.log | cut -c1-200
