#!/bin/bash
# Box visit: stage_tma build variants on C3 / C2 (M3B_LIBRARY), launch list of the nested c4 workload.
mkdir -p gpurun_out
{
for rep in 1 2; do
python tools/stage_time.py c3 20
for v in rotate nolate plm1 plm5; do echo "variant $v"; M3B_LIBRARY=$PWD/build/variants/$v.so python tools/stage_time.py c3 20; done
done
python tools/stage_time.py c2 100
for v in rotate nolate; do echo "variant $v"; M3B_LIBRARY=$PWD/build/variants/$v.so python tools/stage_time.py c2 100; done
} > gpurun_out/r2q_variants.log 2>&1
grep -v "^$" gpurun_out/r2q_variants.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r2q_launches_c4.csv python tools/stage_time.py c4 12 > gpurun_out/r2q_ncu_c4.log 2>&1; echo "ncu rc $?"
