#!/bin/bash
# Round 2 (final kernel) profile visit (1 GPU): the bench command plain, its launch list, and one full capture of the stage kernel.
mkdir -p gpurun_out
# the default bench line first (what the driver runs), with its wall time
( time python bench.py > gpurun_out/r02f_bench_default.json 2> gpurun_out/r02f_bench_default.err ) 2> gpurun_out/r02f_bench_default.time; tail -3 gpurun_out/r02f_bench_default.time
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary"
$CMD > gpurun_out/r02f_plain.json 2> gpurun_out/r02f_plain.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r02f_launches_bench_c3.csv $CMD > gpurun_out/r02f_ncu_list.log 2>&1
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stage_tma -s 8 -c 2 -o gpurun_out/r02f_stage_tma -f $CMD > gpurun_out/r02f_ncu_full.log 2>&1
tail -2 gpurun_out/r02f_ncu_full.log; wc -l gpurun_out/r02f_launches_bench_c3.csv
