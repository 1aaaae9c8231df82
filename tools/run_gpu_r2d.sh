#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2d_bench1.json 2> gpurun_out/r2d_bench1.err; echo "bench rc $?"
tail -c 3000 gpurun_out/r2d_bench1.json; tail -5 gpurun_out/r2d_bench1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_ref1.json 2> gpurun_out/r2d_ref1.err; echo "ref rc $?"
cat gpurun_out/r2d_ref1.json | cut -c1-600
