set -x
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/pytest_gpu.log | head -40
CMD="python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/bench3.json 2> gpurun_out/bench3.err; cat gpurun_out/bench3.json
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stage_strip -s 6 -c 2 -o gpurun_out/prof_strip2 $CMD > gpurun_out/ncu2.log 2>&1
