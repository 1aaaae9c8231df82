# One GPU box visit of the development loop: parity tests, the default bench line, kernel timing.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err; tail -c 600 gpurun_out/bench_latest.json
python tools/kernel_bench.py latest 2>&1 | grep -E "PARITY|TIMING"
