#!/bin/bash
# End-of-round check on one GPU: the GPU suite, smoke(), the default bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py > gpurun_out/r02f_bench_default.json 2> gpurun_out/r02f_bench_default.err ) 2> gpurun_out/r02f_bench_default.time; tail -3 gpurun_out/r02f_bench_default.time
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02f_bench_default.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
