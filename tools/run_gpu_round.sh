set -x
MGC_QUICK=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu_check.py 2>&1 | grep -E "field_err|MULTI_GPU_CHECK|Error|error" | head
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/bench_8gpu_s12.json 2> gpurun_out/bench_8gpu_s12.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_8gpu_s12.json').read().strip().splitlines()[-1]); print('C3 N=8 value', d['value'], 'ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'ref1', d['scaling_reference']['value'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --workload c5 --steps 15 --warmup 3 --no-scaling-reference > gpurun_out/bench_c5_8gpu_s12.json 2> gpurun_out/bench_c5_8gpu_s12.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_c5_8gpu_s12.json').read().strip().splitlines()[-1]); print('C5 N=8 value', d['value'], 'ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], d['roofline']['frac'])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/trace_steps.py 6 2>&1 | grep -v OMP | tail -12
