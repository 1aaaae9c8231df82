set -x
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --workload c5 --steps 20 --warmup 3 --no-scaling-reference > gpurun_out/bench_c5_8gpu.json 2> gpurun_out/bench_c5_8gpu.err; tail -2 gpurun_out/bench_c5_8gpu.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_c5_8gpu.json').read().strip().splitlines()[-1]); print('C5 N=8 value', d['value'], 'ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], d['roofline']['frac'], d['exchange'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_8gpu.json').read().strip().splitlines()[-1]); print('C3 N=8 value', d['value'], 'ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'ref1', d['scaling_reference']['value'])"
