set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; tail -60 gpurun_out/multi_check.log
