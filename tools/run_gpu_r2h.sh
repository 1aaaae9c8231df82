#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/stalled_peer_check.py > gpurun_out/r2h_stalled.log 2>&1; echo "stalled rc $?"; grep "STALLED\|Error\|error" gpurun_out/r2h_stalled.log | head -5
timeout 400 python -m pytest tests/test_multi_gpu.py -x -q -k "products" > gpurun_out/r2h_pytest_products.log 2>&1; echo "pytest products rc $?"; tail -5 gpurun_out/r2h_pytest_products.log
