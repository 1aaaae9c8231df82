#!/bin/bash
# 2 GPUs: everything that needs a GPU once more after the split of kernels.cu (1-GPU suite, then the 2-GPU tests)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r2n_pytest.log
