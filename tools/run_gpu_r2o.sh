#!/bin/bash
# 2 GPUs: the single-buffer (4 CTAs per SM) variant of stage_tma with the fused exchange, forced; bench line with parity key
mkdir -p gpurun_out
export M3B_TMA_CTAS=4 TAG=_ctas4
bash tools/run_gpu_r2f.sh 2
