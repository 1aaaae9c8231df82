#!/bin/bash
# Instruction census of stage_tma's common path (no sink within reach of the tile, no negative density), per stage variant.
# No GPU needed: compiles stage_tma.cu with -DM3B_HOT_PATH_ONLY into /tmp and counts SASS opcodes (tools/sass_count.py).
set -e
cd "$(dirname "$0")/../mara3_b200/csrc"
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -DM3B_HOT_PATH_ONLY $EXTRA -cubin -o /tmp/stage_tma_hot.cubin stage_tma.cu
for v in "${@:-stage_tma<3, 64, true, 1, 2>}" ; do python ../../tools/sass_count.py /tmp/stage_tma_hot.cubin "$v"; done
