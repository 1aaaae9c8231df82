#!/bin/bash
# Development aid: a second build of the library that differs in stage_tma.cu's compile flags only, for M3B_LIBRARY.
# usage: tools/build_variant.sh NAME -DFLAG ...   ->  build/variants/NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
nvcc "$@" -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -c mara3_b200/csrc/stage_tma.cu -o build/variants/$name.o
objs=$(ls build/csrc/*.o | grep -v stage_tma.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so $objs build/variants/$name.o -cudart shared
echo build/variants/$name.so
