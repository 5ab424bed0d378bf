#!/bin/bash
# A second build of the library with in-kernel cycle counters compiled in (-DGSAGE_POOL_TIMING / -DGSAGE_WS_TIMING: linear_pool_ws_umma.cu and
# linear_ws_umma.cu print per-role cycles per tile after every (big) launch).  Run after `python -m pytorch_graphsage_b200.build`; use with
#   GSAGE_B200_LIB=$PWD/pytorch_graphsage_b200/libgsage_b200_timing.so python ...
set -e
cd "$(dirname "$0")/../../pytorch_graphsage_b200"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p build/timing
nvcc $FLAGS -DGSAGE_POOL_TIMING -c csrc/linear_pool_ws_umma.cu -o build/timing/linear_pool_ws_umma.o
nvcc $FLAGS -DGSAGE_WS_TIMING -c csrc/linear_ws_umma.cu -o build/timing/linear_ws_umma.o
OBJS=$(ls build/*.o | grep -v -E "/linear_pool_ws_umma.o|/linear_ws_umma.o")
nvcc -shared -o libgsage_b200_timing.so $OBJS build/timing/linear_pool_ws_umma.o build/timing/linear_ws_umma.o -gencode arch=compute_100a,code=sm_100a -lcudart_static -ldl -lrt -lpthread
echo built $PWD/libgsage_b200_timing.so
