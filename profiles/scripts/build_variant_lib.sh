#!/bin/bash
# Variant build for an A/B measurement: the gather producers of linear_ws_umma / attention_umma / wgrad_umma with ONE elected lane per
# warp issuing its gather4 back to back (profiles/experiments/*_elect_producers.cu.txt) instead of 4-8 lanes per warp each issuing one.
#   GSAGE_B200_LIB=$PWD/pytorch_graphsage_b200/libgsage_b200_electprod.so python bench.py ...
set -e
cd "$(dirname "$0")/../../pytorch_graphsage_b200"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Icsrc -I../include"
mkdir -p build/variant
SKIP=""
for f in linear_ws_umma attention_umma wgrad_umma; do
  cp ../profiles/experiments/${f}_elect_producers.cu.txt build/variant/$f.cu
  nvcc $FLAGS -c build/variant/$f.cu -o build/variant/$f.o
  SKIP="$SKIP|/$f.o"
done
OBJS=$(ls build/*.o | grep -v -E "${SKIP:1}")
nvcc -shared -o libgsage_b200_electprod.so $OBJS build/variant/*.o -gencode arch=compute_100a,code=sm_100a -lcudart_static -ldl -lrt -lpthread
echo built $PWD/libgsage_b200_electprod.so
