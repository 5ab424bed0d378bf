#!/bin/bash
# round 2, GPU call Q: where do the pool kernel's cycles go?  (in-kernel cycle counters, gathered vs in-place rows, pipelined vs not)
O=gpurun_out/r02q; mkdir -p $O
T=$PWD/pytorch_graphsage_b200/libgsage_b200_timing.so
{
echo "== product build"; python profiles/bench_pool.py
GSAGE_POOL_PIPE=0 python profiles/bench_pool.py
SEQ=1 ROWS=4096000 python profiles/bench_pool.py
SEQ=1 ROWS=4096000 GSAGE_POOL_PIPE=0 python profiles/bench_pool.py
echo "== timing build"
GSAGE_B200_LIB=$T GSAGE_POOL_PIPE=0 python profiles/bench_pool.py 2>&1 | tail -3
GSAGE_B200_LIB=$T python profiles/bench_pool.py 2>&1 | tail -3
GSAGE_B200_LIB=$T SEQ=1 ROWS=4096000 GSAGE_POOL_PIPE=0 python profiles/bench_pool.py 2>&1 | tail -3
} > $O/pool.txt 2>&1
cat $O/pool.txt
