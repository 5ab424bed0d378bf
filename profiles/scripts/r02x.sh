#!/bin/bash
# round 2, GPU call X: 64-column pool kernel with one full / empty barrier pair per (buffer, hidden block)
O=gpurun_out/r02x; mkdir -p $O
T=$PWD/pytorch_graphsage_b200/libgsage_b200_timing.so
GSAGE_NO_POOL_N128=1 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -x -k "pooled" > $O/test_pool_old.log 2>&1
echo "test_pool(64-col kernel) rc=$? $(tail -1 $O/test_pool_old.log)" >> $O/summary.txt
for n in test_gpu_model test_gpu_backward; do
  GSAGE_NO_POOL_N128=1 timeout 300 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider -x > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
{
GSAGE_NO_POOL_N128=1 timeout 120 python profiles/bench_pool.py
GSAGE_NO_POOL_N128=1 SEQ=1 ROWS=4096000 timeout 120 python profiles/bench_pool.py
GSAGE_NO_POOL_N128=1 S=25 N=163840 timeout 120 python profiles/bench_pool.py
GSAGE_NO_POOL_N128=1 D=256 ROWS=425984 timeout 120 python profiles/bench_pool.py
GSAGE_NO_POOL_N128=1 GSAGE_B200_LIB=$T timeout 120 python profiles/bench_pool.py 2>&1 | tail -2
timeout 120 python profiles/bench_pool.py
} > $O/micro.txt 2>&1
cat $O/summary.txt; cat $O/micro.txt
