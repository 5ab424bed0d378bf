#!/bin/bash
# round 2, GPU call V: the 128-column pool kernel (linear_pool_n128_umma.cu): parity, then LSU vs TMA fill vs rows in place vs the 64-column kernel
O=gpurun_out/r02v; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -x -k "pool" > $O/test_pool.log 2>&1
echo "test_pool rc=$? $(tail -1 $O/test_pool.log)" >> $O/summary.txt
for n in test_gpu_model test_gpu_backward; do
  timeout 300 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider -x > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
{
GSAGE_POOL_FILL=lsu timeout 120 python profiles/bench_pool.py
GSAGE_POOL_FILL=tma timeout 120 python profiles/bench_pool.py
SEQ=1 ROWS=4096000 timeout 120 python profiles/bench_pool.py
GSAGE_NO_POOL_N128=1 timeout 120 python profiles/bench_pool.py
S=25 N=163840 GSAGE_POOL_FILL=lsu timeout 120 python profiles/bench_pool.py
S=25 N=163840 GSAGE_NO_POOL_N128=1 timeout 120 python profiles/bench_pool.py
} > $O/micro.txt 2>&1
timeout 300 python bench.py --workload pokec-maxpool --legs none --no-cpu-baseline --steps 100 > $O/maxpool.json 2> $O/maxpool.err
cat $O/summary.txt; cat $O/micro.txt; tail -5 $O/test_pool.log
