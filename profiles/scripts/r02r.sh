#!/bin/bash
# round 2, GPU call R: elect.sync issue loops in every tcgen05 / TMA kernel -- tests, pool micro-benchmark with cycle counters, bench
O=gpurun_out/r02r; mkdir -p $O
T=$PWD/pytorch_graphsage_b200/libgsage_b200_timing.so
for n in test_gpu_ops test_gpu_model test_gpu_backward test_gpu_fused test_gpu_autograd test_gpu_train_loop; do
  timeout 300 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider -x > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
{
python profiles/bench_pool.py
GSAGE_POOL_PIPE=0 python profiles/bench_pool.py
GSAGE_B200_LIB=$T GSAGE_POOL_PIPE=0 python profiles/bench_pool.py 2>&1 | tail -2
GSAGE_B200_LIB=$T python profiles/bench_pool.py 2>&1 | tail -2
python profiles/bench_fused.py
} > $O/micro.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --steps 100 > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; cat $O/micro.txt
