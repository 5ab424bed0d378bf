#!/bin/bash
# round 2, GPU call T: pool producers with staged ids; attention with TMA + cp.async producers (split sweep); vector-RED embedding scatter
O=gpurun_out/r02t; mkdir -p $O
T=$PWD/pytorch_graphsage_b200/libgsage_b200_timing.so
for n in test_gpu_ops test_gpu_model test_gpu_backward test_gpu_fused test_gpu_autograd test_gpu_train_loop; do
  timeout 300 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider -x > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
{
python profiles/bench_pool.py
GSAGE_B200_LIB=$T python profiles/bench_pool.py 2>&1 | tail -2
for r in 128 96 64 32 0; do GSAGE_ATT_TMA_ROWS=$r python profiles/bench_attention.py; done
for r in 128 64 0; do GSAGE_ATT_TMA_ROWS=$r D=602 ROWS=232966 python profiles/bench_attention.py; done
} > $O/micro.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --steps 100 > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt; cat $O/micro.txt
