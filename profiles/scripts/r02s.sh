#!/bin/bash
# round 2, GPU call S: elect.sync MMA issuers + multi-lane gather producers (A/B against elected-lane producers); lean pool epilogue
O=gpurun_out/r02s; mkdir -p $O
T=$PWD/pytorch_graphsage_b200/libgsage_b200_timing.so
V=$PWD/pytorch_graphsage_b200/libgsage_b200_electprod.so
for n in test_gpu_ops test_gpu_model test_gpu_backward test_gpu_fused; do
  timeout 300 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider -x > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
{
python profiles/bench_pool.py
GSAGE_B200_LIB=$T python profiles/bench_pool.py 2>&1 | tail -2
GSAGE_B200_LIB=$T GSAGE_POOL_DBG_SKIP=1 python profiles/bench_pool.py 2>&1 | tail -2
} > $O/micro.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --steps 100 > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" >> $O/summary.txt
for w in plaw2m-attention big10m; do
  GSAGE_B200_LIB=$V timeout 300 python bench.py --workload $w --legs none --no-cpu-baseline --no-train --steps 100 > $O/variant_$w.json 2> $O/variant_$w.err
done
cat $O/summary.txt; cat $O/micro.txt
