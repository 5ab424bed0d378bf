#!/bin/bash
# round 2, GPU call E: fused kernel v3 (unconditional loads), LSTM backward + fallback training, ncu of the attention kernel, RNG lane shapes
O=gpurun_out/r02e; mkdir -p $O
export GSAGE_FUSED_LAYER=1
timeout 120 python profiles/bench_fused.py > $O/bench_fused.log 2>&1
echo "bench_fused rc=$?" >> $O/summary.txt
D=256 timeout 120 python profiles/bench_fused.py > $O/bench_fused_d256.log 2>&1
timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q --no-header -p no:cacheprovider > $O/test_gpu_fused.log 2>&1
echo "test_gpu_fused rc=$? $(tail -1 $O/test_gpu_fused.log)" >> $O/summary.txt
timeout 300 python bench.py --legs none --no-cpu-baseline --no-train --steps 100 > $O/reddit_fused.json 2> $O/reddit_fused.err
N=142080 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_mean_project -c 1 -o $O/fused_v3 python profiles/bench_fused.py > $O/ncu_fused.log 2>&1
unset GSAGE_FUSED_LAYER
for n in test_gpu_backward test_gpu_autograd; do
  timeout 600 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
for cfg in "16 512" "32 256" "64 128" "128 64"; do
  set -- $cfg
  GSAGE_RNG_LANES=$1 GSAGE_RNG_LANE_BLOCKS=$2 timeout 300 python bench.py --legs none --no-cpu-baseline --no-train --steps 100 > $O/rng_$1x$2.json 2> $O/rng_$1x$2.err
done
timeout 300 python bench.py --legs none --no-cpu-baseline --no-train --no-ahead --steps 100 > $O/reddit_noahead.json 2> $O/reddit_noahead.err
timeout 300 python bench.py --workload big10m --no-cpu-baseline --no-train --no-ahead --steps 50 > $O/big10m_noahead.json 2> $O/big10m_noahead.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_fused -c 1 -s 4 -o $O/attention python bench.py --workload plaw2m-attention --batch 8192 --no-train --no-cpu-baseline --steps 4 --warmup 1 > $O/ncu_att.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_pool_ws -c 1 -s 6 -o $O/pool python bench.py --workload pokec-maxpool --batch 8192 --no-train --no-cpu-baseline --steps 4 --warmup 1 > $O/ncu_pool.log 2>&1
cat $O/summary.txt; cat $O/bench_fused*.log | grep -v "^ "
