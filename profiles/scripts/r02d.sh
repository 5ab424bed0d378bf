#!/bin/bash
# round 2, GPU call D: fused kernel v2 with a register budget that fits + attention kernel with W2 on the parent side
O=gpurun_out/r02d; mkdir -p $O
timeout 120 python profiles/bench_fused.py > $O/bench_fused.log 2>&1
echo "bench_fused rc=$?" >> $O/summary.txt
if grep -q "^fused" $O/bench_fused.log; then
  for n in test_gpu_fused test_gpu_model test_gpu_ops test_gpu_backward; do
    timeout 600 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider > $O/$n.log 2>&1
    echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
  done
  S=25 N=16384 timeout 200 python profiles/bench_fused.py > $O/bench_fused_S25.log 2>&1
  D=256 timeout 200 python profiles/bench_fused.py > $O/bench_fused_d256.log 2>&1
  timeout 300 python bench.py --legs plaw2m-attention:16384:40,big10m:16384:40 --no-cpu-baseline --steps 100 > $O/bench.json 2> $O/bench.err
  for cfg in "8 1024" "16 512" "32 256" "64 128"; do
    set -- $cfg
    GSAGE_RNG_LANES=$1 GSAGE_RNG_LANE_BLOCKS=$2 timeout 300 python bench.py --legs none --no-cpu-baseline --no-train --steps 100 > $O/rng_$1x$2.json 2> $O/rng_$1x$2.err
  done
  N=142080 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_mean_project -c 1 -o $O/fused_v2 python profiles/bench_fused.py > $O/ncu_fused.log 2>&1
else
  GSAGE_FUSED_LAYER=0 timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider > $O/unfused_tests.log 2>&1
  echo "unfused tests rc=$? $(tail -1 $O/unfused_tests.log)" >> $O/summary.txt
fi
cat $O/summary.txt; cat $O/bench_fused*.log | grep -v "^  "
