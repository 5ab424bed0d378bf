#!/bin/bash
# round 2, GPU call A: first run of the fused gather+mean+project kernel; baselines for the round
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
GSAGE_TEST_EXPERIMENTAL=1 timeout 240 python -m pytest tests/test_gpu_experimental.py -m gpu -x -q > $O/exp_tests.log 2>&1
echo "exp tests rc=$?" >> $O/exp_tests.log
GSAGE_FUSED_LAYER=1 timeout 200 python profiles/bench_fused.py > $O/bench_fused.log 2>&1
echo "rc=$?" >> $O/bench_fused.log
timeout 300 python bench.py --no-train --no-cpu-baseline --steps 100 > $O/reddit_base.json 2> $O/reddit_base.err
GSAGE_FUSED_LAYER=1 timeout 300 python bench.py --no-train --no-cpu-baseline --steps 100 > $O/reddit_fused.json 2> $O/reddit_fused.err
for B in 8192 32768 131072; do
  timeout 300 python bench.py --workload pokec-mean --batch $B --no-train --no-cpu-baseline --steps 50 > $O/pokec_mean_B$B.json 2> $O/pokec_mean_B$B.err
done
GSAGE_FUSED_LAYER=1 N=142080 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_mean_project -c 1 -o $O/fused python profiles/bench_fused.py > $O/ncu_fused.log 2>&1
ls -la $O
tail -3 $O/exp_tests.log; cat $O/bench_fused.log
