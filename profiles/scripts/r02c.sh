#!/bin/bash
# round 2, GPU call C: fused gather+mean+project kernel v2 (setmaxnreg, 16 producers, compile-time fanout) + the re-runs of call B's failures
O=gpurun_out/r02c; mkdir -p $O
for n in test_gpu_fused test_gpu_autograd test_gpu_model test_gpu_train_loop; do
  timeout 600 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
timeout 200 python profiles/bench_fused.py > $O/bench_fused.log 2>&1
S=25 N=16384 timeout 200 python profiles/bench_fused.py > $O/bench_fused_S25.log 2>&1
D=256 timeout 200 python profiles/bench_fused.py > $O/bench_fused_d256.log 2>&1
timeout 300 python bench.py --legs none --no-cpu-baseline --steps 100 > $O/reddit_fused.json 2> $O/reddit_fused.err
GSAGE_FUSED_LAYER=0 timeout 300 python bench.py --legs none --no-cpu-baseline --no-train --steps 100 > $O/reddit_unfused.json 2> $O/reddit_unfused.err
timeout 300 python bench.py --workload big10m --no-cpu-baseline --no-train --steps 50 > $O/big10m_fused.json 2> $O/big10m_fused.err
N=142080 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_mean_project -c 1 -o $O/fused_v2 python profiles/bench_fused.py > $O/ncu_fused.log 2>&1
cat $O/summary.txt; cat $O/bench_fused*.log
