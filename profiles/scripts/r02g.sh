#!/bin/bash
# round 2, GPU call G: attention kernel with LSU row producers; RNG jump with a shared prefix; gather_reduce at 4 CTAs per SM
O=gpurun_out/r02g; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -k "attention" > $O/ops_attention.log 2>&1
echo "ops_attention rc=$? $(tail -1 $O/ops_attention.log)" >> $O/summary.txt
for n in test_gpu_rng test_gpu_fullsize test_gpu_sampler test_gpu_ops test_gpu_model test_gpu_backward test_gpu_autograd; do
  timeout 600 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
timeout 600 python bench.py --no-cpu-baseline --steps 100 > $O/bench.json 2> $O/bench.err
GSAGE_RNG_LANES=16 GSAGE_RNG_LANE_BLOCKS=512 timeout 300 python bench.py --legs none --no-cpu-baseline --no-train --steps 100 > $O/reddit_rng16x512.json 2> $O/reddit_rng16x512.err
GSAGE_RNG_LANES=64 GSAGE_RNG_LANE_BLOCKS=128 timeout 300 python bench.py --legs none --no-cpu-baseline --no-train --steps 100 > $O/reddit_rng64x128.json 2> $O/reddit_rng64x128.err
timeout 300 python bench.py --legs none --no-cpu-baseline --no-train --steps 100 > $O/reddit_rng32x256.json 2> $O/reddit_rng32x256.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_reddit.csv python bench.py --legs none --no-cpu-baseline --no-train --steps 6 --warmup 3 > $O/launches_reddit.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_fused -c 1 -s 4 -o $O/attention3 python bench.py --workload plaw2m-attention --batch 8192 --no-train --no-cpu-baseline --steps 4 --warmup 1 > $O/ncu_att.log 2>&1
cat $O/summary.txt
