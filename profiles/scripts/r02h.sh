#!/bin/bash
# round 2, GPU call H: attention kernel with per-parent LSU producers (setmaxnreg), 3 x TF32 blocks for the exact pool / attention models
O=gpurun_out/r02h; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -k "attention" > $O/ops_attention.log 2>&1
echo "ops_attention rc=$? $(tail -1 $O/ops_attention.log)" >> $O/summary.txt
for n in test_gpu_ops test_gpu_model test_gpu_backward test_gpu_autograd test_gpu_train_loop test_gpu_fused; do
  timeout 600 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
timeout 300 python bench.py --workload plaw2m-attention --no-cpu-baseline --steps 60 > $O/att.json 2> $O/att.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_fused -c 1 -s 4 -o $O/attention4 python bench.py --workload plaw2m-attention --batch 8192 --no-train --no-cpu-baseline --steps 4 --warmup 1 > $O/ncu_att.log 2>&1
timeout 900 python bench.py --steps 100 > $O/bench.json 2> $O/bench.err
cat $O/summary.txt
