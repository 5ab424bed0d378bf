#!/bin/bash
# round 2, GPU call AF: normalise + classifier as one launch (head_kernel); tests + bench
O=gpurun_out/r02af; mkdir -p $O
for n in test_gpu_model test_gpu_backward test_gpu_train_loop test_gpu_fullsize test_gpu_autograd; do
  timeout 300 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider -x > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
timeout 600 python bench.py --no-cpu-baseline --steps 100 > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt
