#!/bin/bash
# round 2, GPU call Z: projection kernel epilogue with stores transposed through shared memory; tests + default bench
O=gpurun_out/r02z; mkdir -p $O
for n in test_gpu_ops test_gpu_model test_gpu_backward test_gpu_fused test_gpu_autograd test_gpu_train_loop; do
  timeout 300 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider -x > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
timeout 600 python bench.py --no-cpu-baseline --steps 100 > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt
