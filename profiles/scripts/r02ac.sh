#!/bin/bash
# round 2, GPU call AC: 32-byte epilogue stores in every projection kernel + twelve epilogue warps outside 3 x TF32 mode; tests + bench
O=gpurun_out/r02ac; mkdir -p $O
for n in test_gpu_ops test_gpu_model test_gpu_backward test_gpu_fused test_gpu_autograd test_gpu_train_loop; do
  timeout 300 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider -x > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
{
MODE=x3 python profiles/bench_linear.py
MODE=tf32 python profiles/bench_linear.py
MODE=bf16 python profiles/bench_linear.py
MODE=bf16 D=256 ROWS=10000000 N=425984 python profiles/bench_linear.py
MODE=bf16 D=602 ROWS=232966 N=425984 python profiles/bench_linear.py
} > $O/linear.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline --steps 100 > $O/bench.json 2> $O/bench.err
echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt $O/linear.txt
