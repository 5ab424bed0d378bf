#!/bin/bash
# round 2, GPU call AB: projection epilogue with 32-byte stores
O=gpurun_out/r02ab; mkdir -p $O
T=$PWD/pytorch_graphsage_b200/libgsage_b200_timing.so
{
MODE=x3 python profiles/bench_linear.py
MODE=tf32 python profiles/bench_linear.py
MODE=bf16 python profiles/bench_linear.py
MODE=bf16 D=256 ROWS=10000000 N=425984 python profiles/bench_linear.py
MODE=bf16 D=602 ROWS=232966 N=425984 python profiles/bench_linear.py
GSAGE_B200_LIB=$T MODE=bf16 python profiles/bench_linear.py 2>&1 | tail -2
GSAGE_B200_LIB=$T MODE=tf32 python profiles/bench_linear.py 2>&1 | tail -2
} > $O/linear.txt 2>&1
for n in test_gpu_ops test_gpu_model; do
  timeout 300 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider -x > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)"
done
cat $O/linear.txt
