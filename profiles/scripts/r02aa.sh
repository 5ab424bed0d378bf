#!/bin/bash
# round 2, GPU call AA: where do the projection kernel's cycles go?  (3 x TF32 on the pokec shape; bf16 on pokec / big10m / reddit shapes)
O=gpurun_out/r02aa; mkdir -p $O
T=$PWD/pytorch_graphsage_b200/libgsage_b200_timing.so
{
MODE=x3 python profiles/bench_linear.py
MODE=tf32 python profiles/bench_linear.py
MODE=bf16 python profiles/bench_linear.py
MODE=bf16 D=256 ROWS=10000000 N=425984 python profiles/bench_linear.py
MODE=bf16 D=602 ROWS=232966 N=425984 python profiles/bench_linear.py
echo "== timing build"
GSAGE_B200_LIB=$T MODE=x3 python profiles/bench_linear.py 2>&1 | tail -2
GSAGE_B200_LIB=$T MODE=tf32 python profiles/bench_linear.py 2>&1 | tail -2
GSAGE_B200_LIB=$T MODE=bf16 python profiles/bench_linear.py 2>&1 | tail -2
GSAGE_B200_LIB=$T MODE=bf16 D=256 ROWS=10000000 N=425984 python profiles/bench_linear.py 2>&1 | tail -2
GSAGE_B200_LIB=$T MODE=bf16 D=602 ROWS=232966 N=425984 python profiles/bench_linear.py 2>&1 | tail -3
} > $O/linear.txt 2>&1
cat $O/linear.txt
