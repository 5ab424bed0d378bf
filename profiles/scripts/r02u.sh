#!/bin/bash
# round 2, GPU call U: attention kernel, split of a tile's rows between TMA gather4 and cp.async (sweep)
O=gpurun_out/r02u; mkdir -p $O
{
for r in 128 96 64 32 0; do GSAGE_ATT_TMA_ROWS=$r python profiles/bench_attention.py; done
for r in 128 64; do GSAGE_ATT_TMA_ROWS=$r D=602 ROWS=232966 python profiles/bench_attention.py; done
for r in 128 64; do GSAGE_ATT_TMA_ROWS=$r D=64 ROWS=1632803 python profiles/bench_attention.py; done
} > $O/micro.txt 2>&1
timeout 200 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -x -k attention > $O/test_att.log 2>&1; tail -1 $O/test_att.log
cat $O/micro.txt
