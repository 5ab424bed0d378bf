#!/bin/bash
# round 2, GPU call AG: the fused gather+mean+project kernel again, after the issue-loop fix and the 32-byte epilogue stores
O=gpurun_out/r02ag; mkdir -p $O
{
GSAGE_FUSED_LAYER=1 python profiles/bench_fused.py
GSAGE_FUSED_LAYER=1 D=256 python profiles/bench_fused.py
} > $O/fused.txt 2>&1
GSAGE_FUSED_LAYER=1 timeout 300 python bench.py --legs big10m --no-cpu-baseline --no-train --steps 100 > $O/bench_fused.json 2> $O/bench_fused.err
timeout 300 python bench.py --legs big10m --no-cpu-baseline --no-train --steps 100 > $O/bench_plain.json 2> $O/bench_plain.err
cat $O/fused.txt
