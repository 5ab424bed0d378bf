#!/bin/bash
# round 2, GPU call AM: sample-ahead with the RNG draws ahead of the gate (GSAGE_AHEAD_SPLIT=1) on the final build
O=gpurun_out/r02am; mkdir -p $O
GSAGE_AHEAD_SPLIT=1 timeout 600 python bench.py --no-cpu-baseline --no-train --steps 100 > $O/bench_split.json 2> $O/bench_split.err
echo "rc=$?"
