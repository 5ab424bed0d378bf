#!/bin/bash
O=gpurun_out/r02l; mkdir -p $O
GSAGE_ATT_DEBUG=1 timeout 300 python bench.py --workload plaw2m-attention --no-cpu-baseline --no-train --steps 5 --warmup 3 > $O/att_dbg.json 2> $O/att_dbg.err
grep "\[att\]" $O/att_dbg.err | sort | uniq -c | sort -rn | head -20
