#!/bin/bash
# round 2, GPU call AL: --set full captures of the dominant projection launch (reddit) and of the attention kernel on the final build
O=gpurun_out/r02al; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_ws_umma_kernel -c 1 -s 3 -o $O/project python bench.py --legs none --no-train --no-cpu-baseline --no-ahead --steps 3 --warmup 1 > $O/ncu_project.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_fused_kernel -c 1 -s 4 -o $O/attention python bench.py --workload plaw2m-attention --legs none --no-train --no-cpu-baseline --no-ahead --steps 3 --warmup 1 > $O/ncu_attention.log 2>&1
ls -la $O
