#!/bin/bash
# round 2, GPU call K: attention v5 with backed-off producer polling
O=gpurun_out/r02k; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -k "attention" > $O/ops_attention.log 2>&1
echo "ops_attention rc=$? $(tail -1 $O/ops_attention.log)" >> $O/summary.txt
timeout 300 python bench.py --workload plaw2m-attention --no-cpu-baseline --no-train --steps 60 > $O/att_g3.json 2> $O/att_g3.err
timeout 300 python bench.py --workload pokec-mean --batch 32768 --no-cpu-baseline --no-train --steps 40 > $O/pokec_mean.json 2> $O/pokec_mean.err
timeout 300 python bench.py --workload pokec-maxpool-f32 --batch 4096 --no-cpu-baseline --no-train --steps 10 > $O/pool_f32.json 2> $O/pool_f32.err
cat $O/summary.txt
