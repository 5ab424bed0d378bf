#!/bin/bash
# round 2, GPU call I: attention kernel v5 (self-contained compute groups, one per tile buffer)
O=gpurun_out/r02i; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -k "attention" > $O/ops_attention.log 2>&1
echo "ops_attention rc=$? $(tail -1 $O/ops_attention.log)" >> $O/summary.txt
if ! grep -q "failed\|error" $O/ops_attention.log; then
  for n in test_gpu_model test_gpu_backward test_gpu_fused test_gpu_autograd; do
    timeout 600 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider > $O/$n.log 2>&1
    echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
  done
  timeout 300 python bench.py --workload plaw2m-attention --no-cpu-baseline --steps 60 > $O/att_g3.json 2> $O/att_g3.err
  GSAGE_ATT_GROUPS=2 timeout 300 python bench.py --workload plaw2m-attention --no-cpu-baseline --no-train --steps 60 > $O/att_g2.json 2> $O/att_g2.err
  GSAGE_ATT_GROUPS=1 timeout 300 python bench.py --workload plaw2m-attention --no-cpu-baseline --no-train --steps 60 > $O/att_g1.json 2> $O/att_g1.err
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_fused -c 1 -s 4 -o $O/attention5 python bench.py --workload plaw2m-attention --batch 8192 --no-train --no-cpu-baseline --steps 4 --warmup 1 > $O/ncu_att.log 2>&1
fi
cat $O/summary.txt
