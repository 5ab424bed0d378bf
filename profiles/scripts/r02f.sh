#!/bin/bash
# round 2, GPU call F: attention kernel with the weighted sum on the tensor core; fused mean kernel at S=25 / d=256
O=gpurun_out/r02f; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -k "attention" > $O/ops_attention.log 2>&1
echo "ops_attention rc=$? $(tail -1 $O/ops_attention.log)" >> $O/summary.txt
if grep -q "passed" $O/ops_attention.log && ! grep -q "failed" $O/ops_attention.log; then
  for n in test_gpu_ops test_gpu_model test_gpu_backward; do
    timeout 600 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider > $O/$n.log 2>&1
    echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
  done
  timeout 300 python bench.py --workload plaw2m-attention --no-cpu-baseline --steps 60 > $O/att_nbuf3.json 2> $O/att_nbuf3.err
  GSAGE_ATT_NBUF=2 timeout 300 python bench.py --workload plaw2m-attention --no-cpu-baseline --no-train --steps 60 > $O/att_nbuf2.json 2> $O/att_nbuf2.err
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_fused -c 1 -s 4 -o $O/attention2 python bench.py --workload plaw2m-attention --batch 8192 --no-train --no-cpu-baseline --steps 4 --warmup 1 > $O/ncu_att.log 2>&1
fi
export GSAGE_FUSED_LAYER=1
S=25 N=16384 D=256 timeout 120 python profiles/bench_fused.py > $O/bench_fused_S25_d256.log 2>&1
S=25 N=16384 timeout 120 python profiles/bench_fused.py > $O/bench_fused_S25_d602.log 2>&1
S=10 D=64 timeout 120 python profiles/bench_fused.py > $O/bench_fused_S10_d64.log 2>&1
timeout 300 python bench.py --workload big10m --no-cpu-baseline --no-train --steps 50 > $O/big10m_fused.json 2> $O/big10m_fused.err
unset GSAGE_FUSED_LAYER
GSAGE_RNG_LANES=16 GSAGE_RNG_LANE_BLOCKS=512 timeout 300 python bench.py --workload big10m --no-cpu-baseline --no-train --steps 50 > $O/big10m_rng16.json 2> $O/big10m_rng16.err
cat $O/summary.txt; cat $O/bench_fused*.log | grep -v "^ "
