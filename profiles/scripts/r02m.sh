#!/bin/bash
# round 2, GPU call M: attention v6 (cp.async producers per buffer, u precomputed per parent)
O=gpurun_out/r02m; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -k "attention" > $O/ops_attention.log 2>&1
echo "ops_attention rc=$? $(tail -1 $O/ops_attention.log)" >> $O/summary.txt
GSAGE_ATT_DEBUG=1 timeout 300 python bench.py --workload plaw2m-attention --no-cpu-baseline --no-train --steps 5 --warmup 3 > $O/att_dbg.json 2> $O/att_dbg.err
grep "\[att\] n=409600" $O/att_dbg.err | tail -2 >> $O/summary.txt
timeout 300 python bench.py --workload plaw2m-attention --no-cpu-baseline --steps 60 > $O/att.json 2> $O/att.err
for n in test_gpu_model test_gpu_backward; do
  timeout 600 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
cat $O/summary.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02m/att.json')); r=d['roofline']
print('att ms/step %.3f launch %.3f frac %.3f train %s' % (d['ms_per_step'], r['avg_launch_ms'], r['frac'], d.get('train',{}).get('ms_per_step')), {k:round(v,3) for k,v in d['breakdown_ms_per_step'].items() if v})
PY
