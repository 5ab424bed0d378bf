#!/bin/bash
# round 2, GPU call B: whole GPU suite file by file (a poisoned CUDA context must not take the other files with it) + the default bench
O=gpurun_out/r02b; mkdir -p $O
for f in tests/test_gpu_*.py; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q -x --no-header -p no:cacheprovider > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt
