#!/bin/bash
# round 2, GPU call AI: final build (after the projection grid change) -- whole GPU suite, default bench, reference arm, smoke, evidence captures
O=gpurun_out/r02ai; mkdir -p $O
for f in tests/test_gpu_*.py; do
  n=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q --no-header -p no:cacheprovider > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench rc=$?" >> $O/summary.txt
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
echo "smoke rc=$?" >> $O/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_reddit.csv python bench.py --legs none --no-cpu-baseline --no-train --steps 4 --warmup 3 > $O/launches_reddit.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_pokec-maxpool.csv python bench.py --workload pokec-maxpool --legs none --no-cpu-baseline --no-train --steps 4 --warmup 3 > $O/launches_maxpool.log 2>&1
cat $O/summary.txt
