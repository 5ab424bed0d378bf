#!/bin/bash
# round 2, GPU call P: TMEM read micro-benchmark; pool kernel with the software-pipelined epilogue; launch lists of the train steps
O=gpurun_out/r02p; mkdir -p $O
timeout 120 profiles/build/tmem_read_bw > $O/tmem_read_bw.txt 2>&1
for n in test_gpu_ops test_gpu_model test_gpu_backward; do
  timeout 300 python -m pytest tests/$n.py -m gpu -q --no-header -p no:cacheprovider > $O/$n.log 2>&1
  echo "$n rc=$? $(tail -1 $O/$n.log)" >> $O/summary.txt
done
timeout 300 python bench.py --workload pokec-maxpool --legs none --no-cpu-baseline --no-train --steps 100 > $O/maxpool_pipe.json 2> $O/maxpool_pipe.err
GSAGE_POOL_PIPE=0 timeout 300 python bench.py --workload pokec-maxpool --legs none --no-cpu-baseline --no-train --steps 100 > $O/maxpool_nopipe.json 2> $O/maxpool_nopipe.err
for w in pokec-mean pokec-maxpool plaw2m-attention; do
  B=16384; [ $w = pokec-mean ] && B=32768
  GSAGE_BENCH_PROFILE_TRAIN=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train_launches_$w.csv \
    python bench.py --workload $w --batch $B --legs none --no-cpu-baseline --steps 3 --warmup 3 > $O/train_launches_$w.log 2>&1
  echo "launches $w rc=$?" >> $O/summary.txt
done
cat $O/summary.txt; cat $O/tmem_read_bw.txt
