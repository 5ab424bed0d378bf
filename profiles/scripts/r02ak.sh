#!/bin/bash
# round 2, GPU call AK: the LSTM aggregator's line on the final build (forward; B = 2048) and its launch list
O=gpurun_out/r02ak; mkdir -p $O
timeout 300 python bench.py --workload reddit-lstm --batch 2048 --legs none --no-cpu-baseline --steps 30 > $O/bench_lstm.json 2> $O/bench_lstm.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_lstm.csv python bench.py --workload reddit-lstm --batch 2048 --legs none --no-cpu-baseline --no-train --steps 2 --warmup 3 > $O/launches_lstm.log 2>&1
tail -c 600 $O/bench_lstm.json
