#!/bin/bash
# round 2, GPU call AJ: the default bench line and the reference arm of the final build
O=gpurun_out/r02aj; mkdir -p $O
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench rc=$?" >> $O/summary.txt
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 200 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider > $O/tests.log 2>&1
echo "tests rc=$? $(tail -1 $O/tests.log)" >> $O/summary.txt
cat $O/summary.txt
