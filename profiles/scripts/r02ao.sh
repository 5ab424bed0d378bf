#!/bin/bash
# round 2, GPU call AO: last check of the final binary (smoke, model + sampler + train-loop tests, a short headline bench)
O=gpurun_out/r02ao; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $O/smoke.log)" >> $O/summary.txt
timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_sampler.py tests/test_gpu_train_loop.py tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider > $O/tests.log 2>&1
echo "tests rc=$? $(tail -1 $O/tests.log)" >> $O/summary.txt
timeout 300 python bench.py --legs none --no-cpu-baseline --steps 100 > $O/bench_reddit.json 2> $O/bench_reddit.err
echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt
