#!/bin/bash
# round 2, GPU call AN: final build (RNG draws of the next batch ahead of the gate for the mean / pool aggregators): the tests that touch the
# sampler pipeline, then the default bench line
O=gpurun_out/r02an; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_rng.py tests/test_gpu_model.py tests/test_gpu_train_loop.py tests/test_gpu_fullsize.py tests/test_gpu_backward.py -m gpu -q --no-header -p no:cacheprovider > $O/tests.log 2>&1
echo "tests rc=$? $(tail -1 $O/tests.log)" >> $O/summary.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench rc=$?" >> $O/summary.txt
cat $O/summary.txt
