#!/bin/bash
# round 2, GPU call AE (2 GPUs): multi-GPU parity tests + the default bench under torchrun
O=gpurun_out/r02ae; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider -x -s > $O/test_gpu_multi.log 2>&1
echo "test_gpu_multi rc=$? $(tail -1 $O/test_gpu_multi.log)" >> $O/summary.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err
echo "bench2 rc=$?" >> $O/summary.txt
cat $O/summary.txt; grep -i "collective\|peer\|nccl" $O/test_gpu_multi.log | head -5
