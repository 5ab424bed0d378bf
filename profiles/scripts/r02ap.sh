#!/bin/bash
# round 2, GPU call AP (2 GPUs): multi-GPU parity tests + the headline leg under torchrun on the final binary
O=gpurun_out/r02ap; mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider -x > $O/test_gpu_multi.log 2>&1
echo "test_gpu_multi rc=$? $(tail -1 $O/test_gpu_multi.log)" >> $O/summary.txt
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --legs none --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err
echo "bench2 rc=$?" >> $O/summary.txt
cat $O/summary.txt
