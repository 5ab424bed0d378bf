#!/bin/bash
# round 2, GPU call J (2 GPUs): multi-GPU parity tests + the bench under torchrun with the peer-memory all-reduce
O=gpurun_out/r02j; mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -p no:cacheprovider -s > $O/test_gpu_multi.log 2>&1
echo "test_gpu_multi rc=$? $(tail -1 $O/test_gpu_multi.log)" >> $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 3 --no-cpu-baseline --legs big10m:16384:30,plaw2m-attention:16384:30 > $O/bench_2gpu.json 2> $O/bench_2gpu.err
echo "bench_2gpu rc=$?" >> $O/summary.txt
GSAGE_SYMM_ALLREDUCE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 60 --warmup 3 --no-cpu-baseline --legs big10m:16384:30 > $O/bench_2gpu_nccl.json 2> $O/bench_2gpu_nccl.err
echo "bench_2gpu_nccl rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 60 --no-cpu-baseline --legs big10m:16384:30 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
cat $O/summary.txt; tail -5 $O/test_gpu_multi.log; tail -3 $O/bench_2gpu.err
