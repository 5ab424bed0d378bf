#!/bin/bash
# round 2, GPU call N (8 GPUs): the bench under torchrun at N = 8 -- peer-memory all-reduce vs NCCL, sharded parity on every rank
O=gpurun_out/r02n; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 40 --warmup 3 --no-cpu-baseline --legs big10m:16384:20 > $O/bench_8gpu.json 2> $O/bench_8gpu.err
echo "bench_8gpu rc=$?" >> $O/summary.txt
GSAGE_SYMM_ALLREDUCE=0 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 40 --warmup 3 --no-cpu-baseline --legs big10m:16384:20 > $O/bench_8gpu_nccl.json 2> $O/bench_8gpu_nccl.err
echo "bench_8gpu_nccl rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -3 $O/bench_8gpu.err
