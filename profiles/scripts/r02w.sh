#!/bin/bash
# round 2, GPU call W: tcgen05.mma rate micro-benchmark; pool kernel parity again (TMA fill, ids outside the table)
O=gpurun_out/r02w; mkdir -p $O
timeout 120 profiles/build/umma_rate > $O/umma_rate.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -x -k "pool" > $O/test_pool.log 2>&1
echo "test_pool rc=$? $(tail -1 $O/test_pool.log)"
cat $O/umma_rate.txt
