#!/bin/bash
# round 2, GPU call AH: persistent tcgen05 kernels with 2 / 4 CTAs per SM-slot (smaller static tile shares) under the sample-ahead pipeline
O=gpurun_out/r02ah; mkdir -p $O
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --workload pokec-maxpool --legs none --no-cpu-baseline --no-train --steps 100 > $O/maxpool_$name.json 2> $O/maxpool_$name.err
  env "$@" timeout 200 python bench.py --legs plaw2m-attention --no-cpu-baseline --no-train --steps 100 > $O/reddit_att_$name.json 2> $O/reddit_att_$name.err
}
run w1 GSAGE_POOL_WAVES=1
run w2 GSAGE_POOL_WAVES=2 GSAGE_WS_WAVES=2 GSAGE_ATT_WAVES=2
run w4 GSAGE_POOL_WAVES=4 GSAGE_WS_WAVES=4 GSAGE_ATT_WAVES=4
ls $O
