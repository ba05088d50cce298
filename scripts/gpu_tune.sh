#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/tune.log
for cfg in "4 0" "4 1" "4 2" "1 0" "2 0" "8 0"; do
  set -- $cfg
  URF_TUNE_A=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --groups $1 2>/dev/null | sed "s/^/groups=$1 tune=$2 /" >> gpurun_out/tune.log
done
cat gpurun_out/tune.log | cut -c1-400
