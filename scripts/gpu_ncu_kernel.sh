#!/bin/bash
# Runs on the GPU box (under gpurun): one `ncu --set full` capture with source counters of ONE kernel of a batch-32,
# single-stream bench step.  usage: gpu_ncu_kernel.sh <kernel-name-regex> <tag> [batch]
mkdir -p gpurun_out
K=${1:-k_ring_detect}; TAG=${2:-cap}; B=${3:-32}
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$K" --launch-skip 3 -c 1 -f -o gpurun_out/prof_$TAG \
  python bench.py --steps 1 --warmup 3 --batch $B --groups 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$TAG.log
