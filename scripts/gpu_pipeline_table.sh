#!/bin/bash
# Runs on the GPU box (under gpurun): whole-pipeline ncu table of ONE batch step on one stream (every kernel of the step):
# duration, DRAM bytes read/written, L2->L1 bytes, warp instructions, issue %, occupancy, the main stall reasons.
# usage: gpu_pipeline_table.sh <tag> [batch] [extra bench args]   -> gpurun_out/pipeline_<tag>.csv (+ a live bench line)
mkdir -p gpurun_out
TAG=${1:-cur}; B=${2:-128}; shift; shift
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
M=$M,launch__registers_per_thread,launch__grid_size,launch__block_size
M=$M,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
# warm-up = 3 steps; the 4th (timed) step is captured: skip the launches of 3 steps, then take one step's worth
timeout 1200 ncu --metrics $M --clock-control none --launch-skip ${SKIP:-72} -c ${COUNT:-24} --csv --log-file gpurun_out/pipeline_$TAG.csv \
  python bench.py --steps 1 --warmup 3 --batch $B --groups 1 --no-cpu-baseline --no-e2e "$@" > gpurun_out/pipeline_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/pipeline_$TAG.log
