#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests, smoke, sanitizer, short bench, ncu launch list. Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
if [ "$1" != "nosan" ]; then
echo "== sanitizer"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/sanitizer.log
fi
echo "== latency"; timeout 600 python scripts/latency.py > gpurun_out/latency.json 2> gpurun_out/latency.err; cut -c1-1500 gpurun_out/latency.json; tail -2 gpurun_out/latency.err
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --groups 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_bench.log
