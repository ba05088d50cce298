#!/bin/bash
# Round-end evidence on one B200 (under gpurun): smoke, GPU tests, the default bench line (with the CPU reference legs),
# the reference arm, the ncu launch list and the whole-pipeline table of the same step, one full capture of the dominant
# kernel. Everything lands in gpurun_out/ with the tag given as $1.
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 1500 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_$TAG.json; tail -4 gpurun_out/bench_$TAG.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "rc=$?"; cut -c1-500 gpurun_out/bench_ref_$TAG.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --groups 1 --no-cpu-baseline --no-e2e --no-with-order > gpurun_out/ncu_launches.log 2>&1; echo "ncu rc=$?"
echo "== pipeline table"; SKIP=48 COUNT=16 bash scripts/gpu_pipeline_table.sh $TAG 128
echo "== full capture of the dominant kernel"; timeout 900 ncu --set full --import-source on --clock-control none -k regex:"${DOM:-k_points}" --launch-skip 3 -c 1 -f -o gpurun_out/prof_${TAG}_dominant python bench.py --steps 1 --warmup 3 --batch 128 --groups 1 --no-cpu-baseline --no-e2e --no-with-order > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
