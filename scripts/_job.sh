mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_quick.log; tail -3 gpurun_out/bench_quick.err
SKIP=42 COUNT=14 bash scripts/gpu_pipeline_table.sh r02_v11b 128
