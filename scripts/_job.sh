mkdir -p gpurun_out
SW="4,0,0,0;4,0,1,0;8,0,1,0;16,0,1,0;8,4,1,0;8,4,1,1;8,2,1,1;8,1,1,1;16,1,1,1;16,2,1,1;16,4,1,1;4,4,1,1;4,8,1,1;8,8,1,1;2,16,1,1;8,2,0,1;1,0,0,0"
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sweep "$SW" > gpurun_out/sweep1.log 2> gpurun_out/sweep1.err; echo "sweep rc=$?"; cat gpurun_out/sweep1.log; tail -3 gpurun_out/sweep1.err
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
