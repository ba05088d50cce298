#!/bin/bash
# Runs on a multi-GPU box (gpurun --gpus N): GPU tests, then the bench at N=1 and N=$1 the way the driver launches it.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench N=1"; timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_n1.log; tail -3 gpurun_out/bench_n1.err
echo "== bench N=$N (torchrun)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_n$N.log; tail -5 gpurun_out/bench_n$N.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_ref.log; tail -3 gpurun_out/bench_ref.err
