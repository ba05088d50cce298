mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt; nproc >> gpurun_out/gpus.txt
echo "== mq, 8 devices"; timeout 240 python scripts/bench_mq.py --gpus 8 --scans 8000 --producers 1,8 --slots 96 --max-batch 64 > gpurun_out/mq_r02_v11q_8gpu.jsonl 2> gpurun_out/mq_r02_v11q8.err; echo "mq rc=$?"; cat gpurun_out/mq_r02_v11q_8gpu.jsonl; tail -3 gpurun_out/mq_r02_v11q8.err
