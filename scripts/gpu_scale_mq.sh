#!/bin/bash
# Runs on an 8-GPU box (gpurun --gpus 8): the bench the way the driver launches it at N = 8, 4, 2 (torchrun, one rank per
# GPU), then ONE process streaming OS2-128 scans into all GPUs through urf_mq (BASELINE config 4). -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt; nproc >> gpurun_out/gpus.txt
for N in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/scale_r02_n$N.json 2> gpurun_out/scale_r02_n$N.err
  echo "N=$N rc=$? $(python -c "import json; d=json.loads(open('gpurun_out/scale_r02_n$N.json').read().strip().splitlines()[-1]); print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'lean', round(d['e2e_lean']['value']), 'with_order', round(d['with_order']['value']))" 2>&1)"
done
echo "== mq, 8 devices"; timeout 900 python scripts/bench_mq.py --gpus 8 --scans 8000 --producers 1,8 --slots 96 --max-batch 64 > gpurun_out/mq_r02_8gpu.jsonl 2> gpurun_out/mq_r02.err; echo "mq rc=$?"; cat gpurun_out/mq_r02_8gpu.jsonl; tail -3 gpurun_out/mq_r02.err
echo "== mq, 4 and 1 devices"; for G in 4 1; do timeout 600 python scripts/bench_mq.py --gpus $G --scans 3000 --producers 4 --slots 96 --max-batch 64 >> gpurun_out/mq_r02_fewer.jsonl 2>> gpurun_out/mq_r02.err; done; cat gpurun_out/mq_r02_fewer.jsonl
