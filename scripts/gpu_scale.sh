#!/bin/bash
# Scaling check the way the driver runs it: N = 1, 2, 4, 8 back to back (torchrun for N > 1).
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "N=$N rc=$? $(python -c "import json,sys; d=json.loads(open('gpurun_out/scale_n$N.json').read().strip().splitlines()[-1]); print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))" 2>&1)"
  grep "rank" gpurun_out/scale_n$N.err | head -8
done
