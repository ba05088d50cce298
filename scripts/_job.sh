bash scripts/gpu_scale_mq.sh
