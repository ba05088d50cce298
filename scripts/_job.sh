# scratch entry point of `gpurun -- 'bash scripts/_job.sh'`; the round-end evidence run:
bash scripts/gpu_final.sh r02_v11q
