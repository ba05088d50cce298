mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=8 -k "near_first or radius or tie or golden or fallback or random_parameter or edge or batch" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-with-order "$@" > gpurun_out/v11p_$name.json 2> gpurun_out/v11p_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/v11p_$name.json | cut -c1-900; }
run C2 --steps 100
run C3 --shape C3 --batch 256 --steps 30
run C4 --shape C4 --batch 64 --steps 30
