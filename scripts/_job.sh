mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=8 -k "order or ring_sort or sort or golden or packed or emission" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline "$@" > gpurun_out/v11m_$name.json 2> gpurun_out/v11m_$name.err; echo "$name rc=$? $(python -c "import json; d=json.loads(open('gpurun_out/v11m_$name.json').read().strip().splitlines()[-1]); print('scans/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'with_order', d.get('with_order'))" 2>&1)"; tail -3 gpurun_out/v11m_$name.err; }
run C2 --steps 100
run C3 --shape C3 --batch 256 --steps 30
