#!/bin/bash
# every config shape through bench.py (HBM-resident + e2e), CPU reference beside it where it fits.
EXTRA=${1:-}   # e.g. --no-cpu-baseline to skip the CPU reference legs
mkdir -p gpurun_out
run() { shape=$1; batch=$2; extra=$3; timeout 900 python bench.py --shape $shape --batch $batch --steps 50 --warmup 3 $extra > gpurun_out/cfg_$shape.json 2> gpurun_out/cfg_$shape.err; echo "$shape rc=$? $(python -c "import json; d=json.loads(open('gpurun_out/cfg_$shape.json').read().strip().splitlines()[-1]); cb=d.get('cpu_baseline',{}); print('scans/s', round(d['value']), 'Mpts/s', round(d.get('mpoints_per_sec', d['config'].get('mpoints_per_sec', 0))), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), 'pipeline_frac', round(d['roofline']['pipeline_frac'],4), 'dom', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'cpu', round(cb.get('value',0),1), cb.get('cores'))" 2>&1)"; }
run C1 512 "$EXTRA"
run C3 256 "$EXTRA"
run C4 64 "$EXTRA"
run C5 16 "--no-cpu-baseline"
