#!/bin/bash
# BASELINE config 5: synthetic 256-ring 1M-point scans, detector ablation — all three detectors, then star / x_zero / z_zero
# one at a time (the other two switched off), same scans, batch 16. One bench line each -> gpurun_out/c5_<name>.json
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py --shape C5 --batch 16 --steps 50 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/c5_$name.json 2> gpurun_out/c5_$name.err; echo "C5 $name rc=$? $(python -c "import json; d=json.loads(open('gpurun_out/c5_$name.json').read().strip().splitlines()[-1]); print('scans/s', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'with_order', round(d.get('with_order',{}).get('value',0)), 'dom', d['roofline']['kernel'], round(d['roofline']['frac'],3))" 2>&1)"; }
run all
run star_only --no-xzero --no-zzero
run xzero_only --no-star --no-zzero
run zzero_only --no-star --no-xzero
