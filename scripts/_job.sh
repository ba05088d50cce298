mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline "$@" > gpurun_out/v11q_$name.json 2> gpurun_out/v11q_$name.err; echo "$name rc=$? $(python -c "import json; d=json.loads(open('gpurun_out/v11q_$name.json').read().strip().splitlines()[-1]); print('scans/s', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'with_order', round(d.get('with_order',{}).get('value',0)), 'dom', d['roofline']['kernel'], round(d['roofline']['frac'],3))" 2>&1)"; }
run C1 --shape C1 --batch 512 --steps 50
run C3 --shape C3 --batch 256 --steps 50
run C4 --shape C4 --batch 64 --steps 50
run C5_all --shape C5 --batch 16 --steps 50
run C5_star_only --shape C5 --batch 16 --steps 50 --no-xzero --no-zzero
run C5_xzero_only --shape C5 --batch 16 --steps 50 --no-star --no-zzero
run C5_zzero_only --shape C5 --batch 16 --steps 50 --no-star --no-xzero
