mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== initcheck smoke"; timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/san_initcheck_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/san_initcheck_smoke.log
echo "== bench (driver-style)"; timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02_v11k.json 2> gpurun_out/bench_r02_v11k.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_v11k.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'lean', round(d['e2e_lean']['value']), 'with_order', round(d['with_order']['value']), d['config'], d['roofline']['kernel'], round(d['roofline']['frac'],4))"; tail -2 gpurun_out/bench_r02_v11k.err
echo "== reference arm (driver-style)"; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_r02_v11k.json 2> gpurun_out/bench_ref_r02_v11k.err; echo "rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_ref_r02_v11k.json').read().strip().splitlines()[-1]); print(d['value'], d['config'], d['ms_per_step'])"
