mkdir -p gpurun_out
cat > /tmp/var_test.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from oracle.pyoracle import PortOracle
from urban_road_filter_b200 import api, make_params, FULL_ROI
from urban_road_filter_b200.synth import SHAPES, make_scan, random_cloud
port = PortOracle()
bad = 0
cases = [("C1",0,"column","full"),("C2",2,"column","full"),("C2",3,"ring","full"),("C3",4,"column","full"),("C2",6,"column","def")]
exp = {}
for mk in (2, 0):
    for shape, seed, order, roi in cases:
        sh = SHAPES[shape]
        pts = make_scan(shape, seed, order=order)
        prm = make_params(channels=sh.channels, interval=sh.interval, **(FULL_ROI if roi == "full" else {}))
        key = (shape, seed, order, roi)
        if key not in exp: exp[key] = port.run(pts, prm)
        o = exp[key]
        det = api.Detector(max_points=pts.shape[0], max_batch=1, params=prm)
        det.set_option(9, mk)
        r = det.filtered(pts)
        ok = np.array_equal(r.label, o.label) and np.array_equal(r.order, o.order) and r.n_vert == o.n_vert and np.array_equal(r.vert, o.vert)
        if not ok: print("MISMATCH", mk, key, int((r.label != o.label).sum()), r.n_vert, o.n_vert)
        bad += not ok
        det.close()
print("VARIANTS", "ALL OK" if not bad else f"{bad} FAILED")
PY
timeout 900 python /tmp/var_test.py 2>&1 | tail -6
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/bench_r02_v11h.json 2> gpurun_out/bench_r02_v11h.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_v11h.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'lean', round(d['e2e_lean']['value']), 'with_order', round(d['with_order']['value']), round(d['with_order']['ms_per_step'],4), 'e2e_order', round(d['with_order']['e2e']['value']))
print({k: round(v,4) for k,v in d['roofline']['kernel_ms_per_step'].items()})"; tail -3 gpurun_out/bench_r02_v11h.err
echo "== configs"; bash scripts/gpu_configs.sh --no-cpu-baseline
echo "== C5 ablation"; bash scripts/gpu_ablation_c5.sh
