mkdir -p gpurun_out
cat > /tmp/var_test.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from oracle.pyoracle import PortOracle
from urban_road_filter_b200 import api, make_params, FULL_ROI
from urban_road_filter_b200.synth import SHAPES, make_scan, random_cloud
port = PortOracle()
bad = 0
cases = [("C1",0,"column","full"),("C1",1,"ring","def"),("C2",2,"column","full"),("C2",3,"ring","full"),("C4",5,"ring","full"),("C2",6,"column","def")]
exp = {}
for rd, mk in [(45,0),(46,1),(4,1),(8,1)]:
    for shape, seed, order, roi in cases:
        sh = SHAPES[shape]
        pts = make_scan(shape, seed, order=order)
        prm = make_params(channels=sh.channels, interval=sh.interval, **(FULL_ROI if roi == "full" else {}))
        key = (shape, seed, order, roi)
        if key not in exp: exp[key] = port.run(pts, prm)
        o = exp[key]
        det = api.Detector(max_points=pts.shape[0], max_batch=1, params=prm)
        det.set_option(8, rd); det.set_option(9, mk)
        r = det.filtered(pts)
        ok = np.array_equal(r.label, o.label) and np.array_equal(r.order, o.order) and r.n_vert == o.n_vert and np.array_equal(r.vert, o.vert)
        if not ok: print("MISMATCH", rd, mk, key, int((r.label != o.label).sum()), r.n_vert, o.n_vert)
        bad += not ok
        det.close()
print("VARIANTS", "ALL OK" if not bad else f"{bad} FAILED")
PY
timeout 900 python /tmp/var_test.py 2>&1 | tail -8
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
SW="4,0,0,0,4,0;4,0,0,0,45,0;4,0,0,0,46,0;4,0,0,0,4,1;4,0,0,0,45,1;1,0,0,0,4,1;4,0,1,0,4,1"
timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --sweep "$SW" > gpurun_out/sweep3.log 2> gpurun_out/sweep3.err; echo "sweep rc=$?"; cat gpurun_out/sweep3.log; tail -3 gpurun_out/sweep3.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-with-order --mk 1 > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_quick.log; tail -3 gpurun_out/bench_quick.err
SKIP=48 COUNT=16 bash scripts/gpu_pipeline_table.sh r02_v11e 128 --mk 1
