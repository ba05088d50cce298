"""Single-scan latency through the host-buffer C-ABI call (urf_process): wall clock per call, pinned input, labels only."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np, torch
from urban_road_filter_b200 import api, make_params, FULL_ROI
from urban_road_filter_b200.synth import make_scan, SHAPES

out = {}
for cfg in ("C1", "C2", "C4", "C5"):
    sh = SHAPES[cfg]
    pts = torch.from_numpy(make_scan(cfg, 0)).pin_memory().numpy()
    det = api.Detector(max_points=pts.shape[0], max_batch=1, params=make_params(channels=sh.channels, interval=sh.interval, **FULL_ROI))
    for graph in (0, 1):
        det.set_option(3, graph)
        for _ in range(5):
            det.filtered(pts, want_ring=False, want_order=False)
        t = []
        for _ in range(30):
            t0 = time.perf_counter(); det.filtered(pts, want_ring=False, want_order=False); t.append(time.perf_counter() - t0)
        out[f"{cfg}_graph{graph}"] = {"points": int(pts.shape[0]), "wall_ms_median": 1e3 * float(np.median(t)), "device_ms": det.last_device_ms()}
    det.close()
print(json.dumps(out))
