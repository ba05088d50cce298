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

# The ROS-shaped path: raw 48-byte PointCloud2 records in, the four published clouds out, both ends on
# the device, cfg-default ROI (what the node runs with) and full ROI; compared with labels + emission order + host packing.
def records48(pts):
    rec = np.zeros((pts.shape[0], 48), np.uint8)
    rec[:, 0:12] = pts[:, 0:3].copy().view(np.uint8)
    rec[:, 16:20] = pts[:, 3:4].copy().view(np.uint8)
    return torch.from_numpy(rec.reshape(-1)).pin_memory().numpy()

def host_pack(pts, r):
    lab, order = r.label, r.order
    sel = lambda ids: np.concatenate([pts[ids, :3], np.ones((len(ids), 1), np.float32), pts[ids, 3:4], np.zeros((len(ids), 3), np.float32)], 1)
    return [sel(order[lab[order] == 1]), sel(order[lab[order] == 2]), sel(np.flatnonzero(lab >= 0)), sel(order[r.ring_start[10]:r.ring_start[11]] if len(r.ring_start) > 11 else order[:0])]

packed = {}
for cfg in ("C2", "C4"):
    sh = SHAPES[cfg]
    pts = make_scan(cfg, 0)
    raw = records48(pts)
    n = pts.shape[0]
    det = api.Detector(max_points=n, max_batch=1)
    for roi_name, roi in (("default_roi", {}), ("full_roi", FULL_ROI)):
        det.set_params(make_params(channels=sh.channels, interval=sh.interval, **roi))
        for mode in ("device_pack", "labels_then_host_pack"):
            t = []
            for it in range(25):
                t0 = time.perf_counter()
                if mode == "device_pack":
                    r, cl = det.filtered_cloud2_packed(raw, n, 48, 0, 4, 8, 16)
                else:
                    r = det.filtered_cloud2(raw, n, 48, 0, 4, 8)
                    cl = host_pack(pts, r)
                if it >= 5:
                    t.append(time.perf_counter() - t0)
            packed[f"{cfg}_{roi_name}_{mode}"] = {"wall_ms_median": 1e3 * float(np.median(t)), "n_roi": int(r.n_roi), "n_road": int(r.n_road), "n_curb": int(r.n_curb)}
    det.close()
print(json.dumps({"cloud2_packed": packed}))
