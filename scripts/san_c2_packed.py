"""One C2 scan (full ROI, half of it flat so that some star sectors are refined) through urf_process_cloud2_packed:
exercises the near-first star sort, its refine/resume pass and the device-side cloud packing under compute-sanitizer."""
import sys
sys.path.insert(0, ".")
import numpy as np
from urban_road_filter_b200 import api, make_params, FULL_ROI
from urban_road_filter_b200.synth import make_scan

pts = make_scan("C2", 5)
pts[pts[:, 0] < 0, 2] = -1.8
n = pts.shape[0]
rec = np.zeros((n, 32), np.uint8)
rec[:, 0:12] = pts[:, 0:3].copy().view(np.uint8)
rec[:, 16:20] = pts[:, 3:4].copy().view(np.uint8)
det = api.Detector(max_points=n, max_batch=1, params=make_params(**FULL_ROI))
r, cl = det.filtered_cloud2_packed(rec.reshape(-1), n, 32, 0, 4, 8, 16, want_labels=True)
assert r.status == 0 and len(cl["road"]) == r.n_road and len(cl["curb"]) == r.n_curb
print(f"ok: road={r.n_road} curb={r.n_curb} roi={len(cl['roi'])} road_probably={len(cl['road_probably'])} launches={det.last_launch_count() if hasattr(det, 'last_launch_count') else '-'}")
det.close()
