#!/usr/bin/env python
"""One ingest stream over N GPUs (BASELINE config 4: OS2-128 262,144-point scans): writes a few distinct synthetic scans to a
file and runs tools/mq_bench (built in-tree by urban_road_filter_b200/build.py) for several producer counts.
usage: python scripts/bench_mq.py --gpus 8 [--shape C4] [--scans 3000] [--producers 1,2,4,8]"""
import argparse, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from urban_road_filter_b200.synth import SHAPES, make_scan
from urban_road_filter_b200 import build

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--shape", default="C4")
ap.add_argument("--scans", type=int, default=3000)
ap.add_argument("--producers", default="1,4")
ap.add_argument("--slots", type=int, default=24)
ap.add_argument("--max-batch", type=int, default=16)
args = ap.parse_args()
exe = build.build_tools()
sh = SHAPES[args.shape]
K = 16
path = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(), f"urf_{args.shape}.bin")
with open(path, "wb") as f:
    for k in range(K):
        f.write(np.ascontiguousarray(make_scan(args.shape, 500 + k), np.float32).tobytes())
n = sh.rings * sh.cols
for p in args.producers.split(","):
    subprocess.run([exe, path, str(n), str(K), str(args.gpus), p, str(args.scans), str(args.slots), str(args.max_batch), "1", str(sh.channels),
                    str(sh.interval)], check=True, env={**os.environ, "LD_LIBRARY_PATH": os.path.join(ROOT, "urban_road_filter_b200")})
os.remove(path)
