#!/usr/bin/env python
"""Markdown tables for DESIGN.md / profiles/README.md from the committed evidence files.
usage: make_tables.py kernels <pipeline.csv> <bench.json>   |   configs <json> ...   |   scale <json> ..."""
import csv, json, sys
def last_json(path): return json.loads(open(path).read().strip().splitlines()[-1])
mode = sys.argv[1]
if mode == "kernels":
    rows = list(csv.DictReader(open(sys.argv[2])))
    d = last_json(sys.argv[3])
    live = d["roofline"]["kernel_ms_per_step"]
    tot = sum(live.values())
    print("| kernel | live ms (share) | ncu µs | DRAM B/pt | L2 B/pt | warp-inst / 32 pts | issue % | occupancy % | regs |")
    print("|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        k = r["kernel"].replace("void ", "")
        if k == "TOTAL":
            print(f"| **step** | {tot:.3f} (serial) | {float(r['time_us']):.0f} | {r['dram_B_per_pt']} | | {r['warp_inst_per_32pts']} | | | |")
            continue
        base = k.split("<")[0]
        ms = live.get(base, live.get(k, 0.0))
        print(f"| `{k}` | {ms:.3f} ({100 * ms / tot:.1f} %) | {float(r['time_us']):.0f} | {r['dram_B_per_pt']} | {r['l2_B_per_pt']} | {r['warp_inst_per_32pts']} | {r['issue_pct']} | {r['occupancy_pct']} | {r['regs']} |")
elif mode == "configs":
    print("| config | points | batch | scans/s | Mpoints/s | ms/step | e2e scans/s | with order scans/s | dominant kernel (frac of HBM peak) |")
    print("|---|---|---|---|---|---|---|---|---|")
    for p in sys.argv[2:]:
        d = last_json(p)
        c = d["config"]; r = d["roofline"]
        c = {**c, "batch_per_gpu": d.get("batch_per_gpu", c.get("batch_per_gpu")), "mpoints_per_sec": d.get("mpoints_per_sec", c.get("mpoints_per_sec"))}
        print(f"| {c['workload'].split(',')[0]} | {c['points_per_scan']:,} | {c['batch_per_gpu']} | {d['value']:,.0f} | {c['mpoints_per_sec']:,.0f} | {d['ms_per_step']:.3f} | {d['e2e']['value']:,.0f} | {d.get('with_order', {}).get('value', 0):,.0f} | `{r['kernel']}` {r['frac']:.3f} |")
elif mode == "scale":
    print("| GPUs | scans/s (HBM-resident) | efficiency | e2e scans/s | efficiency | lean e2e | with order |")
    print("|---|---|---|---|---|---|---|")
    base = None
    for p in sys.argv[2:]:
        d = last_json(p)
        n = d["n_gpus"]
        if base is None: base = (d["value"] / n, d["e2e"]["value"] / n)
        print(f"| {n} | {d['value']:,.0f} | {d['value'] / n / base[0]:.3f} | {d['e2e']['value']:,.0f} | {d['e2e']['value'] / n / base[1]:.3f} | {d.get('e2e_lean', {}).get('value', 0):,.0f} | {d.get('with_order', {}).get('value', 0):,.0f} |")
