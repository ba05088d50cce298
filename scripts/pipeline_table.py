#!/usr/bin/env python
"""Turn the raw `ncu --csv` log of scripts/gpu_pipeline_table.sh into one row per kernel launch of the captured step.
usage: pipeline_table.py gpurun_out/pipeline_<tag>.csv profiles/<out>.csv [points_per_launch]"""
import csv, os, sys, collections
src, dst = sys.argv[1], sys.argv[2]
pts = float(sys.argv[3]) if len(sys.argv) > 3 else 131072 * 128
lines = open(src).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
launch = collections.OrderedDict()
for r in rows:
    k = int(r["ID"])
    d = launch.setdefault(k, {"kernel": r["Kernel Name"].split("(")[0].replace("urf::", "")})
    v = r["Metric Value"].replace(",", "")
    try: v = float(v)
    except ValueError: pass
    d[r["Metric Name"]] = (v, r["Metric Unit"])
def val(d, name, scale=1.0):
    if name not in d: return ""
    v, u = d[name]
    if u == "Kbyte": v *= 1e3
    elif u == "Mbyte": v *= 1e6
    elif u == "Gbyte": v *= 1e9
    elif u == "ns": v *= 1e-3
    elif u in ("msecond", "ms"): v *= 1e3
    elif u == "second": v *= 1e6
    return v * scale
out = []
tot = collections.Counter()
for k, d in launch.items():
    us = val(d, "gpu__time_duration.sum"); rd = val(d, "dram__bytes_read.sum"); wr = val(d, "dram__bytes_write.sum")
    inst = val(d, "smsp__inst_executed.sum"); l2 = val(d, "lts__t_bytes.sum")
    row = {"kernel": d["kernel"], "grid": int(val(d, "launch__grid_size")), "block": int(val(d, "launch__block_size")),
           "regs": int(val(d, "launch__registers_per_thread")), "time_us": round(us, 2), "dram_read_MB": round(rd / 1e6, 2),
           "dram_write_MB": round(wr / 1e6, 2), "dram_B_per_pt": round((rd + wr) / pts, 2), "l2_B_per_pt": round(l2 / pts, 2),
           "warp_inst": int(inst), "inst_per_pt": round(inst * 32 / pts / 32, 3) if inst else 0,
           "warp_inst_per_32pts": round(inst / (pts / 32), 1),
           "issue_pct": round(val(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
           "occupancy_pct": round(val(d, "sm__warps_active.avg.pct_of_peak_sustained_active"), 1)}
    for short, m in (("long_sb", "long_scoreboard"), ("short_sb", "short_scoreboard"), ("barrier", "barrier"), ("math", "math_pipe_throttle"),
                     ("mio", "mio_throttle"), ("lg", "lg_throttle"), ("wait", "wait"), ("branch", "branch_resolving")):
        row["stall_" + short] = round(val(d, f"smsp__average_warps_issue_stalled_{m}_per_issue_active.ratio"), 2)
    out.append(row)
    for f in ("time_us", "dram_read_MB", "dram_write_MB", "warp_inst"): tot[f] += row[f]
w = csv.DictWriter(open(dst, "w", newline=""), fieldnames=list(out[0].keys()))
w.writeheader(); w.writerows(out)
w.writerow({"kernel": "TOTAL", "time_us": round(tot["time_us"], 1), "dram_read_MB": round(tot["dram_read_MB"], 1),
            "dram_write_MB": round(tot["dram_write_MB"], 1), "dram_B_per_pt": round((tot["dram_read_MB"] + tot["dram_write_MB"]) * 1e6 / pts, 1),
            "warp_inst": tot["warp_inst"], "warp_inst_per_32pts": round(tot["warp_inst"] / (pts / 32), 1)})
for r in out: print(f'{r["kernel"]:22s} {r["time_us"]:8.1f}us dram {r["dram_B_per_pt"]:6.1f} B/pt  l2 {r["l2_B_per_pt"]:6.1f} B/pt  inst/warp {r["warp_inst_per_32pts"]:7.1f}  issue {r["issue_pct"]:5.1f}%  occ {r["occupancy_pct"]:5.1f}%  regs {r["regs"]}  long_sb {r["stall_long_sb"]}')
if len(sys.argv) > 4:            # per-kernel DRAM bytes of this capture -> the file bench.py reads for roofline.traffic
    import json
    tj = {}
    for r in out:
        name = r["kernel"].replace("void ", "").split("<")[0]
        tj[name] = tj.get(name, 0.0) + (r["dram_read_MB"] + r["dram_write_MB"]) * 1e6
    tj["_source"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch, {os.path.basename(src)} (scripts/gpu_pipeline_table.sh, batch of {int(pts) // 131072} C2 scans)"
    json.dump(tj, open(sys.argv[4], "w"), indent=1)
print("TOTAL", round(tot["time_us"], 1), "us; dram", round((tot["dram_read_MB"] + tot["dram_write_MB"]) * 1e6 / pts, 1), "B/pt; warp inst per 32 pts", round(tot["warp_inst"] / (pts / 32), 1))
