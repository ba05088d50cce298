#!/usr/bin/env python
"""Per-source-line dynamic instruction counts of one kernel from an ncu report captured with --import-source on.
usage: ncu_lines.py <report.ncu-rep> [top N]   (runs `ncu -i ... --page source --csv --print-source cuda,sass` here, no GPU)"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fpath = None; hdr = None; per = collections.Counter(); stall = collections.Counter(); src = {}; total = 0; ops = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ie = r.index("Instructions Executed"); ss = r.index("# Samples"); continue
    if r[0] != "":      # a source line
        cur = (fpath, int(r[0])); src[cur] = r[1].strip(); continue
    # a SASS row under the current source line
    try: n = int(r[ie]); s = int(r[ss])
    except ValueError: continue
    per[cur] += n; stall[cur] += s; total += n
    ops[r[3].split()[0].split(".")[0] if not r[3].strip().startswith("@") else r[3].split()[1].split(".")[0]] += n
stot = sum(stall.values())
print(f"total warp instructions {total}, stall samples {stot}")
for (k, n) in per.most_common(top):
    print(f"{100*n/total:5.1f}% inst {100*stall[k]/max(stot,1):5.1f}% stall  {k[0]}:{k[1]:<5d} {src.get(k,'')[:110]}")
print("--- opcode mix")
for k, n in ops.most_common(25): print(f"{100*n/total:5.1f}%  {k}")
