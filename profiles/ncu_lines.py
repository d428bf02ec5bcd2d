#!/usr/bin/env python
"""Summarise an ncu source page (cuda,sass view) per CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv; python ncu_lines.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
hdr = None
out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
        ss = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ie or r[0] in ("", "Line No"):
        continue
    try:
        n = float(r[ie]); smp = float(r[ss])
    except ValueError:
        continue
    out.append((n, smp, cur, r[0], r[1].strip()[:100]))
tot = sum(o[0] for o in out)
tots = sum(o[1] for o in out)
print(f"total warp instructions {tot:.4g}, stall samples {tots:.4g}")
for n, smp, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100*n/tot:6.2f}% inst {100*smp/max(tots,1):6.2f}% smp  {f}:{ln:<5s} {src}")
