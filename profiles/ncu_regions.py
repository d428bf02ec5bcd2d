#!/usr/bin/env python
"""Aggregate ncu per-line instruction counts into named source regions.
usage: python ncu_regions.py src.csv file:lo-hi=name ... ; prints instr per warp-sample given --per N"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
per = float(sys.argv[2])
regions = []
for a in sys.argv[3:]:
    spec, name = a.split("=")
    f, rng = spec.split(":")
    lo, hi = rng.split("-")
    regions.append((f, int(lo), int(hi), name))
cur = None; hdr = None
agg = {}; tot = 0; smp = {}; tots = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); ss = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= ie or r[0] in ("", "Line No"): continue
    try: n = float(r[ie]); s = float(r[ss]); ln = int(r[0])
    except ValueError: continue
    name = f"other:{cur}"
    for f, lo, hi, nm in regions:
        if f == cur and lo <= ln <= hi: name = nm; break
    agg[name] = agg.get(name, 0) + n; smp[name] = smp.get(name, 0) + s; tot += n; tots += s
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"{k:32s} {v/per:8.1f} instr/warp-sample  {100*v/tot:5.1f}% inst  {100*smp[k]/tots:5.1f}% stall samples")
print(f"{'total':32s} {tot/per:8.1f}")
