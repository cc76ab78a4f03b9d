#!/usr/bin/env python
"""Executed warp-instructions and stall samples of one kernel split at its block barriers (BAR.SYNC), i.e. per phase of
the tile kernel, from an ncu SASS source page.  Inlined helpers lose their call site in the line table, so the per-line
profile cannot say which phase an FFMA2 belongs to; the position in the instruction stream can.

    ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv
    python profiles/tools/segment_profile.py sass.csv
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
seg, segs = dict(n=0, inst=0, samp=0, ops=Counter(), first=0), []
tot = sum(int(r[ci["Instructions Executed"]] or 0) for r in body)
tots = sum(int(r[ci["# Samples"]] or 0) for r in body)
for k, r in enumerate(body):
    src = r[ci["Source"]]
    ie, ss = int(r[ci["Instructions Executed"]] or 0), int(r[ci["# Samples"]] or 0)
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    seg["n"] += 1; seg["inst"] += ie; seg["samp"] += ss; seg["ops"][op.split(".")[0]] += ie
    if op.startswith("BAR") or op.startswith("EXIT"):
        seg["last"] = k; seg["end"] = src; segs.append(seg)
        seg = dict(n=0, inst=0, samp=0, ops=Counter(), first=k + 1)
if seg["n"]:
    seg["last"] = len(body) - 1; seg["end"] = "(tail)"; segs.append(seg)
print(f"total warp-instructions {tot}, samples {tots}")
for i, s in enumerate(segs):
    top = ", ".join(f"{o} {100.0 * c / max(s['inst'], 1):.0f}%" for o, c in s["ops"].most_common(6))
    print(f"seg {i:2d} sass[{s['first']:5d}..{s['last']:5d}] static {s['n']:5d}  inst {100.0 * s['inst'] / tot:5.1f}%  samples {100.0 * s['samp'] / tots:5.1f}%  "
          f"ipc-proxy {s['inst'] / max(s['samp'], 1) / (tot / tots):4.2f}  | {top}")
