#!/usr/bin/env python
"""Per-source-line summary of an ncu `--page source --csv --print-source cuda,sass` export.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass -k regex:KERNEL > src.csv; python tools/ncu_lines.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; data = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; hdr = None; continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].strip().isdigit():
        data.append((cur, dict(zip(range(len(hdr)), r)), hdr))
if not data:
    sys.exit("no per-line records")
hdr = data[0][2]
ie = hdr.index("Instructions Executed"); te = hdr.index("Thread Instructions Executed"); ss = hdr.index("# Samples")
tot = sum(float(d[ie] or 0) for _, d, _ in data); tots = sum(float(d[ss] or 0) for _, d, _ in data)
print(f"total warp instructions {tot:.3e}, samples {tots:.0f}")
data.sort(key=lambda x: -float(x[1][ss] or 0))
for f, d, _ in data[:top]:
    n = float(d[ie] or 0)
    print(f"{f:18s} {d[0]:>4s} inst {n / tot * 100:5.1f}%  samp {float(d[ss] or 0) / max(tots, 1) * 100:5.1f}%  thr/inst {float(d[te] or 0) / max(1.0, n):5.1f}  {d[1].strip()[:100]}")
