#!/usr/bin/env python
"""Per-source-line instruction shares of one kernel from an ncu report (source page, cuda+sass correlation).
    python scripts/ncu_lines.py report.ncu-rep 'regex:k_scan' [launch-skip] [top]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
per = defaultdict(lambda: [0.0, 0.0, ""])
fname, hdr = "", None
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ie, st = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ie:
        continue
    try:
        n = float(r[ie]); s = float(r[st] or 0)
    except ValueError:
        continue
    key = (fname, r[0])
    per[key][0] += n; per[key][1] += s
    if r[1]:
        per[key][2] = r[1]
tot = sum(v[0] for v in per.values()); ts = sum(v[1] for v in per.values()) or 1
print(f"total warp instructions {tot:.0f}, stall samples {ts:.0f}")
for (f, l), v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0] / tot * 100:5.1f}% inst {v[1] / ts * 100:5.1f}% smp  {f}:{l:>4}  {v[2].strip()[:140]}")
