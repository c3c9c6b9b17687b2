#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: srcprof.py dump.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = None; lines = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) > 8 and r[0] not in ("", "Line No") and r[0].isdigit():
        lines.append(r)
ci = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed"); si = hdr.index("# Samples")
tot = sum(int(r[ci]) for r in lines if r[ci].isdigit()); ts = sum(int(r[si]) for r in lines if r[si].isdigit())
print("total warp-inst", tot, "samples", ts)
for r in lines:
    if r[ci].isdigit() and (int(r[ci]) > minp / 100 * tot or int(r[si]) > minp / 100 * ts):
        print("%5s %6.2f%% inst  lanes %5.1f  samples %5.2f%% | %s" % (r[0], 100 * int(r[ci]) / tot, int(r[ti]) / max(1, int(r[ci])), 100 * int(r[si]) / ts, r[1].strip()[:120]))
