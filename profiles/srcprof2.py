#!/usr/bin/env python
"""Per-source-line summary of one kernel launch from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` (a dump holds
one block of sections per launch).  usage: srcprof2.py dump.csv <launch index | -1> [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]); minp = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
# a launch = a run of sections; a new launch starts when a "Function Name" row follows a section that is not the first file of a launch.
starts = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
# sections alternate (file A, file B, ...) per launch: group by repeating pattern of the first section's first data row
first_keys = []
for s in starts:
    k = rows[s + 2][:2] if s + 2 < len(rows) else None
    first_keys.append(tuple(k) if k else None)
launch_starts = [s for s, k in zip(starts, first_keys) if k == first_keys[0]]
launch_starts.append(len(rows))
L = launch_starts[which] if which >= 0 else launch_starts[len(launch_starts) - 2]
E = launch_starts[launch_starts.index(L) + 1]
hdr = None; lines = []
for r in rows[L:E]:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[7].isdigit(): lines.append(r)
ci = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed"); si = hdr.index("# Samples")
tot = sum(int(r[ci]) for r in lines); ts = max(1, sum(int(r[si]) for r in lines)); tt = sum(int(r[ti]) for r in lines)
print("launch rows %d..%d  warp-inst %d  thread-inst %d  avg lanes %.1f  samples %d" % (L, E, tot, tt, tt / max(1, tot), ts))
for r in lines:
    if int(r[ci]) > minp / 100 * tot or int(r[si]) > minp / 100 * ts:
        print("%5s %6.2f%% inst  lanes %5.1f  samples %5.2f%% | %s" % (r[0], 100 * int(r[ci]) / tot, int(r[ti]) / max(1, int(r[ci])), 100 * int(r[si]) / ts, r[1].strip()[:110]))
