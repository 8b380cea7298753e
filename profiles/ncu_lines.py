#!/usr/bin/env python
"""Per-source-line executed warp instructions and stall samples of an .ncu-rep, all lines, sorted by file / line.
usage: python profiles/ncu_lines.py rep.ncu-rep > lines.txt"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--csv", "--page", "source", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None
agg = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name"):
        continue
    if r[0] != "" and r[2] == "-":
        try:
            agg[(cur, int(r[0]))] = (int(r[6]), int(r[7]), r[1][:110])
        except ValueError:
            pass
tot_s = sum(v[0] for v in agg.values()) or 1
tot_e = sum(v[1] for v in agg.values()) or 1
print("total samples %d, total executed warp instructions %d" % (tot_s, tot_e))
for (f, l), (s, e, src) in sorted(agg.items()):
    if e or s:
        print("%-22s %4d  exec %10d (%5.2f%%)  samples %6d (%5.2f%%)  %s" % (f, l, e, 100.0 * e / tot_e, s, 100.0 * s / tot_s, src))
