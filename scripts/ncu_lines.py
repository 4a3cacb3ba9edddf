#!/usr/bin/env python
"""Per-source-line hot spots of an ncu report (needs -lineinfo + --import-source on):
   ncu_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur = None; hdr = None; agg = []
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        agg.append((cur, int(r[0]), r[1], int(r[hdr.index("# Samples")] or 0), int(r[hdr.index("Instructions Executed")] or 0), r))
tot = sum(a[3] for a in agg)
names = ["stall_short_sb", "stall_barrier", "stall_math", "stall_wait", "stall_mio", "stall_long_sb", "stall_lg", "stall_not_selected"]
idx = [hdr.index(n) for n in names]
print("total samples", tot)
byfile = {}
for a in agg: byfile[a[0]] = byfile.get(a[0], 0) + a[3]
print({k: "%.1f%%" % (100 * v / tot) for k, v in byfile.items()})
print("file line  share  executed |", " ".join(n.replace("stall_", "") for n in names))
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    r = a[5]
    print("%-14s %4d %5.1f%% %.2e | %s | %s" % (a[0], a[1], 100 * a[3] / tot, a[4], " ".join("%6s" % r[i] for i in idx), a[2].strip()[:70]))
