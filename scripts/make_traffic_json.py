#!/usr/bin/env python
"""profiles/r02_traffic.json from ncu reports: dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel.
usage: make_traffic_json.py out.json name=report.ncu-rep[:kernel-substring] ...   (bench.py reads the result)"""
import csv, io, json, subprocess, sys


def traffic(path, pattern):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ik, ir, iw, it = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                             "gpu__time_duration.sum"))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    best = None
    for r in rows[2:]:
        if len(r) != len(hdr) or (pattern and pattern not in r[ik]):
            continue
        rd = float(r[ir].replace(",", "")) * scale.get(units[ir], 1)
        wr = float(r[iw].replace(",", "")) * scale.get(units[iw], 1)
        best = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "kernel": r[ik][:120],
                "ncu_time": r[it] + " " + units[it]}
    return best


out = {}
for spec in sys.argv[2:]:
    name, rest = spec.split("=", 1)
    path, _, pat = rest.partition(":")
    t = traffic(path, pat or None)
    if t:
        t["source"] = f"ncu --set full --clock-control none ({path.split('/')[-1]})"
        out[name] = t
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1))
