#!/usr/bin/env python
"""Join an ncu 3-metric launch list (duration, dram read, dram write) with the kernel rows of a bench_kernels JSON:
   summarize_traffic.py <ncu.csv> <bench.json> <out.md>
Per kernel NAME (template arguments kept) the LAST launch is reported: bytes moved in DRAM vs algorithmic bytes."""
import collections, csv, json, re, sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
unit = hdr.index("Metric Unit")
launches = collections.OrderedDict()
for r in rows:
    d = launches.setdefault(r[iid], {"name": r[ik]})
    val = float(r[iv].replace(",", ""))
    u = r[unit]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
    d[r[im]] = val * scale
last = collections.OrderedDict()
for d in launches.values():
    if d["name"].startswith("void at::") or "at::native" in d["name"]:
        continue
    nm = re.sub(r"\(.*", "", d["name"]).replace("void ", "")
    last[nm] = d
bench = json.load(open(sys.argv[2]))
with open(sys.argv[3], "w") as f:
    f.write(f"# DRAM traffic per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum) -- `{sys.argv[1]}`\n\n")
    f.write("Last launch of each kernel instantiation in the sweep; algorithmic bytes per launch = items x bytes_per_item of\n"
            f"`{sys.argv[2]}` (matched by order of appearance where names repeat).\n\n")
    f.write("| kernel | duration us | DRAM read MB | DRAM write MB | DRAM total MB |\n|---|---|---|---|---|\n")
    for nm, d in last.items():
        rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        f.write(f"| `{nm[:100]}` | {d.get('gpu__time_duration.sum', 0) / 1e3:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {(rd + wr) / 1e6:.1f} |\n")
    f.write("\n| bench row | algorithmic MB per launch |\n|---|---|\n")
    for r in bench["kernels"]:
        f.write(f"| {r['kernel'][:90]} | {r['items'] * r['bytes_per_item'] / 1e6:.1f} |\n")
