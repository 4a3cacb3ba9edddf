#!/usr/bin/env python
"""Condense ncu outputs (gpurun_out/, scratch) into small text summaries under profiles/ (tracked).

  summarize_ncu.py launches <launches.csv> <out.md>          per-kernel launch count / total / share
  summarize_ncu.py full <report.ncu-rep> <out.md> [kernel]   selected raw metrics of the last matching launch
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__cycles_active.avg", "smsp__cycles_active.avg", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores"]
STALL = "smsp__average_warps_issue_stalled_"


def launches(path, out):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows:
        name = r[ik].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, r[ig], r[ib]])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary of `{path}` (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("| kernel | launches | total ms | mean us | share | grid | block |\n|---|---|---|---|---|---|---|\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name[:90]}` | {a[0]} | {a[1] / 1e6:.3f} | {a[1] / a[0] / 1e3:.1f} | {a[1] / tot:.1%} | {a[2]} | {a[3]} |\n")


def full(path, out, pattern=None):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    ik = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{path}`\n")
        for r in body:
            if pattern and pattern not in r[ik]:
                continue
            f.write(f"\n## {r[ik][:120]}\n\n| metric | value |\n|---|---|\n")
            for h, v in zip(hdr, r):
                if h in KEYS or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and float(v or 0) >= 0.05):
                    f.write(f"| {h} | {v} |\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
