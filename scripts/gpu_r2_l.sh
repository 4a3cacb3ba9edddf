#!/bin/bash
# round 2, call L: which "streaming" kernels are really HBM-bound?  pipe utilisation + DRAM throughput per kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2l_build.log 2>&1
M=gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.per_cycle_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,sm__maximum_warps_per_active_cycle_pct
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2l_streaming_metrics.csv python bench.py --workload streaming > gpurun_out/r2l_streaming.json 2> gpurun_out/r2l_streaming.err
python - <<PY
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/r2l_streaming_metrics.csv") if l.startswith('"'))]
h = rows[0]; ik, im, iv, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault((r[ii], r[ik].split("(")[0][-60:]), {})[r[im]] = r[iv]
seen = {}
for (i, k), m in agg.items():
    seen[k] = m
for k, m in seen.items():
    print(k.ljust(60), "ms", round(float(m["gpu__time_duration.sum"].replace(",", "")) / 1e6, 3), "dram%", m["dram__throughput.avg.pct_of_peak_sustained_elapsed"], "fp64%", m["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"], "ipc", m["smsp__issue_active.avg.per_cycle_active"], "warps%", m["sm__warps_active.avg.pct_of_peak_sustained_active"], "regs", m["launch__registers_per_thread"])
PY
