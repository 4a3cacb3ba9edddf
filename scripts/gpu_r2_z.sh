#!/bin/bash
# round 2, call Z: per-launch times of the two TNI passes (ncu launch list; the cold-cache times are only read as shares)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2z_build.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:"tni_|proj_tp_kernel|proj_cp" --csv --log-file gpurun_out/r2z_tni.csv python bench.py --workload streaming --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2z_ncu.log 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2z_tni.csv")) if len(r) > 8]
h = rows[0]
ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
iid = h.index("ID")
seen = {}
for r in rows[1:]:
    seen.setdefault((r[iid], r[ik][:60]), {})[r[im]] = r[iv]
last = {}
for (i, k), m in seen.items():
    last[k] = m
for k, m in last.items(): print(k, m)
PY
