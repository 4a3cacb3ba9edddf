#!/bin/bash
# round 2, call K: fused two-pass PTM (L2 ring) at n = 4, 5: parity, conversion sweep, DRAM traffic
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2k_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_convert.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log
timeout 900 python bench.py --workload convert --no-cpu-baseline > gpurun_out/r2k_bench_convert.json 2> gpurun_out/r2k_bench_convert.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2k_bench_convert.json"))
for r in d["kernels"]:
    if "pauli" in r["kernel"]: print(r["kernel"][:60].ljust(60), r["ms"], r["frac_of_hbm_peak"])
PY
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"pl_fused_kernel" -c 6 --csv --log-file gpurun_out/r2k_traffic_fused.csv python bench.py --workload convert --no-cpu-baseline > /dev/null 2>&1
grep pl_fused gpurun_out/r2k_traffic_fused.csv | head -12 | cut -c1-260
