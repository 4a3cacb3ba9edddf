#!/bin/bash
# round 2, call X: compacting stream kernel for 3/5/6/7-column shot arrays -- parity tests + the "next" rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2x_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_moments.py -m gpu -x -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2x_pytest.log
timeout 900 python bench.py --workload next --no-cpu-baseline > gpurun_out/r2x_bench_next.json 2> gpurun_out/r2x_bench_next.err; echo "next rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2x_bench_next.json"))
for r in d["kernels"]:
    if "moments" in r["kernel"]: print("%.2f %8.3f ms  %s" % (r["frac_of_hbm_peak"], r["ms"], r["kernel"]))
PY
