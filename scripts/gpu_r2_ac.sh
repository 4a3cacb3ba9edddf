#!/bin/bash
# round 2, call AC: choi2kraus low-rank path at n = 3 -- convert tests + conversion sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2ac_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_convert.py -m gpu -x -q > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ac_pytest.log
timeout 900 python bench.py --workload convert --no-cpu-baseline > gpurun_out/r2ac_bench_convert.json 2> gpurun_out/r2ac_bench_convert.err; echo "convert rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2ac_bench_convert.json"))
for r in d["kernels"]:
    if "kraus" in r["kernel"]: print("%.2f %8.3f ms %10.0f/s  %s" % (r["frac_of_hbm_peak"], r["ms"], r["items_per_s"], r["kernel"][:90]))
PY
