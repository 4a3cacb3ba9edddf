#!/bin/bash
# round 2, call AB: blocked linear-inversion process kernel -- parity tests + next rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2ab_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_process.py tests/test_gpu_moments.py tests/test_gpu_convert.py -m gpu -x -q -k "linear_inv or golden or empty" > gpurun_out/r2ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ab_pytest.log
timeout 900 python bench.py --workload next --no-cpu-baseline > gpurun_out/r2ab_bench_next.json 2> gpurun_out/r2ab_bench_next.err; echo "next rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2ab_bench_next.json"))
for r in d["kernels"]:
    if "moments" not in r["kernel"]: print("%.2f %8.3f ms  %s" % (r["frac_of_hbm_peak"], r["ms"], r["kernel"]))
PY
