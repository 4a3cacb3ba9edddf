#!/bin/bash
# round 2, call J: register-pass PTM kernels at n = 4, 5: parity + conversion sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2j_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_convert.py -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest.log
timeout 900 python bench.py --workload convert --no-cpu-baseline > gpurun_out/r2j_bench_convert.json 2> gpurun_out/r2j_bench_convert.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2j_bench_convert.json"))
for r in d["kernels"]:
    print(r["kernel"][:60].ljust(60), r["ms"], r["frac_of_hbm_peak"])
PY
