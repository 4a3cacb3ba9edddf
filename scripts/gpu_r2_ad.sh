#!/bin/bash
# round 2, call AD: padded transposition tiles of the thread-per-item kernels -- tests + streaming / convert rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2ad_build.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_convert.py tests/test_gpu_state.py tests/test_gpu_project.py -m gpu -x -q > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ad_pytest.log
timeout 900 python bench.py --workload streaming --no-cpu-baseline > gpurun_out/r2ad_bench_streaming.json 2> gpurun_out/r2ad_bench_streaming.err; echo "rc=$?"
timeout 900 python bench.py --workload convert --no-cpu-baseline > gpurun_out/r2ad_bench_convert.json 2> gpurun_out/r2ad_bench_convert.err; echo "rc=$?"
python - <<PY
import json
for f in ("streaming", "convert"):
    d = json.load(open(f"gpurun_out/r2ad_bench_{f}.json"))
    for r in d["kernels"]:
        if f == "streaming" or "n=1" in r["kernel"]: print("%.2f %8.3f ms  %s" % (r["frac_of_hbm_peak"], r["ms"], r["kernel"][:90]))
PY
