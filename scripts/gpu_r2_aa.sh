#!/bin/bash
# round 2, call AA: streaming rows after the mle_step_herm register cap
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2aa_build.log 2>&1
timeout 900 python bench.py --workload streaming --no-cpu-baseline > gpurun_out/r2aa_bench_streaming.json 2> gpurun_out/r2aa_bench_streaming.err; echo "rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2aa_bench_streaming.json"))
for r in d["kernels"]: print("%.2f %8.3f ms  %s" % (r["frac_of_hbm_peak"], r["ms"], r["kernel"]))
PY
