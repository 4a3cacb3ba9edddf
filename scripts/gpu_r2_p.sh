#!/bin/bash
# round 2, call P: full GPU suite on the current tree + conversion sweep (choi2kraus fast path rows)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2p_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2p_pytest.log
timeout 900 python bench.py --workload convert --no-cpu-baseline > gpurun_out/r2p_bench_convert.json 2> gpurun_out/r2p_bench_convert.err; echo "convert rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2p_bench_convert.json"))
for r in d["kernels"]:
    if "kraus" in r["kernel"]: print(r["kernel"][:100].ljust(100), r["ms"], round(r["items_per_s"]))
PY
