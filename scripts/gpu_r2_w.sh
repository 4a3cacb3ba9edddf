#!/bin/bash
# round 2, call W: structured-input fidelity test + distances bench on the kept tri kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2w_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_state.py -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2w_pytest.log
timeout 600 python bench.py --workload distances --no-cpu-baseline > gpurun_out/r2w_bench_dist.json 2> gpurun_out/r2w_bench_dist.err; echo "dist rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2w_bench_dist.json"))
print(d["value"], d["ms_per_step"], json.dumps(d["kernels"]), json.dumps(d["e2e"]))
PY
