#!/bin/bash
# round 2, call AG: mle_warp_kernel block size for one-wave batches -- 3-qubit MLE row + state tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2ag_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_state.py -m gpu -x -q > gpurun_out/r2ag_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ag_pytest.log
timeout 600 python bench.py --workload mle3q --no-cpu-baseline > gpurun_out/r2ag_bench_mle3q.json 2> gpurun_out/r2ag_bench_mle3q.err; echo "rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2ag_bench_mle3q.json") if l.startswith("{")][0])
print(json.dumps(d)[:700])
PY
