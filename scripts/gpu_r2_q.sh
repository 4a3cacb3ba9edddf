#!/bin/bash
# round 2, call Q: tridiagonal + QL fidelity kernel -- distance tests, empty batch test, distances bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2q_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_state.py tests/test_gpu_convert.py -m gpu -x -q -k "distance or fidelity or empty or purity" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2q_pytest.log
timeout 600 python bench.py --workload distances --no-cpu-baseline > gpurun_out/r2q_bench_dist.json 2> gpurun_out/r2q_bench_dist.err; echo "dist rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2q_bench_dist.json"))
print(d["value"], d["ms_per_step"], json.dumps(d["kernels"]))
PY
timeout 600 python scripts/fid_accuracy.py > gpurun_out/r2q_fid_accuracy.txt 2>&1; tail -12 gpurun_out/r2q_fid_accuracy.txt
