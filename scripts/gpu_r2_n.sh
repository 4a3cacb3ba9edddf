#!/bin/bash
# round 2, call N: MLE warp kernel fast path (unit coefficients): parity + 3-qubit throughput
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2n_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_state.py -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
timeout 300 python bench.py --workload mle3q --no-cpu-baseline > gpurun_out/r2n_bench_mle3q.json 2> gpurun_out/r2n_bench_mle3q.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2n_bench_mle3q.json") if l.startswith("{")][0])
print("mle3q", round(d["value"]), "recon/s", d["ms_per_step"], "ms", d["roofline"]["frac"])
PY
