#!/bin/bash
# round 2, call D: role-split ring Jacobi (A warps / V warps) vs the plain ring solver; Kronecker-structured T / gradient
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 scripts/ubench_jacobi.bin > gpurun_out/r2d_ubench_jacobi.txt 2>&1; echo "ubench rc=$?"
head -12 gpurun_out/r2d_ubench_jacobi.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2d_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_process.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest.log
timeout 300 python bench.py --workload pgdb3q --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2d_bench_pgdb3q.json 2> gpurun_out/r2d_bench_pgdb3q.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2d_bench_pgdb3q.json"))
print("pgdb3q", round(d["value"], 1), "recon/s", d["config"]["jacobi_sweeps_per_eigh"], d["roofline"]["frac"])
PY
