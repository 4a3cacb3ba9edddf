#!/bin/bash
# round 2, call G: rotation-warp ring Jacobi vs the plain ring solver (results + cycles per step); PGDB parity + bench with it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 scripts/ubench_jacobi.bin > gpurun_out/r2g_ubench_jacobi.txt 2>&1; echo "ubench rc=$?"; cat gpurun_out/r2g_ubench_jacobi.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2g_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_process.py tests/test_gpu_project.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest.log
timeout 300 python bench.py --workload pgdb3q --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2g_bench_pgdb3q.json 2> gpurun_out/r2g_bench_pgdb3q.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2g_bench_pgdb3q.json"))
print("pgdb3q", round(d["value"], 1), "recon/s", d["config"]["jacobi_sweeps_per_eigh"], d["roofline"]["frac"])
PY
