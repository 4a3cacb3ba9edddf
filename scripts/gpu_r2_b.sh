#!/bin/bash
# round 2, call B: second-generation Dykstra (early-stop Jacobi + Loewner correction) -- parity, then throughput vs tolerance
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2b_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_process.py tests/test_gpu_project.py -m gpu -x -q -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
for tol in default 1e-6 1e-8; do
  extra=""; [ "$tol" != default ] && extra="--eigh-tol $tol"
  timeout 300 python bench.py --workload pgdb3q --steps 2 --warmup 1 --no-cpu-baseline $extra > gpurun_out/r2b_bench_pgdb3q_$tol.json 2> gpurun_out/r2b_bench_pgdb3q_$tol.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2b_bench_pgdb3q_$tol.json"))
    print("$tol", round(d["value"], 1), "recon/s", d["config"]["jacobi_sweeps_per_eigh"], d["config"]["eigh_calls_mean"], d["roofline"]["frac"])
except Exception as e:
    print("$tol failed", e); print(open("gpurun_out/r2b_bench_pgdb3q_$tol.err").read()[-1500:])
PY
done
timeout 900 python bench.py > gpurun_out/r2b_bench_all.json 2> gpurun_out/r2b_bench_all.err; echo "bench all rc=$?"; tail -c 400 gpurun_out/r2b_bench_all.err
