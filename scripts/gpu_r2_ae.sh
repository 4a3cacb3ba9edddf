#!/bin/bash
# round 2, call AE: full GPU suite + the three kernel tables on the current tree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2ae_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ae_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ae_smoke.log 2>&1; echo "smoke rc=$?"
for w in streaming convert next; do timeout 900 python bench.py --workload $w > gpurun_out/r2ae_bench_$w.json 2> gpurun_out/r2ae_bench_$w.err; echo "$w rc=$?"; done
