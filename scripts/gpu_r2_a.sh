#!/bin/bash
# round 2, call A: parity suite after the contract fixes (counters printed), DFMA/DMMA peaks, r01 bench for reference
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2a_build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 120 scripts/ubench_dmma.bin > gpurun_out/r2a_fp64_peaks.json 2> gpurun_out/r2a_ubench.err; cat gpurun_out/r2a_fp64_peaks.json
timeout 300 python bench.py --workload pgdb3q --steps 2 --warmup 1 > gpurun_out/r2a_bench_pgdb3q.json 2> gpurun_out/r2a_bench_pgdb3q.err; tail -c 600 gpurun_out/r2a_bench_pgdb3q.json
