#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_project.py tests/test_gpu_process.py tests/test_gpu_state.py -m gpu -x -q ) > gpurun_out/pytest_c.log 2>&1
tail -15 gpurun_out/pytest_c.log
timeout 600 python bench.py --workload pgdb3q --batch 296 --steps 2 --warmup 3 > gpurun_out/bench_pgdb3q_v1.json 2> gpurun_out/bench_pgdb3q_v1.err
cut -c1-1200 gpurun_out/bench_pgdb3q_v1.json; tail -3 gpurun_out/bench_pgdb3q_v1.err
timeout 300 python bench.py --workload pgdb2q --batch 4096 --steps 2 --warmup 3 > gpurun_out/bench_pgdb2q_v1.json 2> gpurun_out/bench_pgdb2q_v1.err
cut -c1-700 gpurun_out/bench_pgdb2q_v1.json
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_mle2q_v2.json 2> gpurun_out/bench_mle2q_v2.err
cut -c1-400 gpurun_out/bench_mle2q_v2.json
