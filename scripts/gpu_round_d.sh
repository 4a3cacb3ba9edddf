#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_convert.py tests/test_gpu_process.py -m gpu -x -q ) > gpurun_out/pytest_d.log 2>&1
tail -15 gpurun_out/pytest_d.log
timeout 600 python bench.py --workload pgdb3q --batch 148 --steps 1 --warmup 3 > gpurun_out/bench_pgdb3q_v2.json 2> gpurun_out/bench_pgdb3q_v2.err
cut -c1-1000 gpurun_out/bench_pgdb3q_v2.json; tail -3 gpurun_out/bench_pgdb3q_v2.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pgdb_kernel -s 1 -c 1 -o gpurun_out/prof_pgdb3 -f python bench.py --workload pgdb3q --batch 148 --steps 1 --warmup 3 > gpurun_out/ncu_full_pgdb3.log 2>&1
tail -3 gpurun_out/ncu_full_pgdb3.log
