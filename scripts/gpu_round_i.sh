#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_project.py tests/test_gpu_process.py tests/test_gpu_convert.py -m gpu -x -q ) > gpurun_out/pytest_i.log 2>&1
tail -8 gpurun_out/pytest_i.log
timeout 900 python bench.py --workload pgdb3q --batch 1024 --steps 1 --warmup 3 > gpurun_out/bench_pgdb3q_v5.json 2> gpurun_out/bench_pgdb3q_v5.err
cut -c1-900 gpurun_out/bench_pgdb3q_v5.json; tail -3 gpurun_out/bench_pgdb3q_v5.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pgdb_kernel -s 1 -c 1 -o gpurun_out/prof_pgdb3_v5 -f python bench.py --workload pgdb3q --batch 148 --steps 1 --warmup 3 > gpurun_out/ncu_full_pgdb3_v5.log 2>&1
tail -2 gpurun_out/ncu_full_pgdb3_v5.log
