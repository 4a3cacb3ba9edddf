#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_e.log 2>&1
tail -8 gpurun_out/pytest_e.log
timeout 600 python bench.py --workload pgdb3q --batch 148 --steps 1 --warmup 3 > gpurun_out/bench_pgdb3q_v3.json 2> gpurun_out/bench_pgdb3q_v3.err
cut -c1-1000 gpurun_out/bench_pgdb3q_v3.json; tail -3 gpurun_out/bench_pgdb3q_v3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pgdb_kernel -s 1 -c 1 -o gpurun_out/prof_pgdb3_v3 -f python bench.py --workload pgdb3q --batch 148 --steps 1 --warmup 3 > gpurun_out/ncu_full_pgdb3_v3.log 2>&1
tail -2 gpurun_out/ncu_full_pgdb3_v3.log
