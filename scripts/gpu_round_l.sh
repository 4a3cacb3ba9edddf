#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_l.log 2>&1
tail -5 gpurun_out/pytest_l.log
timeout 600 python bench.py > gpurun_out/bench_mle2q_v4.json 2> gpurun_out/bench_mle2q_v4.err
cut -c1-400 gpurun_out/bench_mle2q_v4.json; tail -3 gpurun_out/bench_mle2q_v4.err
timeout 900 python bench.py --workload pgdb3q --batch 1024 --steps 2 --warmup 3 > gpurun_out/bench_pgdb3q_v8.json 2> gpurun_out/bench_pgdb3q_v8.err
cut -c1-300 gpurun_out/bench_pgdb3q_v8.json; tail -3 gpurun_out/bench_pgdb3q_v8.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_mle2q_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mle_quad_kernel -s 1 -c 1 -o gpurun_out/prof_mle_quad_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_quad_final.log 2>&1
tail -2 gpurun_out/ncu_full_quad_final.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pgdb_kernel -s 1 -c 1 -o gpurun_out/prof_pgdb3_final -f python bench.py --workload pgdb3q --batch 148 --steps 1 --warmup 3 > gpurun_out/ncu_full_pgdb3_final.log 2>&1
tail -2 gpurun_out/ncu_full_pgdb3_final.log
