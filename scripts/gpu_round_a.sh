#!/bin/bash
# first GPU pass of a session: parity tests, bench lines, ncu launch list, one full capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_mle2q.json 2> gpurun_out/bench_mle2q.err
tail -c 3000 gpurun_out/bench_mle2q.json
timeout 600 python bench.py --workload pgdb3q --batch 296 --steps 2 --warmup 3 > gpurun_out/bench_pgdb3q.json 2> gpurun_out/bench_pgdb3q.err
tail -c 2500 gpurun_out/bench_pgdb3q.json
timeout 300 python bench.py --workload pgdb2q --batch 4096 --steps 2 --warmup 3 > gpurun_out/bench_pgdb2q.json 2> gpurun_out/bench_pgdb2q.err
tail -c 1500 gpurun_out/bench_pgdb2q.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_mle2q.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mle_reg_kernel -s 1 -c 1 -o gpurun_out/prof_mle_reg -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
