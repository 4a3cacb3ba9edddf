#!/bin/bash
set -x
mkdir -p gpurun_out
./scripts/ubench_fp64.bin > gpurun_out/ubench_fp64.txt 2>&1; cat gpurun_out/ubench_fp64.txt
( timeout 900 python -m pytest tests/test_gpu_state.py -m gpu -x -q ) > gpurun_out/pytest_state.log 2>&1
tail -15 gpurun_out/pytest_state.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_mle2q_quad.json 2> gpurun_out/bench_mle2q_quad.err
cut -c1-900 gpurun_out/bench_mle2q_quad.json; tail -3 gpurun_out/bench_mle2q_quad.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mle_quad_kernel -s 1 -c 1 -o gpurun_out/prof_mle_quad -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_quad.log 2>&1
tail -3 gpurun_out/ncu_full_quad.log
