#!/bin/bash
# what the driver does at round end: GPU tests, smoke(), reference arm, our arm; then the ncu evidence of the bench command
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_final.log 2>&1
tail -6 gpurun_out/pytest_final.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_final.log 2>&1
tail -4 gpurun_out/smoke_final.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 ) > gpurun_out/bench_reference_final.json 2> gpurun_out/bench_reference_final.err
cut -c1-300 gpurun_out/bench_reference_final.json
( time timeout 900 python bench.py --gpus 1 ) > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
cut -c1-2500 gpurun_out/bench_final.json; tail -4 gpurun_out/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_mle2q_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mle_quad_kernel -s 1 -c 1 -o gpurun_out/prof_mle_quad_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_quad_final.log 2>&1
tail -2 gpurun_out/ncu_full_quad_final.log
