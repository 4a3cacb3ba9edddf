#!/bin/bash
# driver-like end-of-round flow on the final tree
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_final.log 2>&1
tail -4 gpurun_out/pytest_final.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke_final.log 2>&1
tail -3 gpurun_out/smoke_final.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 ) > gpurun_out/bench_reference_final.json 2> gpurun_out/bench_reference_final.err
cut -c1-200 gpurun_out/bench_reference_final.json
( time timeout 900 python bench.py --gpus 1 ) > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
cut -c1-400 gpurun_out/bench_final.json; tail -4 gpurun_out/bench_final.err
