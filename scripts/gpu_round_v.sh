#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_state.py tests/test_gpu_moments.py tests/test_gpu_process.py -m gpu -x -q ) > gpurun_out/pytest_v.log 2>&1
tail -4 gpurun_out/pytest_v.log
timeout 300 python bench.py --workload distances > gpurun_out/bench_distances_v.json 2> gpurun_out/bench_distances_v.err
cut -c1-330 gpurun_out/bench_distances_v.json
