#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_project.py tests/test_gpu_process.py tests/test_gpu_convert.py -m gpu -x -q ) > gpurun_out/pytest_k.log 2>&1
tail -4 gpurun_out/pytest_k.log
timeout 900 python bench.py --workload pgdb3q --batch 1024 --steps 1 --warmup 3 > gpurun_out/bench_pgdb3q_v7.json 2> gpurun_out/bench_pgdb3q_v7.err
cut -c1-200 gpurun_out/bench_pgdb3q_v7.json; tail -3 gpurun_out/bench_pgdb3q_v7.err
timeout 900 python bench.py --workload convert > gpurun_out/bench_convert_v4.json 2> gpurun_out/bench_convert_v4.err
python -c "
import json
d=json.load(open('gpurun_out/bench_convert_v4.json'))
for r in d['kernels']:
    if 'kraus2choi' in r['kernel']: print('%-70s %10.3f ms %8.1f GB/s %.3f  items=%d' % (r['kernel'][:70], r['ms'], r['achieved_gbs'], r['frac_of_hbm_peak'], r['items']))
"; tail -3 gpurun_out/bench_convert_v4.err
