#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_project.py -m gpu -x -q ) > gpurun_out/pytest_z.log 2>&1
tail -5 gpurun_out/pytest_z.log
timeout 600 python bench.py --workload streaming > gpurun_out/bench_streaming_z.json 2> gpurun_out/bench_streaming_z.err
python -c "
import json
d=json.load(open('gpurun_out/bench_streaming_z.json'))
for r in d['kernels']:
    print('%-70s %10.3f ms %8.1f GB/s %.3f' % (r['kernel'][:70], r['ms'], r['achieved_gbs'], r['frac_of_hbm_peak']))
"; tail -3 gpurun_out/bench_streaming_z.err
