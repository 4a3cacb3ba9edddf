#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_state.py -m gpu -x -q -k "single_step" ) > gpurun_out/pytest_ab.log 2>&1
tail -3 gpurun_out/pytest_ab.log
timeout 600 python bench.py --workload streaming > gpurun_out/bench_streaming_ab.json 2> gpurun_out/bench_streaming_ab.err
python -c "
import json
d=json.load(open('gpurun_out/bench_streaming_ab.json'))
for r in d['kernels'][:6]:
    print('%-70s %10.3f ms %8.1f GB/s %.3f' % (r['kernel'][:70], r['ms'], r['achieved_gbs'], r['frac_of_hbm_peak']))
"
