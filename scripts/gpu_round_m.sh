#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_convert.py -m gpu -x -q ) > gpurun_out/pytest_m.log 2>&1
tail -6 gpurun_out/pytest_m.log
timeout 900 python bench.py --workload convert > gpurun_out/bench_convert_v5.json 2> gpurun_out/bench_convert_v5.err
python -c "
import json
d=json.load(open('gpurun_out/bench_convert_v5.json'))
for r in d['kernels']:
    if 'liouville' in r['kernel']: print('%-70s %10.3f ms %8.1f GB/s %.3f  items=%d' % (r['kernel'][:70], r['ms'], r['achieved_gbs'], r['frac_of_hbm_peak'], r['items']))
"; tail -3 gpurun_out/bench_convert_v5.err
