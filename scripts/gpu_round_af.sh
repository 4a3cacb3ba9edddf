#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_state.py tests/test_gpu_project.py -m gpu -x -q ) > gpurun_out/pytest_af.log 2>&1
tail -3 gpurun_out/pytest_af.log
timeout 600 python bench.py --workload streaming > gpurun_out/bench_streaming_af.json 2> gpurun_out/bench_streaming_af.err
python -c "
import json
d=json.load(open('gpurun_out/bench_streaming_af.json'))
for r in d['kernels']:
    print('%-70s %10.3f ms %8.1f GB/s %.3f' % (r['kernel'][:70], r['ms'], r['achieved_gbs'], r['frac_of_hbm_peak']))
"
