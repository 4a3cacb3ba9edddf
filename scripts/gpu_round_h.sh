#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_state.py -m gpu -x -q ) > gpurun_out/pytest_h.log 2>&1
tail -5 gpurun_out/pytest_h.log
python scripts/exp_mle_batch.py 2>&1 | tee gpurun_out/exp_mle_batch.txt
timeout 600 python bench.py --workload distances > gpurun_out/bench_distances_v2.json 2> gpurun_out/bench_distances_v2.err
python -c "
import json
d=json.load(open('gpurun_out/bench_distances_v2.json'))
for r in d['kernels']: print('%-70s %10.3f ms %8.1f GB/s %.3f  batch_ms=%s' % (r['kernel'][:70], r['ms'], r['achieved_gbs'], r['frac_of_hbm_peak'], r['batch_ms']))
"; tail -3 gpurun_out/bench_distances_v2.err
