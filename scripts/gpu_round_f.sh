#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_f.log 2>&1
tail -5 gpurun_out/pytest_f.log
timeout 600 python bench.py --workload streaming > gpurun_out/bench_streaming.json 2> gpurun_out/bench_streaming.err
python -c "
import json
d=json.load(open('gpurun_out/bench_streaming.json'))
for r in d['kernels']: print('%-70s %10.3f ms %8.1f GB/s %.3f' % (r['kernel'][:70], r['ms'], r['achieved_gbs'], r['frac_of_hbm_peak']))
"; tail -3 gpurun_out/bench_streaming.err
timeout 900 python bench.py --workload convert > gpurun_out/bench_convert.json 2> gpurun_out/bench_convert.err
python -c "
import json
d=json.load(open('gpurun_out/bench_convert.json'))
for r in d['kernels']: print('%-70s %10.3f ms %8.1f GB/s %.3f  items=%d' % (r['kernel'][:70], r['ms'], r['achieved_gbs'], r['frac_of_hbm_peak'], r['items']))
"; tail -3 gpurun_out/bench_convert.err
timeout 600 python bench.py --workload distances > gpurun_out/bench_distances.json 2> gpurun_out/bench_distances.err
python -c "
import json
d=json.load(open('gpurun_out/bench_distances.json'))
for r in d['kernels']: print('%-70s %10.3f ms %8.1f GB/s %.3f  batch_ms=%s' % (r['kernel'][:70], r['ms'], r['achieved_gbs'], r['frac_of_hbm_peak'], r['batch_ms']))
"; tail -3 gpurun_out/bench_distances.err
timeout 900 python bench.py --workload pgdb3q --batch 1024 --steps 1 --warmup 3 > gpurun_out/bench_pgdb3q_v4.json 2> gpurun_out/bench_pgdb3q_v4.err
cut -c1-900 gpurun_out/bench_pgdb3q_v4.json; tail -3 gpurun_out/bench_pgdb3q_v4.err
