#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_g.log 2>&1
tail -12 gpurun_out/pytest_g.log
timeout 900 python bench.py --workload convert > gpurun_out/bench_convert_v2.json 2> gpurun_out/bench_convert_v2.err
python -c "
import json
d=json.load(open('gpurun_out/bench_convert_v2.json'))
for r in d['kernels']: print('%-70s %10.3f ms %8.1f GB/s %.3f  items=%d' % (r['kernel'][:70], r['ms'], r['achieved_gbs'], r['frac_of_hbm_peak'], r['items']))
"; tail -3 gpurun_out/bench_convert_v2.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_mle2q_v3.json 2> gpurun_out/bench_mle2q_v3.err
cut -c1-330 gpurun_out/bench_mle2q_v3.json; tail -3 gpurun_out/bench_mle2q_v3.err
