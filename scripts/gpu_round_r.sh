#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_moments.py -m gpu -x -q > gpurun_out/pytest_r2.log 2>&1; tail -5 gpurun_out/pytest_r2.log
timeout 600 python bench.py --workload next > gpurun_out/bench_next_r.json 2> gpurun_out/bench_next_r.err
tail -3 gpurun_out/bench_next_r.err
python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_next_r.json').read().strip().splitlines()[-1])
for row in r['kernels']:
    print(f"{row['kernel'][:75]:75s} {row['ms']:9.3f} ms {row['achieved_gbs']:8.1f} GB/s {row['frac_of_hbm_peak']:.3f}")
PY
