#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_u.log 2>&1
tail -4 gpurun_out/pytest_u.log
timeout 300 python bench.py --workload distances > gpurun_out/bench_distances_u.json 2> gpurun_out/bench_distances_u.err
cut -c1-330 gpurun_out/bench_distances_u.json
timeout 600 python bench.py --workload next > gpurun_out/bench_next_u.json 2> gpurun_out/bench_next_u.err
python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_next_u.json').read().strip().splitlines()[-1])
for row in r['kernels'][:7]:
    print(f"{row['kernel'][:75]:75s} {row['ms']:9.3f} ms {row['achieved_gbs']:8.1f} GB/s {row['frac_of_hbm_peak']:.3f}")
PY
timeout 300 python bench.py --workload pgdb2q --batch 8192 > gpurun_out/bench_pgdb2q_b8192.json 2> gpurun_out/bench_pgdb2q_b8192.err
cut -c1-200 gpurun_out/bench_pgdb2q_b8192.json
