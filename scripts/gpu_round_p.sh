#!/bin/bash
# ring eigensolver on one warp (M = 16 / 8): full GPU suite, then the benches it touches
set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_p.log 2>&1
tail -8 gpurun_out/pytest_p.log
timeout 300 python bench.py --workload distances > gpurun_out/bench_distances_p.json 2> gpurun_out/bench_distances_p.err
cut -c1-900 gpurun_out/bench_distances_p.json
timeout 300 python bench.py --workload pgdb2q > gpurun_out/bench_pgdb2q_p.json 2> gpurun_out/bench_pgdb2q_p.err
cut -c1-700 gpurun_out/bench_pgdb2q_p.json
timeout 300 python bench.py --workload pgdb3q > gpurun_out/bench_pgdb3q_p.json 2> gpurun_out/bench_pgdb3q_p.err
cut -c1-400 gpurun_out/bench_pgdb3q_p.json
timeout 300 python bench.py --workload convert > gpurun_out/bench_convert_p.json 2> gpurun_out/bench_convert_p.err
python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_convert_p.json').read().strip().splitlines()[-1])
for row in r.get('kernels', r.get('rows', [])):
    if 'choi2kraus' in row['kernel']:
        print(row['kernel'][:30], row['ms'])
PY
