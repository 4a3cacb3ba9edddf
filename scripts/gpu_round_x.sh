#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload convert > gpurun_out/bench_convert_x.json 2> gpurun_out/bench_convert_x.err
tail -2 gpurun_out/bench_convert_x.err
python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_convert_x.json').read().strip().splitlines()[-1])
for row in r['kernels']:
    if 'choi2kraus' in row['kernel']:
        print(row['kernel'][:60], row['items'], row['ms'])
PY
