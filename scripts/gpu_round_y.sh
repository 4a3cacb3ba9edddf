#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_y.json 2> gpurun_out/bench_y.err
python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_y.json').read().strip().splitlines()[-1])
print(r['value'], r['e2e']['value'], r.get('parity'), r['cpu_baseline']['value'])
PY
tail -3 gpurun_out/bench_y.err
timeout 600 python bench.py --workload pgdb2q > gpurun_out/bench_pgdb2q_y.json 2> gpurun_out/bench_pgdb2q_y.err
python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_pgdb2q_y.json').read().strip().splitlines()[-1])
print(r['value'], r['e2e']['value'], r.get('parity'), r.get('cpu_baseline'))
PY
tail -3 gpurun_out/bench_pgdb2q_y.err
