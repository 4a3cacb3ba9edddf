#!/bin/bash
mkdir -p gpurun_out
( time timeout 420 python bench.py --workload pgdb3q --steps 2 --warmup 3 --cpu-baseline ) > gpurun_out/bench_pgdb3q_parity.json 2> gpurun_out/bench_pgdb3q_parity.err
tail -5 gpurun_out/bench_pgdb3q_parity.err
python - <<'PY'
import json
try:
    r = json.loads(open('gpurun_out/bench_pgdb3q_parity.json').read().strip().splitlines()[-1])
    print(r['value'], r.get('parity'), r.get('cpu_baseline'))
except Exception as e:
    print("no json:", e)
PY
