#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_state.py -m gpu -x -q ) > gpurun_out/pytest_aa.log 2>&1
tail -4 gpurun_out/pytest_aa.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_aa.json 2> gpurun_out/bench_aa.err
python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_aa.json').read().strip().splitlines()[-1])
print(r['value'], r['ms_per_step'], r['e2e']['value'], r['roofline']['frac'])
PY
tail -3 gpurun_out/bench_aa.err
