#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload convert > gpurun_out/bench_convert_ad.json 2> gpurun_out/bench_convert_ad.err
tail -3 gpurun_out/bench_convert_ad.err
timeout 600 python bench.py --workload distances > gpurun_out/bench_distances_ad.json 2> gpurun_out/bench_distances_ad.err
tail -3 gpurun_out/bench_distances_ad.err
python - <<'PY'
import json
for f in ('convert', 'distances'):
    r = json.loads(open(f'gpurun_out/bench_{f}_ad.json').read().strip().splitlines()[-1])
    print(json.dumps(r['cpu_baseline'])[:1500])
PY
