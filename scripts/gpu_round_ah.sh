#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload convert > gpurun_out/bench_convert_ah.json 2> gpurun_out/bench_convert_ah.err
tail -3 gpurun_out/bench_convert_ah.err
python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_convert_ah.json').read().strip().splitlines()[-1])
for row in r['kernels']:
    print(f"{row['kernel'][:40]:40s} items {row['items']:8d} {row['ms']:9.3f} ms {row['frac_of_hbm_peak']:.3f} batch_ms {row.get('batch_ms')}")
PY
