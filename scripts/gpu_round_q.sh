#!/bin/bash
# "next" rows: kernel table + ncu launch list / traffic of the moments kernel
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload next > gpurun_out/bench_next_q.json 2> gpurun_out/bench_next_q.err
tail -3 gpurun_out/bench_next_q.err
python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_next_q.json').read().strip().splitlines()[-1])
for row in r['kernels']:
    print(f"{row['kernel'][:75]:75s} {row['ms']:9.3f} ms {row['achieved_gbs']:8.1f} GB/s {row['frac_of_hbm_peak']:.3f}")
PY
