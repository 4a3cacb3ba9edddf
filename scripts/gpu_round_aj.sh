#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_convert.py -m gpu -x -q -k "not large" ) > gpurun_out/pytest_aj.log 2>&1
tail -3 gpurun_out/pytest_aj.log
timeout 600 python bench.py --workload convert --no-cpu-baseline > gpurun_out/bench_convert_aj.json 2> gpurun_out/bench_convert_aj.err
python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_convert_aj.json').read().strip().splitlines()[-1])
for row in r['kernels']:
    if 'n=1' in row['kernel'] or 'n=2' in row['kernel'] or 'n=3' in row['kernel']:
        if 'kraus' in row['kernel'] and 'choi2kraus' in row['kernel']: continue
        print(f"{row['kernel'][:40]:40s} {row['ms']:9.3f} ms {row['frac_of_hbm_peak']:.3f}")
PY
