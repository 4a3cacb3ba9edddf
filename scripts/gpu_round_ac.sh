#!/bin/bash
set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_convert.py -m gpu -x -q -k "large" --durations=3 ) > gpurun_out/pytest_ac.log 2>&1
tail -8 gpurun_out/pytest_ac.log
python - <<'PY'
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import bench_kernels as bk
from forest_benchmarking_b200.operator_tools import superoperator_transformations as st
for n, b in ((4, 148), (4, 592)):
    d = 2 ** n
    kraus = bk._rand_c128(torch, (b, 2, d, d), 50 + n) * (1.0 / np.sqrt(2 * d))
    a = st.kraus2choi_batch(kraus)
    ms = bk._time(torch, lambda: st.choi2kraus_batch(a), reps=1, warmup=1)
    print("choi2kraus n=%d B=%d: %.2f ms (%.3f ms per matrix)" % (n, b, ms, ms / b))
PY
