#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_state.py -m gpu -x -q -k "purity or distances or variance" ) > gpurun_out/pytest_ai.log 2>&1
tail -3 gpurun_out/pytest_ai.log
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench_kernels as bk
from forest_benchmarking_b200 import distance_measures as dm
for n in (1, 2, 3, 4, 5):
    d = 2 ** n
    b = (1 << 30) // (16 * d * d)
    r = bk._rand_states(torch, b, d, 31 + n)
    o = torch.empty((b,), dtype=torch.float64, device="cuda")
    ms = bk._time(torch, lambda: dm.purity_batch(r, out=o))
    print("purity d=%d %.3f" % (d, b * (16 * d * d + 8) / (ms * 1e-3) / 1e9 / 6650))
    del r, o
PY
