#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_project.py -m gpu -x -q ) > gpurun_out/pytest_ag.log 2>&1
tail -3 gpurun_out/pytest_ag.log
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench_kernels as bk
from forest_benchmarking_b200.operator_tools import project_superoperators as pj
for n in (1, 2, 3):
    m = 4 ** n
    b = (2 << 30) // (2 * 16 * m * m)
    c = bk._rand_c128(torch, (b, m, m), 21 + n)
    out = torch.empty_like(c)
    ms = bk._time(torch, lambda: pj.proj_choi_to_trace_non_increasing_batch(c, out=out))
    print("TNI n=%d %.3f" % (n, b * 32 * m * m / (ms * 1e-3) / 1e9 / 6650))
    del c, out
PY
