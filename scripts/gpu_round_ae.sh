#!/bin/bash
mkdir -p gpurun_out
for t in 4096 2048 1024 512; do
QT_TP_TILE=$t python - <<'PY'
import os, sys, torch
sys.path.insert(0, '.')
import bench_kernels as bk
from forest_benchmarking_b200.operator_tools import project_superoperators as pj
res = []
for n in (1, 2, 3):
    m = 4 ** n
    b = (2 << 30) // (2 * 16 * m * m)
    c = bk._rand_c128(torch, (b, m, m), 21 + n)
    out = torch.empty_like(c)
    ms = bk._time(torch, lambda: pj.proj_choi_to_trace_preserving_batch(c, out=out))
    res.append("n=%d %.3f" % (n, b * 32 * m * m / (ms * 1e-3) / 1e9 / 6650))
    del c, out
print("tile", os.environ["QT_TP_TILE"], res)
PY
done
