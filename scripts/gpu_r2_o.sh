#!/bin/bash
# round 2, call O: certified low-rank fast path of choi2kraus at n = 4, 5: parity + throughput on the BASELINE inputs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2o_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_convert.py -m gpu -x -q -k "choi2kraus" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2o_pytest.log
python - <<PY
import torch, time, sys, numpy as np
sys.path.insert(0, ".")
from forest_benchmarking_b200.operator_tools import superoperator_transformations as st
import bench_kernels as bk
for n, batch in ((4, 2048), (5, 128)):
    d, m = 2 ** n, 4 ** n
    kraus = bk._rand_c128(torch, (batch, 2, d, d), 50 + n) * (1.0 / np.sqrt(2 * d))
    choi = st.kraus2choi_batch(kraus)
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        k, c, e = st.choi2kraus_batch(choi)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    back = st.kraus2choi_batch(k[:, :2].contiguous())
    err = float((back - choi).flatten(1).norm(dim=1).max() / choi.flatten(1).norm(dim=1).max())
    print(f"n={n} batch={batch}: {dt*1e3:.2f} ms  {batch/dt:.0f} matrices/s  counts {c.min().item()}..{c.max().item()}  round-trip err {err:.2e}")
PY
