#!/bin/bash
# compute-sanitizer passes over small invocations of every kernel family (memcheck, racecheck, synccheck)
set -x
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from forest_benchmarking_b200 import tomography as tm, synthetic as sy, distance_measures as dm
from forest_benchmarking_b200.operator_tools import superoperator_transformations as st, project_superoperators as pj
from forest_benchmarking_b200.operator_tools import project_state_matrix as psm
pidx, ex, cnt, _ = sy.state_tomography_batch(1, 13, 2)
plan = tm.MlePlan(2, pidx)
e = torch.from_numpy(ex).cuda()
for k in (1, 2, 3):
    tm.iterative_mle_state_estimate_batch(plan, e, kernel=k, maxiter=60)
tm.linear_inv_state_estimate_batch(plan, e)
for n, b in ((1, 5), (2, 2)):
    codes, pi, x, c, _ = sy.process_tomography_batch(3, b, n, in_basis="sic")
    p = tm.PgdbPlan(n, codes, pi)
    out = tm.pgdb_process_estimate_batch(p, torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda())
    pj.proj_choi_to_physical_batch(out + 0.05)
    pj.proj_choi_to_completely_positive_batch(out - 0.01)
    pj.proj_choi_to_trace_preserving_batch(out)
    st.choi2kraus_batch(out)
# one 64x64 eigendecomposition through the block Jacobi (mbarrier split-phase path) -- a full 3-qubit PGDB run is
# hours under racecheck
c3 = torch.randn(1, 64, 64, dtype=torch.complex128, device="cuda")
pj.proj_choi_to_completely_positive_batch(c3)
for n in (1, 2, 3, 4):
    d = 2 ** n
    k = torch.randn(3, 2, d, d, dtype=torch.complex128, device="cuda")
    c = st.kraus2choi_batch(k)
    s = st.reshuffle_batch(c)
    pl = st.superop2pauli_liouville_batch(s)
    st.pauli_liouville2superop_batch(pl)
    r = c[:, :d, :d].contiguous()
    dm.fidelity_batch(r, r); dm.trace_distance_batch(r, r); psm.project_state_matrix_to_physical_batch(r)
torch.cuda.synchronize()
print("sanitizer workload done")
PY
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_small.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/sanitizer_$tool.log
done
