#!/bin/bash
# round 2: compute-sanitizer passes (memcheck, racecheck, synccheck) over small invocations of the kernels added this round
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r2san_build.log 2>&1
cat > /tmp/san_r2.py <<'PY'
import sys, os, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
heavy = os.environ.get("SAN_HEAVY", "1") == "1"
from oracle import ref_numpy as orc
from forest_benchmarking_b200 import tomography as tm, synthetic as sy, distance_measures as dm
from forest_benchmarking_b200.operator_tools import superoperator_transformations as st, project_superoperators as pj
rng = np.random.default_rng(5)
# fidelity: tridiagonal + QL kernel (d = 4, 8, 16), partial warps, a flagged (pure rho) item in the batch
for n, B in ((2, 70), (3, 37), (4, 35)):
    d = 2 ** n
    rho = np.stack([orc.ginibre_state(rng, d, rank=(1 if b == 3 else None)) for b in range(B)])
    sig = np.stack([orc.ginibre_state(rng, d, rank=(1 if b % 7 == 0 else None)) for b in range(B)])
    f = dm.fidelity_batch(torch.from_numpy(rho).cuda(), torch.from_numpy(sig).cuda())
    assert torch.isfinite(f).all()
# choi2kraus n = 4: certified low-rank path and the general one-sided Jacobi (a full-rank item)
d = 16
k = torch.randn(2, 2, d, d, dtype=torch.complex128, device="cuda")
c = st.kraus2choi_batch(k)
if heavy:
    c[1] += 0.01 * torch.eye(256, dtype=torch.complex128, device="cuda")
st.choi2kraus_batch(c)
# Choi projections n = 4 on the global-memory eigensolver
if heavy:
    c1 = c[:1].contiguous()
    pj.proj_choi_to_completely_positive_batch(c1 - 0.01)
    pj.proj_choi_to_trace_preserving_batch(c1)
    pj.proj_choi_to_trace_non_increasing_batch(c1)
# second-generation Dykstra loop: proj_physical n = 3 and one 3-qubit PGDB item (few outer steps are enough for the tools)
c3 = st.kraus2choi_batch(torch.randn(1, 2, 8, 8, dtype=torch.complex128, device="cuda")) / 8
pj.proj_choi_to_physical_batch(c3 + 0.02)
for nq in ((2, 3) if heavy else (2,)):  # one 3-qubit reconstruction is ~0.2 s natively: memcheck / synccheck only
    codes, pi, x, cn, _ = sy.process_tomography_batch(3, 1, nq)
    p = tm.PgdbPlan(nq, codes, pi)
    tm.pgdb_process_estimate_batch(p, torch.from_numpy(x).cuda(), torch.from_numpy(cn).cuda())
# superop <-> PTM: fused two-pass kernel (n = 4, needs B - 1 > 12), dense FP64-MMA variant (n = 2, 3)
s4 = torch.randn(15, 256, 256, dtype=torch.complex128, device="cuda")
pl = st.superop2pauli_liouville_batch(s4)
st.pauli_liouville2superop_batch(pl)
for n in (2, 3):
    s = torch.randn(5, 4 ** n, 4 ** n, dtype=torch.complex128, device="cuda")
    st.pauli_liouville2superop_batch(st.superop2pauli_liouville_batch(s, variant="dense_mma"), variant="dense_mma")
    st.pauli_liouville2superop_batch(st.superop2pauli_liouville_batch(s))
# 3-qubit MLE warp kernel
pidx, ex, cnt, _ = sy.state_tomography_batch(1, 3, 3)
tm.iterative_mle_state_estimate_batch(tm.MlePlan(3, pidx), torch.from_numpy(ex).cuda(), maxiter=20)
torch.cuda.synchronize()
print("sanitizer workload done")
PY
for tool in memcheck synccheck racecheck; do
  heavy=1; [ $tool = racecheck ] && heavy=0
  SAN_HEAVY=$heavy timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_r2.py > gpurun_out/r2san_$tool.log 2>&1
  echo "== $tool rc=$?"; tail -4 gpurun_out/r2san_$tool.log
done
