"""ncu driver for the round's newer kernels: moments (SWAR, 2 qubits), fidelity, 2-qubit PGDB."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import bench_kernels as bk
from forest_benchmarking_b200 import distance_measures as dm, observable_estimation as oe, tomography as tm, synthetic as sy
which = sys.argv[1]
if which == "moments":
    g = torch.Generator(device="cuda").manual_seed(7)
    b = (1 << 30) // 2000
    bits = torch.randint(0, 2, (b, 1000, 2), device="cuda", generator=g, dtype=torch.uint8)
    masks = torch.randint(1, 4, (b,), device="cuda", generator=g, dtype=torch.int32)
    for _ in range(3):
        oe.shots_to_obs_moments_batch(bits, masks)
elif which == "fidelity":
    rho, sig = bk._rand_states(torch, 1 << 16, 16, 61), bk._rand_states(torch, 1 << 16, 16, 62)
    o = torch.empty((1 << 16,), dtype=torch.float64, device="cuda")
    for _ in range(3):
        dm.fidelity_batch(rho, sig, out=o)
else:
    codes, pidx, ex, cnt, _ = sy.process_tomography_batch(5, 1024, 2)
    plan = tm.PgdbPlan(2, codes, pidx)
    e, c = torch.from_numpy(np.ascontiguousarray(ex)).cuda(), torch.from_numpy(np.ascontiguousarray(cnt)).cuda()
    for _ in range(2):
        tm.pgdb_process_estimate_batch(plan, e, c)
torch.cuda.synchronize()
