import sys, torch
sys.path.insert(0, ".")
import bench_kernels as bk
from forest_benchmarking_b200 import distance_measures as dm
rho, sig = bk._rand_states(torch, 1 << 16, 16, 61), bk._rand_states(torch, 1 << 16, 16, 62)
o = torch.empty((1 << 16,), dtype=torch.float64, device="cuda")
for _ in range(3):
    dm.fidelity_batch(rho, sig, out=o)
torch.cuda.synchronize()
