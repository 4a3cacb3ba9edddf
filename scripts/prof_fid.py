"""fidelity_batch on B random full-rank n-qubit pairs (profiling driver).  usage: python scripts/prof_fid.py [n] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_kernels as bk
from forest_benchmarking_b200 import distance_measures as dm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 12 * 32 * 4
d = 2 ** n
rho, sig = bk._rand_states(torch, B, d, 1), bk._rand_states(torch, B, d, 2)
out = torch.empty(B, dtype=torch.float64, device="cuda")
for _ in range(3):
    dm.fidelity_batch(rho, sig, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
dm.fidelity_batch(rho, sig, out=out)
e1.record()
torch.cuda.synchronize()
print(f"n={n} B={B}: {e0.elapsed_time(e1):.3f} ms")
