"""Experiment: 2-qubit MLE throughput vs batch size for the register / quad kernels."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from forest_benchmarking_b200 import tomography as tm, synthetic as sy
pidx, ex, _, _ = sy.state_tomography_batch(2002, 4096, 2)
plan = tm.MlePlan(2, pidx)
for B in (1024, 4096, 8192, 16384, 32768, 65536, 262144):
    e = torch.from_numpy(np.tile(ex, (B // 4096 + 1, 1))[:B].copy()).cuda()
    line = [f"B={B}"]
    for name, k in (("reg", 1), ("quad", 3), ("auto", 0)):
        for _ in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            rho, it = tm.iterative_mle_state_estimate_batch(plan, e, kernel=k)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        line.append(f"{name}: {dt*1e3:.2f} ms {B/dt/1e3:.0f}k/s")
    print(" | ".join(line), flush=True)
