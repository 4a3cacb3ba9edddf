"""One pgdb_kernel<3> launch over `items` experiments (one wave of the persistent grid) -- the ncu target.
usage: python scripts/prof_pgdb3.py [items=148] [eigh_rel_tol]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from forest_benchmarking_b200 import synthetic as sy, tomography as tm
items = int(sys.argv[1]) if len(sys.argv) > 1 else 148
tol = float(sys.argv[2]) if len(sys.argv) > 2 else None
codes, pidx, ex, cnt, _ = sy.process_tomography_batch(3003, items, 3)
plan = tm.PgdbPlan(3, codes, pidx)
e, c = torch.from_numpy(ex).cuda(), torch.from_numpy(cnt).cuda()
for _ in range(2):
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    out, ctr = tm.pgdb_process_estimate_batch(plan, e, c, return_counters=True, eigh_rel_tol=tol)
    t1.record()
    torch.cuda.synchronize()
ctr = ctr.cpu().numpy()
print(f"items {items} ms {t0.elapsed_time(t1):.1f} eigh/item {ctr[:, 2].mean():.1f} sweeps/eigh {ctr[:, 3].sum() / ctr[:, 2].sum():.3f}")
