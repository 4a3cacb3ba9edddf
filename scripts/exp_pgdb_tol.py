"""3-qubit PGDB: estimate / trip counts vs the eigensolver stop threshold (with the first-order correction).
Reference run: tight threshold (0).  usage: python scripts/exp_pgdb_tol.py [items=296]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from forest_benchmarking_b200 import synthetic as sy, tomography as tm
items = int(sys.argv[1]) if len(sys.argv) > 1 else 296
codes, pidx, ex, cnt, _ = sy.process_tomography_batch(3003, items, 3)
plan = tm.PgdbPlan(3, codes, pidx)
e, c = torch.from_numpy(ex).cuda(), torch.from_numpy(cnt).cuda()
def run(tol):
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    out, ctr = tm.pgdb_process_estimate_batch(plan, e, c, return_counters=True, eigh_rel_tol=tol)
    t1.record(); torch.cuda.synchronize()
    return out.cpu().numpy(), ctr.cpu().numpy(), t0.elapsed_time(t1)
ref, cref, _ = run(0.0)
rows = []
for tol in (1e-8, 1e-6, 1e-5, 1.5e-5, 2e-5, 3e-5, 5e-5):
    run(tol)
    out, ctr, ms = run(tol)
    err = np.linalg.norm((out - ref).reshape(items, -1), axis=1) / np.linalg.norm(ref.reshape(items, -1), axis=1)
    rows.append({"tol": tol, "ms": ms, "sweeps_per_eigh": float(ctr[:, 3].sum() / ctr[:, 2].sum()),
                 "max_rel_err_vs_tight": float(err.max()), "median_rel_err": float(np.median(err)),
                 "items_outer_mismatch": int((ctr[:, 0] != cref[:, 0]).sum()),
                 "items_eigh_mismatch": int((ctr[:, 2] != cref[:, 2]).sum()),
                 "items_cost_mismatch": int((ctr[:, 1] != cref[:, 1]).sum())})
    print(rows[-1], flush=True)
print(json.dumps({"items": items, "rows": rows}))
