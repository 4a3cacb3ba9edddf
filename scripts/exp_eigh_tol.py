"""Experiment: parity and speed of pgdb_process_estimate vs the Jacobi stopping tolerance."""
import ctypes, sys, time
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from forest_benchmarking_b200 import _lib, tomography as tm, synthetic as sy
from util import golden, max_relerr

lib = _lib.lib()
codes3, pidx3, ex3, cnt3, _ = sy.process_tomography_batch(3003, 148, 3)
plan3 = tm.PgdbPlan(3, codes3, pidx3)
e3, c3 = torch.from_numpy(ex3).cuda(), torch.from_numpy(cnt3).cuda()
base = None
for tol in (0.0, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4):
    _lib.check(lib.qt_set_eigh_tolerance(ctypes.c_double(tol)), "tol")
    line = [f"tol={tol:g}"]
    for name in ("pgdb_3q_pauli", "pgdb_3q_sic", "pgdb_2q_pauli", "pgdb_2q_sic_mixed", "pgdb_1q_pauli"):
        g = golden(name)
        n = int(g["n"])
        plan = tm.PgdbPlan(n, g["state_codes"], g["pauli_idx"])
        choi, cn = tm.pgdb_process_estimate_batch(plan, torch.from_numpy(np.ascontiguousarray(g["expectations"])).cuda(),
                                                  torch.from_numpy(np.ascontiguousarray(g["counts"])).cuda(),
                                                  bool(g["trace_preserving"]), return_counters=True)
        cn = cn.cpu().numpy()
        line.append(f"{name}: err={max_relerr(choi.cpu().numpy(), g['choi_ref']):.1e} d_eigh={int(np.abs(cn[:,2]-g['counters_ref'][:,0]).max())}"
                    f" d_cost={int(np.abs(cn[:,1]-g['counters_ref'][:,1]).max())}")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out, cn = tm.pgdb_process_estimate_batch(plan3, e3, c3, True, return_counters=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    cn = cn.cpu().numpy()
    if base is None:
        base = out.cpu().numpy()
    line.append(f"3q batch148: {dt*1e3:.0f} ms sweeps/eigh={cn[:,3].sum()/cn[:,2].sum():.2f} eigh={cn[:,2].mean():.1f} "
                f"err_vs_default={max_relerr(out.cpu().numpy(), base):.1e}")
    print(" | ".join(line), flush=True)
