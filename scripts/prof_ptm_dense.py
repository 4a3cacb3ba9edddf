"""superop -> PTM: the dense FP64-MMA (DMMA) variant vs the Kronecker-factored butterfly kernels, n = 2, 3.
Prints one JSON object (timings with CUDA events, inputs larger than L2).  usage: python scripts/prof_ptm_dense.py [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from forest_benchmarking_b200.operator_tools import superoperator_transformations as st

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
peak = 6554.2
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
out = {"hbm_peak_gbs": peak, "rows": []}
for n, batch in ((2, 1 << 20), (3, 16384)):
    m = 4 ** n
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.complex(torch.randn((batch, m, m), dtype=torch.float64, device="cuda", generator=g),
                      torch.randn((batch, m, m), dtype=torch.float64, device="cuda", generator=g))
    y = torch.empty_like(x)
    for variant in ("butterfly", "dense_mma"):
        for fwd in (True, False):
            fn = st.superop2pauli_liouville_batch if fwd else st.pauli_liouville2superop_batch
            for _ in range(2):
                fn(x, out=y, variant=variant)
            torch.cuda.synchronize()
            ms = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(x, out=y, variant=variant); e1.record(); torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            t = sorted(ms)[len(ms) // 2]
            d = 2 ** n
            out["rows"].append({"n": n, "batch": batch, "variant": variant, "direction": "superop2pl" if fwd else "pl2superop",
                                "ms": t, "matrices_per_s": batch / (t * 1e-3), "hbm_gbs": batch * 32 * m * m / (t * 1e-3) / 1e9,
                                "frac_of_hbm_peak": batch * 32 * m * m / (t * 1e-3) / 1e9 / peak,
                                "dense_tflops": (batch * 16.0 * d ** 6 / (t * 1e-3) / 1e12) if variant == "dense_mma" else None})
    del x, y
print(json.dumps(out))
