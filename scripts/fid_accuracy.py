"""Accuracy + speed sweep of fidelity_batch against the oracle (dev tool, GPU): full-rank, rank-deficient sigma, pure rho
(fallback), all n; then kernel time per n at a streaming-sized batch."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import ref_numpy as orc
from forest_benchmarking_b200 import distance_measures as dm

rng = np.random.default_rng(11)
for n in (1, 2, 3, 4, 5):
    d = 2 ** n
    for kind in ("full", "sigma rank 1", "sigma rank 2", "rho rank 1", "rho near-singular"):
        rho, sig = [], []
        for b in range(200):
            if kind == "rho rank 1":
                r = orc.ginibre_state(rng, d, rank=1)
            elif kind == "rho near-singular":
                r = (1 - 1e-7) * orc.ginibre_state(rng, d, rank=1) + 1e-7 * np.eye(d) / d
            else:
                r = orc.ginibre_state(rng, d)
            s = orc.ginibre_state(rng, d, rank=1 if kind == "sigma rank 1" else (min(2, d) if kind == "sigma rank 2" else None))
            rho.append(r), sig.append(s)
        rho, sig = np.stack(rho), np.stack(sig)
        got = dm.fidelity_batch(torch.from_numpy(rho).cuda(), torch.from_numpy(sig).cuda()).cpu().numpy()
        want = np.array([np.real(orc.fidelity(r, s)) for r, s in zip(rho, sig)])
        rel = np.abs(got - want) / np.maximum(want, 1e-3)
        print(f"n={n} {kind:18s} max rel err {rel.max():.2e}  median {np.median(rel):.2e}")
for n in (2, 3, 4, 5):
    d = 2 ** n
    B = (1 << 30) // (32 * d * d)
    g = torch.Generator(device="cuda").manual_seed(n)
    def states():
        a = torch.randn(B, d, d, dtype=torch.complex128, device="cuda", generator=g) if False else torch.complex(
            torch.randn(B, d, d, dtype=torch.float64, device="cuda", generator=g),
            torch.randn(B, d, d, dtype=torch.float64, device="cuda", generator=g))
        r = a @ a.conj().transpose(1, 2)
        return r / torch.diagonal(r, dim1=1, dim2=2).sum(-1).real[:, None, None]
    rho, sig = states(), states()
    out = torch.empty(B, dtype=torch.float64, device="cuda")
    for _ in range(2):
        dm.fidelity_batch(rho, sig, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        dm.fidelity_batch(rho, sig, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"n={n} B={B}: {ms:.3f} ms, {B / ms * 1e3:.3e} pairs/s, {B * 32 * d * d / ms / 1e6:.0f} GB/s")
