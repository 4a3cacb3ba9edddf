"""Generate tests/golden/*.npz by running the UNMODIFIED reference (through oracle/pyquil_shim)
on seeded synthetic inputs.  Run here (the reference tree does not travel to the GPU box):

    python oracle/make_golden.py [--skip-3q]

TEST INFRASTRUCTURE ONLY.  Counters (iterations, eigh calls, cost evaluations) are obtained by
wrapping reference functions at run time -- the reference source is not modified.
"""
import argparse
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_numpy as orc          # noqa: E402
from oracle import reference_bridge as rb    # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class CallCounter:
    def __init__(self, module, name):
        self.module, self.name, self.n = module, name, 0
        self.orig = getattr(module, name)

    def __enter__(self):
        def wrapped(*a, **k):
            self.n += 1
            return self.orig(*a, **k)
        setattr(self.module, self.name, wrapped)
        return self

    def __exit__(self, *exc):
        setattr(self.module, self.name, self.orig)


def golden_mle(ref, name, seed, batch, n, **kw):
    truth, pidx, ex, cnt = orc.synth_state_tomography(seed, batch, n)
    qubits = list(range(n))
    coeffs = np.ones(len(pidx))
    rho = np.empty_like(truth)
    iters = np.empty(batch, dtype=np.int32)
    maxiter = kw.get("maxiter", 10_000)
    t0 = time.time()
    for b in range(batch):
        res = rb.state_results(ref, pidx, coeffs, ex[b], cnt[b], qubits)
        with CallCounter(ref.tomo, "_R") as c, warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            rho[b] = ref.tomo.iterative_mle_state_estimate(res, qubits, **kw)
        iters[b] = maxiter if len(w) else c.n
    print(f"{name}: {batch} items in {time.time() - t0:.1f}s iters={iters.tolist()}")
    np.savez_compressed(os.path.join(OUT, name), seed=seed, n=n, pauli_idx=pidx, expectations=ex,
                        counts=cnt, rho_true=truth, rho_ref=rho, iters_ref=iters,
                        kwargs=np.array(repr(kw)))


def golden_pgdb(ref, name, seed, batch, n, basis, tp=True, unitary=True):
    import forest.benchmarking.operator_tools.project_superoperators as ps
    truth, settings, ex, cnt = orc.synth_process_tomography(seed, batch, n, in_basis=basis, unitary=unitary)
    qubits = list(range(n))
    coeffs = np.ones(len(settings))
    est = np.empty_like(truth)
    counters = np.empty((batch, 2), dtype=np.int32)
    t0 = time.time()
    for b in range(batch):
        res = rb.process_results(ref, settings, coeffs, ex[b], cnt[b], qubits)
        with CallCounter(ps, "proj_choi_to_completely_positive") as ce, CallCounter(ref.tomo, "_cost") as cc:
            est[b] = ref.tomo.pgdb_process_estimate(res, qubits, trace_preserving=tp)
        counters[b] = (ce.n, cc.n)
    print(f"{name}: {batch} items in {time.time() - t0:.1f}s (eigh, cost evals)={counters.tolist()}")
    np.savez_compressed(os.path.join(OUT, name), seed=seed, n=n, basis=np.array(basis), trace_preserving=tp,
                        state_codes=np.array([s for s, _ in settings], dtype=np.int32),
                        pauli_idx=np.array([k for _, k in settings], dtype=np.int32),
                        expectations=ex, counts=cnt, choi_true=truth, choi_ref=est, counters_ref=counters)


def golden_algebra(ref, name, seed, n, batch):
    rng = np.random.default_rng(seed)
    d = 2 ** n
    ot = ref.ot
    kraus = np.stack([np.stack([np.sqrt(.7) * orc.haar_unitary(rng, d), np.sqrt(.3) * orc.haar_unitary(rng, d)])
                      for _ in range(batch)])
    choi = np.stack([ot.kraus2choi(list(k)) for k in kraus])
    superop = np.stack([ot.choi2superop(c) for c in choi])
    ksup = np.stack([ot.kraus2superop(list(k)) for k in kraus])
    pl = np.stack([ot.superop2pauli_liouville(s) for s in superop])
    back = np.stack([ot.pauli_liouville2superop(p) for p in pl])
    # non-physical Hermitian-perturbed Choi matrices for the projections
    noisy = np.empty_like(choi)
    for b in range(batch):
        x = rng.standard_normal((d * d, d * d)) + 1j * rng.standard_normal((d * d, d * d))
        noisy[b] = choi[b] + (x + x.conj().T) / (4 * d * d)
    cp = np.stack([ot.proj_choi_to_completely_positive(x) for x in noisy])
    tp = np.stack([ot.proj_choi_to_trace_preserving(x) for x in noisy])
    tni = np.stack([ot.proj_choi_to_trace_non_increasing(x) for x in noisy])
    phys = np.stack([ot.proj_choi_to_physical(x) for x in noisy])
    phys_tni = np.stack([ot.proj_choi_to_physical(x, False) for x in noisy])
    np.savez_compressed(os.path.join(OUT, name), seed=seed, n=n, kraus=kraus, choi=choi, superop=superop,
                        kraus2superop=ksup, pauli_liouville=pl, pl2superop=back, noisy=noisy, proj_cp=cp,
                        proj_tp=tp, proj_tni=tni, proj_physical=phys, proj_physical_tni=phys_tni)
    print(f"{name}: done")


def golden_nonhermitian_physical(ref, name="proj_physical_nonherm"):
    """proj_choi_to_physical on NON-Hermitian inputs (the reference does not Hermitise up front: the anti-Hermitian
    part lives on in old_CP_change and enters the stopping rule, project_superoperators.py:112-136), with the
    number of CP projections each call made; plus distance-measure defaults (purity / impurity without keywords)."""
    import forest.benchmarking.operator_tools.project_superoperators as ps
    out = {}
    for n, seed, batch in ((1, 7001, 6), (2, 7002, 4), (3, 7003, 2)):
        rng = np.random.default_rng(seed)
        d = 2 ** n
        xs, ys, ts, cnt, cnt_t = [], [], [], [], []
        for b in range(batch):
            u1, u2 = orc.haar_unitary(rng, d), orc.haar_unitary(rng, d)
            choi = orc.kraus2choi([np.sqrt(.7) * u1, np.sqrt(.3) * u2])
            g = rng.standard_normal((d * d, d * d)) + 1j * rng.standard_normal((d * d, d * d))
            # growing anti-Hermitian share: from a small perturbation to one that changes the trip count
            x = choi + (g + g.conj().T) / (4 * d * d) + (0.02 * 3 ** b) * (g - g.conj().T) / (4 * d * d)
            xs.append(x)
            with CallCounter(ps, "proj_choi_to_completely_positive") as c:
                ys.append(ps.proj_choi_to_physical(x))
            cnt.append(c.n)
            with CallCounter(ps, "proj_choi_to_completely_positive") as c:
                ts.append(ps.proj_choi_to_physical(x, False))
            cnt_t.append(c.n)
        out[f"n{n}_in"], out[f"n{n}_out"], out[f"n{n}_out_tni"] = np.stack(xs), np.stack(ys), np.stack(ts)
        out[f"n{n}_calls"], out[f"n{n}_calls_tni"] = np.array(cnt, dtype=np.int32), np.array(cnt_t, dtype=np.int32)
        # the same inputs Hermitised first: what a wrapper that symmetrises up front would count
        herm = []
        for x in xs:
            with CallCounter(ps, "proj_choi_to_completely_positive") as c:
                ps.proj_choi_to_physical((x + x.conj().T) / 2)
            herm.append(c.n)
        out[f"n{n}_calls_hermitised"] = np.array(herm, dtype=np.int32)
        print(f"{name} n={n}: calls {cnt} (TNI {cnt_t}); Hermitised input would take {herm}")
    rng = np.random.default_rng(7004)
    rho = np.stack([orc.ginibre_state(rng, 4) for _ in range(4)])
    out["rho"] = rho
    out["purity_default"] = np.array([ref.dm.purity(r) for r in rho])
    out["purity_renorm"] = np.array([ref.dm.purity(r, dim_renorm=True) for r in rho])
    out["impurity_default"] = np.array([ref.dm.impurity(r) for r in rho])
    out["impurity_renorm"] = np.array([ref.dm.impurity(r, dim_renorm=True) for r in rho])
    out["fidelity_tol1e6"] = np.array([ref.dm.fidelity(rho[0], r, tol=1e6) for r in rho])
    np.savez_compressed(os.path.join(OUT, name), **out)


def golden_distances(ref, name, seed, n, batch):
    rng = np.random.default_rng(seed)
    d = 2 ** n
    rho = np.stack([orc.ginibre_state(rng, d) for _ in range(batch)])
    sigma = np.stack([orc.ginibre_state(rng, d, rank=(1 if b % 4 == 0 else None)) for b in range(batch)])
    fid = np.array([ref.dm.fidelity(r, s) for r, s in zip(rho, sigma)])
    td = np.array([ref.dm.trace_distance(r, s) for r, s in zip(rho, sigma)])
    pur = np.array([ref.dm.purity(r, dim_renorm=False) for r in rho])
    wiz = np.stack([ref.project_state_matrix_to_physical(r - 0.3 * s) for r, s in zip(rho, sigma)])
    np.savez_compressed(os.path.join(OUT, name), seed=seed, n=n, rho=rho, sigma=sigma, fidelity=fid,
                        trace_distance=td, purity=pur, wizard_in=rho - 0.3 * sigma, wizard_out=wiz)
    print(f"{name}: done")


def golden_next_rows(ref, name="next_rows"):
    """linear-inversion process estimate, closest unitary, log-likelihood, shots -> moments,
    ratio variance -- all outputs from the reference's own functions."""
    from forest.benchmarking.operator_tools.project_superoperators import proj_choi_to_unitary
    from forest.benchmarking.observable_estimation import shots_to_obs_moments, ratio_variance
    from pyquil.paulis import PauliTerm
    out = {}
    # linear_inv_process_estimate: 1 qubit Pauli + SIC, 2 qubits SIC
    for tag, seed, n, basis in (("lip_1q_pauli", 6001, 1, "pauli"), ("lip_1q_sic", 6002, 1, "sic"), ("lip_2q_sic", 6003, 2, "sic")):
        _, settings, ex, cnt = orc.synth_process_tomography(seed, 3, n, in_basis=basis)
        qubits = list(range(n))
        coeffs = np.ones(len(settings))
        out[tag + "_codes"] = np.array([s for s, _ in settings], dtype=np.int32)
        out[tag + "_pidx"] = np.array([k for _, k in settings], dtype=np.int32)
        out[tag + "_ex"] = ex
        out[tag + "_choi"] = np.stack([ref.tomo.linear_inv_process_estimate(
            rb.process_results(ref, settings, coeffs, ex[b], cnt[b], qubits), qubits) for b in range(3)])
    # proj_choi_to_unitary
    rng = np.random.default_rng(6004)
    for n in (1, 2, 3):
        d = 2 ** n
        xs = []
        for _ in range(4):
            u = orc.haar_unitary(rng, d)
            g = rng.standard_normal((d * d, d * d)) + 1j * rng.standard_normal((d * d, d * d))
            xs.append(.9 * orc.kraus2choi(u) + .1 * (g @ g.conj().T) / d ** 2 + .01 * g)
        xs = np.stack(xs)
        out[f"unitary_n{n}_in"] = xs
        out[f"unitary_n{n}_out"] = np.stack([proj_choi_to_unitary(x) for x in xs])
    # state_log_likelihood
    rho, pidx, ex, cnt = orc.synth_state_tomography(6005, 4, 2)
    qubits = [0, 1]
    out["ll_rho"], out["ll_pidx"], out["ll_ex"], out["ll_cnt"] = rho, pidx, ex, cnt
    out["ll_value"] = np.array([ref.tomo.state_log_likelihood(
        rho[b], rb.state_results(ref, pidx, np.ones(len(pidx)), ex[b], cnt[b], qubits), qubits) for b in range(4)])
    # shots_to_obs_moments (qc.run returns int64 bits) and ratio_variance
    rng = np.random.default_rng(6006)
    qubits = [4, 7, 9]
    bits = (rng.random((6, 300, 3)) < rng.uniform(.1, .9, size=(6, 1, 3))).astype(np.uint8)
    masks = np.array([1, 2, 4, 3, 6, 7], dtype=np.int32)
    coeffs = np.array([1.0, -1.0, 0.5, 2.0, 1.0, -0.25])
    mom = np.empty((2, 6, 2))
    for i in range(6):
        ops = [("Z", q) for c, q in enumerate(qubits) if (masks[i] >> c) & 1]
        term = PauliTerm.from_list(ops, coefficient=coeffs[i])
        for prior in (0, 1):
            mom[prior, i] = shots_to_obs_moments(bits[i].astype(np.int64), qubits, term, bool(prior))
    out["mom_bits"], out["mom_masks"], out["mom_coeffs"], out["mom_out"] = bits, masks, coeffs, mom
    a, va, b, vb = rng.uniform(-1, 1, 16), rng.uniform(1e-4, 1e-2, 16), rng.uniform(.7, 1, 16), rng.uniform(1e-5, 1e-3, 16)
    out["rv_in"] = np.stack([a, va, b, vb])
    out["rv_out"] = ratio_variance(a, va, b, vb)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(f"{name}: done")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-3q", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    ref = rb.load()
    jobs = {
        "algebra": lambda: [golden_algebra(ref, f"algebra_n{n}", 4004 + n, n, 4 if n < 3 else 2) for n in (1, 2, 3)],
        "dist": lambda: [golden_distances(ref, f"distances_n{n}", 5005 + n, n, 32) for n in (1, 2, 4)],
        "mle1": lambda: golden_mle(ref, "mle_1q", 1001, 8, 1),
        "mle2": lambda: golden_mle(ref, "mle_2q", 2002, 8, 2),
        "mle2_tol": lambda: golden_mle(ref, "mle_2q_tol1e-4", 2003, 8, 2, tol=1e-4),
        "mle2_maxiter": lambda: golden_mle(ref, "mle_2q_maxiter200", 2004, 4, 2, maxiter=200),
        "mle2_ent": lambda: golden_mle(ref, "mle_2q_maxent", 2005, 2, 2, entropy_penalty=.001, tol=1e-5),
        "mle2_hedge": lambda: golden_mle(ref, "mle_2q_hedged", 2006, 4, 2, epsilon=1e-4, beta=.5, tol=1e-3),
        "mle3": lambda: golden_mle(ref, "mle_3q_tol1e-5", 2007, 2, 3, tol=1e-5),
        "pgdb1": lambda: [golden_pgdb(ref, "pgdb_1q_pauli", 3001, 8, 1, "pauli"),
                          golden_pgdb(ref, "pgdb_1q_sic", 3002, 8, 1, "sic"),
                          golden_pgdb(ref, "pgdb_1q_pauli_tni", 3004, 4, 1, "pauli", tp=False),
                          golden_pgdb(ref, "pgdb_1q_pauli_mixed", 3005, 4, 1, "pauli", unitary=False)],
        "pgdb2": lambda: [golden_pgdb(ref, "pgdb_2q_pauli", 3003, 4, 2, "pauli"),
                          golden_pgdb(ref, "pgdb_2q_sic", 3006, 4, 2, "sic"),
                          golden_pgdb(ref, "pgdb_2q_sic_mixed", 3007, 2, 2, "sic", unitary=False)],
        "next": lambda: golden_next_rows(ref),
        "nonherm": lambda: golden_nonhermitian_physical(ref),
        "pgdb2_tni": lambda: golden_pgdb(ref, "pgdb_2q_pauli_tni", 3009, 2, 2, "pauli", tp=False),
        "pgdb3": lambda: [golden_pgdb(ref, "pgdb_3q_sic", 3008, 1, 3, "sic"),
                          golden_pgdb(ref, "pgdb_3q_pauli", 3003, 1, 3, "pauli")],
    }
    for key, fn in jobs.items():
        if args.only and key not in args.only.split(","):
            continue
        if args.skip_3q and key == "pgdb3":
            continue
        fn()


if __name__ == "__main__":
    main()
