"""Per-kernel roofline sweeps used by bench.py --workload {streaming, convert, distances}.

Every entry times ONE kernel with CUDA events on the launching stream (inputs resident in HBM and larger
than the 126 MB L2, so no flush is needed; the input size is stated per row), and reports achieved GB/s
from the ALGORITHMIC bytes of DESIGN.md section 4 against the measured HBM peak.
"""
import ctypes
import json

import numpy as np


def _time(torch, fn, reps=5, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def _rand_c128(torch, shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    re = torch.randn(shape, dtype=torch.float64, device="cuda", generator=g)
    im = torch.randn(shape, dtype=torch.float64, device="cuda", generator=g)
    return torch.complex(re, im)


def _rand_states(torch, b, d, seed):
    """Full-rank Ginibre states G G^dagger / tr generated on the device."""
    g = _rand_c128(torch, (b, d, d), seed)
    rho = g @ g.conj().transpose(1, 2)
    tr = torch.diagonal(rho, dim1=1, dim2=2).sum(-1).real
    return (rho / tr[:, None, None]).contiguous()


def _row(name, items, bytes_item, ms, peak, extra=None):
    gbs = items * bytes_item / (ms * 1e-3) / 1e9
    r = {"kernel": name, "items": int(items), "bytes_per_item": int(bytes_item), "ms": round(ms, 4),
         "items_per_s": items / (ms * 1e-3), "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 4),
         "input_mb": round(items * bytes_item / 2 ** 20, 1)}
    if extra:
        r.update(extra)
    return r


def streaming_rows(torch, peak, budget_bytes=2 << 30):
    """HBM-roofline kernels: single R-rho-R step, TP / TNI projection, n=1 CP projection, trace distance, purity."""
    from forest_benchmarking_b200 import tomography as tm, distance_measures as dm
    from forest_benchmarking_b200.operator_tools import project_superoperators as pj
    rows = []
    for n in (1, 2):
        d, k = 2 ** n, 4 ** n - 1
        b = budget_bytes // (2 * 16 * d * d + 8 * k)
        ex = torch.rand((k, b), dtype=torch.float64, device="cuda") * 1.2 - .6
        rho = _rand_states(torch, b, d, 11 + n)
        out = torch.empty_like(rho)
        ms = _time(torch, lambda: tm.mle_step_batch(n, ex, rho, .1, out=out))
        rows.append(_row(f"mle_step_kernel<{n}> (one R rho R update)", b, 32 * d * d + 8 * k, ms, peak))
        del ex, rho, out
    for n in (1, 2, 3):
        m = 4 ** n
        b = budget_bytes // (2 * 16 * m * m)
        c = _rand_c128(torch, (b, m, m), 21 + n)
        out = torch.empty_like(c)
        ms = _time(torch, lambda: pj.proj_choi_to_trace_preserving_batch(c, out=out))
        rows.append(_row(f"proj_tp_kernel<{n}> (proj_choi_to_trace_preserving)", b, 32 * m * m, ms, peak))
        ms = _time(torch, lambda: pj.proj_choi_to_trace_non_increasing_batch(c, out=out))
        tni = "tni_fused1_kernel (one pass, thread per item)" if n == 1 else f"tni_correction_kernel<{n}> + tni_apply_kernel<{n}> (two passes)"
        rows.append(_row(f"{tni} (proj_choi_to_trace_non_increasing)", b, 32 * m * m, ms, peak))
        if n == 1:
            ms = _time(torch, lambda: pj.proj_choi_to_completely_positive_batch(c, out=out))
            rows.append(_row("proj_cp_kernel<1> (proj_choi_to_completely_positive, 4x4)", b, 32 * m * m, ms, peak))
        del c, out
    for n in (2, 4):
        d = 2 ** n
        b = budget_bytes // (2 * 16 * d * d)
        r, s = _rand_states(torch, b, d, 31 + n), _rand_states(torch, b, d, 41 + n)
        o = torch.empty((b,), dtype=torch.float64, device="cuda")
        ms = _time(torch, lambda: dm.trace_distance_batch(r, s, out=o))
        rows.append(_row(f"trace_distance_kernel<{d}>", b, 32 * d * d + 8, ms, peak))
        ms = _time(torch, lambda: dm.purity_batch(r, out=o))
        rows.append(_row(f"purity_kernel<{d}>", b, 16 * d * d + 8, ms, peak))
        oc = torch.empty((b,), dtype=torch.complex128, device="cuda")
        ms = _time(torch, lambda: dm.hilbert_schmidt_ip_batch(r, s, out=oc))
        rows.append(_row(f"hs_inner_kernel d={d} (hilbert_schmidt_ip / process_fidelity)", b, 32 * d * d + 16, ms, peak))
        del r, s, o, oc
    return rows


def convert_rows(torch, peak, batch=16384, nk=2, budget_bytes=6 << 30):
    """BASELINE configs[3]: kraus -> choi -> superop -> pauli_liouville -> superop -> choi (-> kraus, n <= 3),
    n = 1..5, `batch` matrices per n, chunked so that three representations fit the byte budget."""
    from forest_benchmarking_b200.operator_tools import superoperator_transformations as st
    rows = []
    for n in (1, 2, 3, 4, 5):
        d, m = 2 ** n, 4 ** n
        chunk = int(min(batch, max(1, budget_bytes // (3 * 16 * m * m))))
        # timing rule: inputs larger than the 126 MB L2 -- at n <= 2 the nominal batch is only 4 / 64 MB per
        # representation, so the kernels are timed on >= 256 MB and `batch_ms` is scaled back to the nominal batch
        chunk = max(chunk, (256 << 20) // (16 * m * m)) if n <= 2 else chunk
        kraus = _rand_c128(torch, (chunk, nk, d, d), 50 + n) * (1.0 / np.sqrt(nk * d))
        a, b_, ws = (torch.empty((chunk, m, m), dtype=torch.complex128, device="cuda") for _ in range(3))
        reps = batch / chunk if chunk > batch else max(1, batch // chunk)
        lib_calls = [
            ("kraus2choi", lambda: st.kraus2choi_batch(kraus), 16 * (nk * d * d + m * m)),
            ("choi2superop", lambda: st.reshuffle_batch(a, out=b_), 32 * m * m),
            ("superop2pauli_liouville", lambda: st.superop2pauli_liouville_batch(b_, out=a, workspace=ws), 32 * m * m),
            ("pauli_liouville2superop", lambda: st.pauli_liouville2superop_batch(a, out=b_, workspace=ws), 32 * m * m),
            ("superop2choi", lambda: st.reshuffle_batch(b_, out=a), 32 * m * m),
        ]
        a.copy_(st.kraus2choi_batch(kraus))
        for name, fn, bytes_item in lib_calls:
            ms = _time(torch, fn, reps=3 if n >= 4 else 5, warmup=2)
            rows.append(_row(f"{name} n={n}", chunk, bytes_item, ms, peak,
                             {"chunks_for_batch": reps, "batch_ms": round(ms * reps, 3)}))
        # choi2kraus: n <= 3 shared-memory eigensolver (FP64-bound); n = 4, 5: certified low-rank fast path (three scans of the
        # matrix) for these 2-Kraus-operator channels, general one-sided Jacobi solver for anything it cannot certify
        ms = _time(torch, lambda: st.choi2kraus_batch(a), reps=3, warmup=1)
        label = "eigensolver: FP64-bound, GB/s for reference only" if n <= 2 else "rank-2 inputs: certified low-rank path"
        rows.append(_row(f"choi2kraus n={n} ({label})", chunk, 16 * m * m + 16 * m * m + 8 * m, ms, peak,
                         {"chunks_for_batch": reps, "batch_ms": round(ms * reps, 3)}))
        if n == 4:
            g = _rand_c128(torch, (148, m, m), 99)
            full = (a[:148] + 1e-3 * (g + g.conj().transpose(1, 2))).contiguous()
            ms = _time(torch, lambda: st.choi2kraus_batch(full), reps=1, warmup=1)
            rows.append(_row("choi2kraus n=4, FULL-RANK inputs (general one-sided Jacobi out of L2, one matrix per SM, 148 matrices)",
                             148, 16 * m * m + 16 * m * m + 8 * m, ms, peak, {"chunks_for_batch": 1, "batch_ms": round(ms, 3)}))
            del g, full
        del kraus, a, b_, ws
        torch.cuda.empty_cache()
    return rows


def distance_rows(torch, peak, pairs, n=4):
    """BASELINE configs[4]: fidelity + trace_distance over `pairs` random n-qubit state pairs on this rank."""
    from forest_benchmarking_b200 import distance_measures as dm
    d = 2 ** n
    chunk = min(pairs, 1 << 18)
    rho, sig = _rand_states(torch, chunk, d, 61), _rand_states(torch, chunk, d, 62)
    o = torch.empty((chunk,), dtype=torch.float64, device="cuda")
    reps = -(-pairs // chunk)
    rows = []
    ms = _time(torch, lambda: dm.fidelity_batch(rho, sig, out=o))
    rows.append(_row(f"fidelity_kernel<{d}> (Cholesky + one values-only warp-Jacobi eigh per pair; FP64/latency-bound)", chunk,
                     32 * d * d + 8, ms, peak, {"batch_ms": round(ms * reps, 3), "pairs_total": pairs}))
    ms = _time(torch, lambda: dm.trace_distance_batch(rho, sig, out=o))
    rows.append(_row(f"trace_distance_kernel<{d}>", chunk, 32 * d * d + 8, ms, peak,
                     {"batch_ms": round(ms * reps, 3), "pairs_total": pairs}))
    return rows


def next_rows(torch, peak, budget_bytes=4 << 30):
    """SURVEY 8(f) "next" rows: raw shots -> moments (HBM-bound byte kernel), log-likelihood, linear-inversion process
    estimate, closest unitary.  GB/s from the algorithmic bytes (inputs read once + outputs written once)."""
    from forest_benchmarking_b200 import observable_estimation as oe, tomography as tm, synthetic as sy
    from forest_benchmarking_b200.operator_tools import project_superoperators as pj
    rows = []
    g = torch.Generator(device="cuda").manual_seed(7)
    for q, shots in ((1, 1000), (2, 1000), (2, 500), (4, 1000), (8, 1000), (3, 1000), (5, 1000), (7, 1000), (11, 1000), (20, 1000)):
        b = budget_bytes // (shots * q)
        bits = torch.randint(0, 2, (b, shots, q), device="cuda", generator=g, dtype=torch.uint8)
        masks = torch.randint(1, 2 ** q, (b,), device="cuda", generator=g, dtype=torch.int32)
        ms = _time(torch, lambda: oe.shots_to_obs_moments_batch(bits, masks))
        kind = "SWAR" if q in (1, 2, 4, 8) else ("stream" if q <= 16 else "bytes")
        rows.append(_row(f"moments_{kind}_kernel n_qubits={q} shots={shots} (shots_to_obs_moments)", b,
                         shots * q + 4 + 8 + 16, ms, peak))
        del bits, masks
    n = 2
    b = 1 << 20
    rho = _rand_states(torch, b, 2 ** n, 71)
    k = 4 ** n - 1
    plan = tm.MlePlan(n, np.arange(1, k + 1, dtype=np.int32))
    ex = torch.rand((b, k), dtype=torch.float64, device="cuda", generator=g) * 1.2 - .6
    cnt = torch.full((b, k), 1000.0, dtype=torch.float64, device="cuda")
    ms = _time(torch, lambda: tm.state_log_likelihood_batch(plan, rho, ex, cnt))
    rows.append(_row("log_likelihood_packed_kernel<2> (state_log_likelihood; 30 log10 per experiment: FP64-bound)", b, 16 * 4 ** n + 16 * k + 8, ms, peak))
    del rho, ex, cnt
    for n, b in ((1, 1 << 20), (2, 1 << 16), (3, 1 << 10)):
        pplan = tm.PgdbPlan.complete(n)
        ex = torch.rand((b, pplan.S), dtype=torch.float64, device="cuda", generator=g) * 1.2 - .6
        out = torch.empty((b, 4 ** n, 4 ** n), dtype=torch.complex128, device="cuda")
        ms = _time(torch, lambda: tm.linear_inv_process_estimate_batch(pplan, ex, out=out))
        name = f"linproc_kernel<{n}>" if n == 1 else f"linproc_accum_kernel<{n}> + linproc_kernel<{n}> (two launches)"
        rows.append(_row(f"{name} (linear_inv_process_estimate, complete Pauli design)", b,
                         8 * pplan.S + 16 * 16 ** n, ms, peak))
        ms = _time(torch, lambda: pj.proj_choi_to_unitary_batch(out), reps=3, warmup=1)
        rows.append(_row(f"proj_unitary_kernel<{n}> (proj_choi_to_unitary; eigensolver: FP64-bound)", b,
                         32 * 16 ** n, ms, peak))
        del ex, out
    return rows


def cpu_convert_baseline(torch, per_n=None):
    """Oracle port (single process) on a subsample of BASELINE configs[3]: seconds per matrix of every conversion, and
    the GPU result's relative Frobenius error on the same matrices.  SURVEY 8(d): 64 matrices per n, 2 at n = 5."""
    import time
    from oracle import ref_numpy as orc
    from forest_benchmarking_b200.operator_tools import superoperator_transformations as st
    per_n = per_n or {1: 64, 2: 64, 3: 16, 4: 2, 5: 1}
    out = {}
    for n, cnt in per_n.items():
        d = 2 ** n
        rng = np.random.default_rng(4004 + n)
        kraus = np.stack([[np.sqrt(.7) * orc.haar_unitary(rng, d), np.sqrt(.3) * orc.haar_unitary(rng, d)]
                          for _ in range(cnt)])
        kd = torch.from_numpy(kraus).cuda()
        g_choi = st.kraus2choi_batch(kd)
        g_sup = st.reshuffle_batch(g_choi)
        g_pl = st.superop2pauli_liouville_batch(g_sup)
        g_back = st.pauli_liouville2superop_batch(g_pl)
        g_kraus, g_cnt, _ = st.choi2kraus_batch(g_choi)
        gpu = {k: v.cpu().numpy() for k, v in dict(choi=g_choi, sup=g_sup, pl=g_pl, back=g_back, kraus=g_kraus).items()}
        t = {k: 0.0 for k in ("kraus2choi", "choi2superop", "superop2pauli_liouville", "pauli_liouville2superop",
                              "choi2kraus")}
        err = 0.0

        def rel(a, b):
            return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
        for i in range(cnt):
            t0 = time.perf_counter(); choi = orc.kraus2choi(list(kraus[i])); t["kraus2choi"] += time.perf_counter() - t0
            t0 = time.perf_counter(); sup = orc.choi2superop(choi); t["choi2superop"] += time.perf_counter() - t0
            t0 = time.perf_counter(); pl = orc.superop2pauli_liouville(sup); t["superop2pauli_liouville"] += time.perf_counter() - t0
            t0 = time.perf_counter(); back = orc.pauli_liouville2superop(pl); t["pauli_liouville2superop"] += time.perf_counter() - t0
            t0 = time.perf_counter(); ks = orc.choi2kraus(choi); t["choi2kraus"] += time.perf_counter() - t0
            err = max(err, rel(gpu["choi"][i], choi), rel(gpu["sup"][i], sup), rel(gpu["pl"][i], pl),
                      rel(gpu["back"][i], back))
            # Kraus operators carry an eigenvector phase: compare through kraus2choi (the reference's own test)
            k_gpu = gpu["kraus"][i][: int(g_cnt[i].item())]
            err = max(err, rel(sum(orc.kraus2choi(k) for k in k_gpu), sum(orc.kraus2choi(k) for k in ks)))
        out[f"n={n}"] = {"matrices": cnt, "sec_per_matrix": {k: v / cnt for k, v in t.items()},
                         "max_rel_frobenius_err_gpu_vs_port": err}
    return out


def cpu_distance_baseline(torch, pairs=2000, n=4):
    """Oracle port on `pairs` pairs of BASELINE configs[4]: microseconds per pair and the GPU's absolute error."""
    import time
    from oracle import ref_numpy as orc
    from forest_benchmarking_b200 import distance_measures as dm
    d = 2 ** n
    rho, sig = _rand_states(torch, pairs, d, 61), _rand_states(torch, pairs, d, 62)
    fid = dm.fidelity_batch(rho, sig).cpu().numpy()
    td = dm.trace_distance_batch(rho, sig).cpu().numpy()
    r, s = rho.cpu().numpy(), sig.cpu().numpy()
    t0 = time.perf_counter()
    f_cpu = np.array([orc.fidelity(a, b) for a, b in zip(r, s)])
    t1 = time.perf_counter()
    t_cpu = np.array([orc.trace_distance(a, b) for a, b in zip(r, s)])
    t2 = time.perf_counter()
    return {"pairs": pairs, "fidelity_us_per_pair": (t1 - t0) / pairs * 1e6,
            "trace_distance_us_per_pair": (t2 - t1) / pairs * 1e6,
            "max_abs_err_fidelity": float(np.abs(fid - f_cpu).max()),
            "max_abs_err_trace_distance": float(np.abs(td - t_cpu).max()), "cores": 1, "kind": "port"}


def dumps(rows):
    return json.dumps(rows)
