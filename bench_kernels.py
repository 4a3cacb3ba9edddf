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
        rows.append(_row(f"proj_tp_kernel<{n}> TNI (proj_choi_to_trace_non_increasing)", b, 32 * m * m, ms, peak))
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
        kraus = _rand_c128(torch, (chunk, nk, d, d), 50 + n) * (1.0 / np.sqrt(nk * d))
        a, b_, ws = (torch.empty((chunk, m, m), dtype=torch.complex128, device="cuda") for _ in range(3))
        reps = max(1, batch // chunk)
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
        if n <= 3:
            ms = _time(torch, lambda: st.choi2kraus_batch(a), reps=3, warmup=1)
            rows.append(_row(f"choi2kraus n={n} (eigensolver: FP64-bound, GB/s for reference only)", chunk,
                             16 * m * m + 16 * m * m + 8 * m, ms, peak, {"chunks_for_batch": reps,
                                                                          "batch_ms": round(ms * reps, 3)}))
        if n == 4:
            sub = a[:148].contiguous()
            ms = _time(torch, lambda: st.choi2kraus_batch(sub), reps=1, warmup=1)
            rows.append(_row("choi2kraus n=4 (one-sided Jacobi out of L2, one 256x256 matrix per SM, 148 matrices)", 148,
                             16 * m * m + 16 * m * m + 8 * m, ms, peak, {"chunks_for_batch": 1, "batch_ms": round(ms, 3)}))
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
    for q, shots in ((1, 1000), (2, 1000), (2, 500), (4, 1000), (8, 1000), (3, 1000)):
        b = budget_bytes // (shots * q)
        bits = torch.randint(0, 2, (b, shots, q), device="cuda", generator=g, dtype=torch.uint8)
        masks = torch.randint(1, 2 ** q, (b,), device="cuda", generator=g, dtype=torch.int32)
        ms = _time(torch, lambda: oe.shots_to_obs_moments_batch(bits, masks))
        kind = "SWAR" if q in (1, 2, 4, 8) else "bytes"
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
    rows.append(_row("log_likelihood_kernel<2> (state_log_likelihood)", b, 16 * 4 ** n + 16 * k + 8, ms, peak))
    del rho, ex, cnt
    for n, b in ((1, 1 << 20), (2, 1 << 16), (3, 1 << 10)):
        pplan = tm.PgdbPlan.complete(n)
        ex = torch.rand((b, pplan.S), dtype=torch.float64, device="cuda", generator=g) * 1.2 - .6
        out = torch.empty((b, 4 ** n, 4 ** n), dtype=torch.complex128, device="cuda")
        ms = _time(torch, lambda: tm.linear_inv_process_estimate_batch(pplan, ex, out=out))
        rows.append(_row(f"linproc_kernel<{n}> (linear_inv_process_estimate, complete Pauli design)", b,
                         8 * pplan.S + 16 * 16 ** n, ms, peak))
        ms = _time(torch, lambda: pj.proj_choi_to_unitary_batch(out), reps=3, warmup=1)
        rows.append(_row(f"proj_unitary_kernel<{n}> (proj_choi_to_unitary; eigensolver: FP64-bound)", b,
                         32 * 16 ** n, ms, peak))
        del ex, out
    return rows


def dumps(rows):
    return json.dumps(rows)
