#!/usr/bin/env python
"""Benchmark harness (driver contract: python bench.py --gpus N --steps K --warmup W [--impl reference]).

Headline workload (BASELINE.json configs[1]): 2-qubit iterative_mle_state_estimate, batch = 4096
synthetic experiments per GPU, reference defaults (epsilon=.1, tol=1e-9, maxiter=10000).  One "step" is
one pass of the hot path over one batch.  Other workloads (--workload pgdb3q | mle_step | convert |
distances) report the other BASELINE configs with the same JSON schema.

  value : reconstructions/s with inputs already resident in HBM (CUDA-event timed, max over ranks)
  e2e   : same metric through the public batch API with pinned HOST buffers, H2D + D2H in the timed region
  roofline / cpu_baseline : see DESIGN.md "Measurement".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MLE tomography reconstructions/sec (batched n-qubit)"
UNIT = "reconstructions/s"
MLE_DEFAULTS = dict(epsilon=.1, tol=1e-9, maxiter=10_000)


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def flop_model_mle(n, k, iters):
    """SURVEY.md 8(d) canonical FLOP model of one reconstruction: iters * F_mle_iter(n, K)."""
    d = 2 ** n
    per_iter = 16 * d ** 3 + 8 * k * d + 2 * k * d + 10 * k + 6 * d * d
    return per_iter * np.asarray(iters, dtype=np.float64)


def dist_info():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores (the Python reference cannot travel)
# ------------------------------------------------------------------------------------------------
def _cpu_mle_item(args):
    pidx, ex, cnt, n, kw = args
    from oracle import ref_numpy as orc
    t0 = time.perf_counter()
    rho, it = orc.mle_state_estimate(pidx, np.ones(len(pidx)), ex, cnt, n, rebuild_paulis=True, **kw)
    return time.perf_counter() - t0, it, rho


def cpu_mle_sample(n, items, procs, seed=2002, data=None, return_states=False):
    """Times the faithful scalar port of iterative_mle_state_estimate (re-krons the Pauli matrices every
    iteration like the reference's lifted_pauli call, tomography.py:327) on `items` experiments spread
    over `procs` processes.  Returns (items/s, per-item seconds, iterations)."""
    from forest_benchmarking_b200 import synthetic as sy
    pidx, ex, cnt = data if data is not None else sy.state_tomography_batch(seed, items, n)[:3]
    jobs = [(pidx, ex[i], cnt[i], n, MLE_DEFAULTS) for i in range(items)]
    t0 = time.perf_counter()
    if procs > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_mle_item, jobs, chunksize=1)
    else:
        res = [_cpu_mle_item(j) for j in jobs]
    wall = time.perf_counter() - t0
    if return_states:
        return items / wall, [r[0] for r in res], [r[1] for r in res], np.stack([r[2] for r in res])
    return items / wall, [r[0] for r in res], [r[1] for r in res]


def _cpu_pgdb_item(args):
    settings, ex, cnt, n = args
    from oracle import ref_numpy as orc
    t0 = time.perf_counter()
    choi, c = orc.pgdb_process_estimate(settings, np.ones(len(settings)), ex, cnt, n, return_counters=True)
    return time.perf_counter() - t0, c, choi


def cpu_pgdb_sample(n, codes, pidx, ex, cnt, procs):
    """Times the oracle port of pgdb_process_estimate (dense design matrix, like the reference) on the given
    experiments, one per process.  Returns (items/s, per-item seconds, counters, choi matrices)."""
    settings = [(tuple(int(c) for c in s), int(k)) for s, k in zip(codes, pidx)]
    jobs = [(settings, ex[i], cnt[i], n) for i in range(len(ex))]
    t0 = time.perf_counter()
    if procs > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_pgdb_item, jobs, chunksize=1)
    else:
        res = [_cpu_pgdb_item(j) for j in jobs]
    wall = time.perf_counter() - t0
    return len(jobs) / wall, [r[0] for r in res], [r[1] for r in res], np.stack([r[2] for r in res])


def run_reference_arm(args):
    rank, world, _ = dist_info()
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    cores = os.cpu_count() or 1
    procs = min(cores, 64)
    items = procs  # one experiment per worker process and step
    vals = []
    t_start = time.perf_counter()
    for s in range(args.warmup + args.steps):
        v, _, _ = cpu_mle_sample(2, items, procs, seed=2002 + s)
        if s >= args.warmup:
            vals.append(v)
        if time.perf_counter() - t_start > 240:  # keep the whole arm within a few minutes
            if not vals:
                vals.append(v)
            break
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * items / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": {"workload": "2-qubit iterative_mle_state_estimate, reference defaults "
                               "(epsilon=.1, tol=1e-9, maxiter=10000), 1000 shots, K=15 Paulis",
                   "sample_items_per_step": items},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": f"{items} experiments per step, one per worker process, "
                                   "oracle/ref_numpy.mle_state_estimate(rebuild_paulis=True) -- the scalar "
                                   "restatement of the reference loop (pinned to the reference to 1e-11)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def measure_fp64_peak(torch, lib, _lib):
    import ctypes
    scratch = torch.zeros(8, dtype=torch.float64, device="cuda")
    blocks, threads, iters = 148 * 8, 256, 20000
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.qt_fp64_probe(blocks, threads, iters, _lib.ptr(scratch), _lib.current_stream_ptr()), "probe")
        e1.record()
        torch.cuda.synchronize()
        fl = blocks * threads * 8.0 * iters * 2.0
        best = max(best, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def run_ours(args):
    import torch
    from forest_benchmarking_b200 import _lib, synthetic as sy, tomography as tm
    rank, world, local = dist_info()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.lib()
    n, B = 2, args.batch
    hbm_peak, peak_src = measured_peaks()

    # synthetic batch for this rank (different experiments per rank: weak scaling, B per GPU)
    pidx, ex, cnt, _ = sy.state_tomography_batch(2002 + rank, B, n)
    K = len(pidx)
    plan = tm.MlePlan(n, pidx)
    ex_host = torch.from_numpy(ex).pin_memory()
    ex_dev = ex_host.cuda()
    rho = torch.empty((B, 4, 4), dtype=torch.complex128, device="cuda")
    iters = torch.empty((B,), dtype=torch.int32, device="cuda")
    rho_host = torch.empty((B, 4, 4), dtype=torch.complex128).pin_memory()
    gathered = torch.empty((world * B, 4, 4), dtype=torch.complex128, device="cuda") if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")  # 256 MB > 126 MB L2

    def step_resident():
        tm.iterative_mle_state_estimate_batch(plan, ex_dev, None, out=rho, iters_out=iters, **MLE_DEFAULTS)
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(torch.float64), rho.view(torch.float64))

    def step_e2e():
        d = ex_host.cuda(non_blocking=True)
        r, _ = tm.iterative_mle_state_estimate_batch(plan, d, None, out=rho, iters_out=iters, **MLE_DEFAULTS)
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(torch.float64), r.view(torch.float64))
        rho_host.copy_(r, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn, steps, warmup, kernel_only=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = []
        for _ in range(steps):
            flush.zero_()  # L2 flush between timed iterations (outside the event pair)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()) / steps, ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_res, _ = timed(step_resident, args.steps, args.warmup)
    ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    clocks = sampler.stop() if rank == 0 else None

    # kernel-only timing of the dominant kernel (no collective), for the roofline
    def kernel_only():
        tm.iterative_mle_state_estimate_batch(plan, ex_dev, None, out=rho, iters_out=iters, **MLE_DEFAULTS)
    ms_kernel, _ = timed(kernel_only, args.steps, 1)
    it_host = iters.cpu().numpy()
    flops = float(flop_model_mle(n, K, it_host).sum())
    fp64_peak = measure_fp64_peak(torch, lib, _lib)
    bytes_item = 8 * K + 16 * 16 + 4  # expectations in, rho out, iteration counter out (counts unused: beta = 0)

    # HBM-roofline view: ONE R-rho-R update per experiment streamed through HBM
    Bs = 1 << 22
    exs = torch.rand((15, Bs), dtype=torch.float64, device="cuda") * 1.2 - .6
    rs = torch.eye(4, dtype=torch.complex128, device="cuda").repeat(Bs, 1, 1) / 4
    ro = torch.empty_like(rs)
    ms_stream, _ = timed(lambda: tm.mle_step_batch(2, exs, rs, .1, out=ro), 5, 3)
    stream_bytes = Bs * (15 * 8 + 256 + 256)
    del exs, rs, ro

    if rank == 0:
        value = world * B / (ms_res * 1e-3)
        e2e = world * B / (ms_e2e * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
            "config": {"workload": "2-qubit iterative_mle_state_estimate (BASELINE configs[1]), reference defaults "
                                   "epsilon=.1 tol=1e-9 maxiter=10000, 1000 shots, K=15 Paulis",
                       "batch_per_gpu": B, "global_batch": world * B,
                       "l2": "flushed (256 MB memset) between timed iterations",
                       "collective": "all_gather of reconstructed states" if world > 1 else "none",
                       "iterations_mean": float(it_host.mean()), "hit_maxiter_frac": float((it_host >= 10000).mean())},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(ex_host.numel() * 8),
                    "d2h_bytes_per_step": int(rho_host.numel() * 16), "ms_per_step": ms_e2e},
            "gpu_launches": args.steps * 1,
            "clocks": clocks,
            "roofline": {
                "kernel": "mle_quad_kernel (fused persistent R-rho-R loop, 4 lanes per experiment, state in registers)",
                "bound": "fp64", "achieved": flops / (ms_kernel * 1e-3) / 1e12, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": flops / (ms_kernel * 1e-3) / 1e12 / fp64_peak,
                "peak_source": "measured live: qt_fp64_probe DFMA chains (MEASURED_PEAKS.json has no FP64 entry)",
                "flop_model": "SURVEY.md 8(d): iters*(16d^3+10Kd+10K+6d^2) = 1870/iteration at n=2,K=15, actual iters",
                "traffic": 524032,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full "
                                "(profiles/r01_ncu_mle_quad_kernel_final.md); the 1.5 MB of inputs/outputs stay in L2",
                "hbm_view": {"bound": "hbm", "achieved": B * bytes_item / (ms_kernel * 1e-3) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": B * bytes_item / (ms_kernel * 1e-3) / 1e9 / hbm_peak,
                             "note": "compulsory bytes only (380 B/item); the loop state never leaves registers, "
                                     "so this kernel is FP64/latency-bound, not HBM-bound"},
            },
            "roofline_streaming": {
                "kernel": "mle_step_kernel<2> (ONE R-rho-R update, rho HBM->HBM)", "bound": "hbm",
                "achieved": stream_bytes / (ms_stream * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": stream_bytes / (ms_stream * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src,
                "bytes_per_item": 632, "items": Bs, "traffic": None},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            procs = min(cores, 8)
            # the first `procs` experiments of the batch the GPU just reconstructed: baseline timing AND parity
            v, per_item, its, rho_cpu = cpu_mle_sample(2, procs, procs, data=(pidx, ex[:procs], cnt[:procs]),
                                                       return_states=True)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": procs, "kind": "port",
                                    "sec_per_item_one_core": float(np.mean(per_item)),
                                    "sample": f"the first {procs} experiments of the GPU batch, one per process "
                                              f"(host has {cores} cores), oracle scalar port with per-iteration "
                                              f"Pauli re-kron like the reference; iterations {its}"}
            rho_gpu, it_gpu = rho[:procs].cpu().numpy(), iters[:procs].cpu().numpy()
            errs = [float(np.linalg.norm(rho_gpu[i] - rho_cpu[i]) / np.linalg.norm(rho_cpu[i])) for i in range(procs)]
            line["parity"] = {"max_rel_frobenius_err": max(errs), "tolerance": 1e-6, "items": procs,
                              "iteration_count_mismatches": int(sum(int(a) != int(b) for a, b in zip(it_gpu, its))),
                              "against": "oracle port (pinned to the reference, tests/test_oracle_vs_reference.py) "
                                         "on the same experiments"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def flop_model_pgdb(n, n_in, counters):
    """eigh_calls*44 m^3 + (cost_evals + 2 outer)*F_A, F_A = 8 d^4 n_in + n_in*4*n*4^n."""
    d, m = 2 ** n, 4 ** n
    fa = 8.0 * d ** 4 * n_in + n_in * 4.0 * n * m
    c = np.asarray(counters, dtype=np.float64)
    return c[:, 2] * 44.0 * m ** 3 + (c[:, 1] + 2 * c[:, 0]) * fa


def run_pgdb(args):
    """BASELINE configs[2]: 3-qubit pgdb_process_estimate (64x64 Choi), batch 1024 per GPU, Pauli inputs."""
    import torch
    from forest_benchmarking_b200 import _lib, synthetic as sy, tomography as tm
    rank, world, local = dist_info()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = {"pgdb3q": 3, "pgdb2q": 2, "pgdb1q": 1}[args.workload]
    if args.eigh_tol is not None:
        import ctypes
        _lib.check(_lib.lib().qt_set_eigh_tolerance(ctypes.c_double(args.eigh_tol)), "qt_set_eigh_tolerance")
    B = args.batch if args.batch != 4096 else 1024
    codes, pidx, ex, cnt, _ = sy.process_tomography_batch(3003 + rank, B, n, in_basis=args.in_basis)
    plan = tm.PgdbPlan(n, codes, pidx)
    m = 4 ** n
    ex_host, cnt_host = torch.from_numpy(ex).pin_memory(), torch.from_numpy(cnt).pin_memory()
    ex_dev, cnt_dev = ex_host.cuda(), cnt_host.cuda()
    out = torch.empty((B, m, m), dtype=torch.complex128, device="cuda")
    out_host = torch.empty((B, m, m), dtype=torch.complex128).pin_memory()
    gathered = torch.empty((world * B, m, m), dtype=torch.complex128, device="cuda") if world > 1 else None
    nbytes = int(_lib.lib().qt_pgdb_workspace_bytes(plan._h, B))
    ws = torch.empty((nbytes // 8,), dtype=torch.float64, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")
    last = {}

    def step_resident():
        _, c = tm.pgdb_process_estimate_batch(plan, ex_dev, cnt_dev, True, out=out, return_counters=True, workspace=ws)
        last["c"] = c
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(torch.float64), out.view(torch.float64))

    def step_e2e():
        e, c = ex_host.cuda(non_blocking=True), cnt_host.cuda(non_blocking=True)
        tm.pgdb_process_estimate_batch(plan, e, c, True, out=out, workspace=ws)
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(torch.float64), out.view(torch.float64))
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.barrier()
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()) / steps

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_res = timed(step_resident, args.steps, args.warmup)
    ms_e2e = timed(step_e2e, args.steps, 1)
    clocks = sampler.stop() if rank == 0 else None
    counters = last["c"].cpu().numpy()
    fp64_peak = measure_fp64_peak(torch, _lib.lib(), _lib)
    flops = float(flop_model_pgdb(n, plan.n_in, counters).sum())
    hbm_peak, _ = measured_peaks()
    if rank == 0:
        bytes_item = 2 * 8 * plan.S + 16 * m * m
        line = {
            "metric": METRIC, "value": world * B / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
            "config": {"workload": f"{n}-qubit pgdb_process_estimate (BASELINE configs[2]), {args.in_basis} inputs, "
                                   f"{plan.S} settings x 1000 shots, Haar-random unitary truth",
                       "batch_per_gpu": B, "global_batch": world * B, "l2": "flushed between timed iterations",
                       "collective": "all_gather of reconstructed Choi matrices" if world > 1 else "none",
                       "outer_mean": float(counters[:, 0].mean()), "cost_evals_mean": float(counters[:, 1].mean()),
                       "eigh_calls_mean": float(counters[:, 2].mean()),
                       "jacobi_sweeps_per_eigh": float(counters[:, 3].sum() / max(1, counters[:, 2].sum()))},
            "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(2 * ex_host.numel() * 8), "d2h_bytes_per_step": int(out_host.numel() * 16)},
            "gpu_launches": args.steps, "clocks": clocks,
            "roofline": {"kernel": f"pgdb_kernel<{n}> (fused PGD + Dykstra + Jacobi eigh, one experiment per "
                                   f"{'block' if n == 3 else 'warp'})",
                         "bound": "fp64", "achieved": flops / (ms_res * 1e-3) / 1e12, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": flops / (ms_res * 1e-3) / 1e12 / fp64_peak,
                         "flop_model": "SURVEY.md 8(d): eigh*44m^3 + (cost_evals+2*outer)*F_A with the items' actual counters",
                         "peak_source": "measured live: qt_fp64_probe", "traffic": None,
                         "hbm_view": {"bound": "hbm", "achieved": B * bytes_item / (ms_res * 1e-3) / 1e9,
                                      "peak": hbm_peak, "unit": "GB/s",
                                      "frac": B * bytes_item / (ms_res * 1e-3) / 1e9 / hbm_peak,
                                      "note": "compulsory bytes only"}},
        }
        want_cpu = (n <= 2 and not args.no_cpu_baseline) or args.cpu_baseline
        if world == 1 and want_cpu:
            os.environ.setdefault("OMP_NUM_THREADS", "1")
            cores = os.cpu_count() or 1
            items = min(cores, 8) if n <= 2 else 2
            v, per_item, cs, choi_cpu = cpu_pgdb_sample(n, codes, pidx, ex[:items], cnt[:items], items)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": items, "kind": "port",
                                    "sec_per_item_one_core": float(np.mean(per_item)),
                                    "sample": f"the first {items} experiments of the GPU batch, one per process (host has "
                                              f"{cores} cores), oracle port with the reference's dense design matrix"}
            choi_gpu = out[:items].cpu().numpy()
            errs = [float(np.linalg.norm(choi_gpu[i] - choi_cpu[i]) / np.linalg.norm(choi_cpu[i])) for i in range(items)]
            line["parity"] = {"max_rel_frobenius_err": max(errs), "tolerance": 1e-6, "items": items,
                              "outer_iteration_mismatches": int(sum(int(counters[i, 0]) != int(cs[i]["outer"])
                                                                    for i in range(items))),
                              "against": "oracle port (pinned to the reference) on the same experiments"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_kernels(args):
    """--workload streaming | convert | distances | next: per-kernel HBM-roofline tables (bench_kernels.py)."""
    import torch
    import bench_kernels as bk
    rank, world, local = dist_info()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    peak, peak_src = measured_peaks()
    sampler = ClockSampler(local)
    sampler.start()
    if args.workload == "streaming":
        rows = bk.streaming_rows(torch, peak)
        headline = next(r for r in rows if r["kernel"].startswith("mle_step_kernel<2>"))
        workload = "HBM-bound kernels of the path, one launch each over a 2 GiB working set (inputs > L2)"
    elif args.workload == "next":
        rows = bk.next_rows(torch, peak)
        headline = next(r for r in rows if "n_qubits=2 shots=1000" in r["kernel"])
        workload = "SURVEY 8(f) next rows: shots -> moments, log-likelihood, linear-inversion process estimate, closest unitary"
    elif args.workload == "convert":
        rows = bk.convert_rows(torch, peak)
        headline = next(r for r in rows if r["kernel"].startswith("superop2pauli_liouville n=3"))
        workload = "BASELINE configs[3]: conversion sweep kraus<->choi<->pauli_liouville, n=1..5, batch 16384 (chunked)"
    else:
        pairs = 1_000_000 // world
        rows = bk.distance_rows(torch, peak, pairs)
        headline = rows[0]
        workload = f"BASELINE configs[4]: fidelity + trace_distance, 10^6 4-qubit pairs / {world} GPU(s)"
    clocks = sampler.stop()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if args.workload == "convert":
            cpu = bk.cpu_convert_baseline(torch)
        elif args.workload == "distances":
            cpu = bk.cpu_distance_baseline(torch)
    if rank == 0:
        print(json.dumps({
            "cpu_baseline": cpu,
            "metric": "kernel roofline sweep", "value": headline["items_per_s"], "unit": "items/s (headline kernel)",
            "n_gpus": world, "steps": 5, "warmup": 3, "ms_per_step": headline["ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
            "config": {"workload": workload, "l2": "inputs larger than the 126 MB L2"},
            "clocks": clocks,
            "roofline": {"kernel": headline["kernel"], "bound": "hbm", "achieved": headline["achieved_gbs"],
                         "peak": peak, "unit": "GB/s", "frac": headline["frac_of_hbm_peak"], "peak_source": peak_src,
                         "traffic": None},
            "kernels": rows}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mle2q", choices=["mle2q", "pgdb3q", "pgdb2q", "pgdb1q", "streaming",
                                                            "convert", "distances", "next"])
    ap.add_argument("--in-basis", default="pauli", choices=["pauli", "sic"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline", action="store_true", help="pgdb3q: also time 2 experiments on the host (minutes)")
    ap.add_argument("--eigh-tol", type=float, default=None, help="experiment: qt_set_eigh_tolerance value")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload.startswith("pgdb"):
        run_pgdb(args)
    elif args.workload in ("streaming", "convert", "distances", "next"):
        run_kernels(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
