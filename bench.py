#!/usr/bin/env python
"""Benchmark harness (driver contract: python bench.py --gpus N --steps K --warmup W [--impl reference]).

ONE JSON line.  Its top-level keys are the headline workload, BASELINE.json configs[1]: 2-qubit
iterative_mle_state_estimate, batch = 4096 synthetic experiments per GPU (weak scaling), reference defaults
(epsilon=.1, tol=1e-9, maxiter=10000).  One "step" is one pass of the hot path over one batch.

  value : reconstructions/s with inputs already resident in HBM (CUDA-event timed, max over ranks)
  e2e   : same metric through the public batch API with pinned HOST buffers, H2D + D2H in the timed region
  roofline / cpu_baseline / parity : see DESIGN.md "Measurement".

The same line carries, as nested objects with the same schema, the other BASELINE configs (so that the driver's
BENCH / SCALE records hold them):

  "pgdb3q"    configs[2]: 3-qubit pgdb_process_estimate, GLOBAL batch 1024 sharded B/N per rank (strong scaling),
              all-gather of the Choi matrices inside the timed region, per-rank kernel time min/max, FP64 roofline,
              parity of 2 experiments against the CPU port (N = 1 only)
  "distances" configs[4]: fidelity + trace_distance over 10^6 random 4-qubit pairs split over the ranks + gather
  "mle_batch_sweep", "mle3q": the headline kernel at batch 4096 .. 262144, and 3-qubit MLE throughput

`--workload X` restricts the run to one part (profiling runs); `--workload streaming|convert|next` print the
per-kernel HBM-roofline tables of bench_kernels.py.
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
DEFAULT_PARTS = ("mle2q", "pgdb3q", "distances", "mle_batch_sweep", "mle3q")


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


def tracked_json(name):
    """profiles/<name> (tracked evidence files written by scripts/, e.g. the FP64 peaks and the ncu DRAM traffic)."""
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def kernel_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the tracked ncu summary, or None."""
    t = tracked_json("r02_traffic.json") or {}
    e = t.get(kernel)
    return (e["dram_bytes_per_launch"], e["source"]) if e else (None, "no ncu --set full capture of this kernel is tracked")


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
        return self

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


def flop_model_pgdb(n, n_in, counters):
    """eigh_calls*44 m^3 + (cost_evals + 2 outer)*F_A, F_A = 8 d^4 n_in + n_in*4*n*4^n."""
    d, m = 2 ** n, 4 ** n
    fa = 8.0 * d ** 4 * n_in + n_in * 4.0 * n * m
    c = np.asarray(counters, dtype=np.float64)
    return c[:, 2] * 44.0 * m ** 3 + (c[:, 1] + 2 * c[:, 0]) * fa


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
class Ctx:
    """One process = one GPU.  Owns the process group, the L2-flush buffer and the timing helper."""

    def __init__(self, args):
        import torch
        from forest_benchmarking_b200 import _lib
        self.torch, self._lib, self.args = torch, _lib, args
        self.rank, self.world, self.local = dist_info()
        assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.lib = _lib.lib()
        self.hbm_peak, self.hbm_src = measured_peaks()
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")  # 256 MB > 126 MB L2
        self._fp64 = None

    def timed(self, fn, steps, warmup):
        """W untimed warm-up steps, then `steps` steps, each between CUDA events on the launching stream with the L2
        flushed (256 MB memset) before it; barrier + synchronize on both sides; returns (mean ms per step as the MAX over
        ranks, this rank's mean ms)."""
        torch, dist = self.torch, self.dist
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        ms = []
        for _ in range(steps):
            self.flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        mine = sum(ms) / steps
        tot = torch.tensor([mine], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()), mine

    def gather_scalar(self, v):
        torch, dist = self.torch, self.dist
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        if not dist:
            return [float(v)]
        out = torch.empty((self.world,), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(out, t)
        return [float(x) for x in out.cpu()]

    def fp64_peak(self):
        """FP64 roofline denominator: max(live DFMA probe, tracked DFMA / DMMA micro-benchmark peaks)."""
        if self._fp64 is None:
            torch, lib, _lib = self.torch, self.lib, self._lib
            scratch = torch.zeros(8, dtype=torch.float64, device="cuda")
            blocks, threads, iters = 148 * 8, 256, 20000
            live = 0.0
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(lib.qt_fp64_probe(blocks, threads, iters, _lib.ptr(scratch), _lib.current_stream_ptr()), "probe")
                e1.record()
                torch.cuda.synchronize()
                live = max(live, blocks * threads * 8.0 * iters * 2.0 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
            tr = tracked_json("r02_fp64_peaks.json") or {}
            cands = {"live qt_fp64_probe (DFMA chains)": live,
                     "profiles/r02_fp64_peaks.json DFMA": float(tr.get("fp64_dfma_peak_tflops", 0.0)),
                     "profiles/r02_fp64_peaks.json DMMA (mma.sync m8n8k4.f64)": float(tr.get("fp64_dmma_peak_tflops", 0.0))}
            src = max(cands, key=cands.get)
            self._fp64 = (cands[src], src, {k: round(v, 2) for k, v in cands.items()})
        return self._fp64


def part_mle2q(ctx):
    """BASELINE configs[1] -- the headline."""
    torch, dist, args = ctx.torch, ctx.dist, ctx.args
    from forest_benchmarking_b200 import synthetic as sy, tomography as tm
    rank, world = ctx.rank, ctx.world
    n, B = 2, args.batch
    # synthetic batch for this rank (different experiments per rank: weak scaling, B per GPU)
    pidx, ex, cnt, _ = sy.state_tomography_batch(2002 + rank, B, n)
    K = len(pidx)
    plan = tm.MlePlan(n, pidx)
    ex_host = torch.from_numpy(ex).pin_memory()
    ex_dev = ex_host.cuda()
    rho = torch.empty((B, 4, 4), dtype=torch.complex128, device="cuda")
    iters = torch.empty((B,), dtype=torch.int32, device="cuda")
    rho_host = torch.empty((B, 4, 4), dtype=torch.complex128).pin_memory()
    iters_host = torch.empty((B,), dtype=torch.int32).pin_memory()
    gathered = torch.empty((world * B, 4, 4), dtype=torch.complex128, device="cuda") if world > 1 else None

    def kernel_only():
        tm.iterative_mle_state_estimate_batch(plan, ex_dev, None, out=rho, iters_out=iters, **MLE_DEFAULTS)

    def step_resident():
        kernel_only()
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(torch.float64), rho.view(torch.float64))

    def step_e2e():
        d = ex_host.cuda(non_blocking=True)
        r, it = tm.iterative_mle_state_estimate_batch(plan, d, None, out=rho, iters_out=iters, **MLE_DEFAULTS)
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(torch.float64), r.view(torch.float64))
        rho_host.copy_(r, non_blocking=True)
        iters_host.copy_(it, non_blocking=True)  # the drop-in needs them to warn at maxiter (tomography.py:244)
        torch.cuda.current_stream().synchronize()

    sampler = ClockSampler(ctx.local).start() if rank == 0 else None
    ms_res, _ = ctx.timed(step_resident, args.steps, args.warmup)
    ms_e2e, _ = ctx.timed(step_e2e, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    ms_kernel, _ = ctx.timed(kernel_only, args.steps, 1)  # dominant kernel alone (no collective), for the roofline
    it_host = iters.cpu().numpy()
    flops = float(flop_model_mle(n, K, it_host).sum())
    fp64_peak, fp64_src, fp64_cands = ctx.fp64_peak()
    bytes_item = 8 * K + 16 * 16 + 4  # expectations in, rho out, iteration counter out (counts unused: beta = 0)

    # HBM-roofline view: ONE R-rho-R update per experiment streamed through HBM
    Bs = 1 << 22
    exs = torch.rand((15, Bs), dtype=torch.float64, device="cuda") * 1.2 - .6
    rs = torch.eye(4, dtype=torch.complex128, device="cuda").repeat(Bs, 1, 1) / 4
    ro = torch.empty_like(rs)
    ms_stream, _ = ctx.timed(lambda: tm.mle_step_batch(2, exs, rs, .1, out=ro), 5, 3)
    stream_bytes = Bs * (15 * 8 + 256 + 256)
    del exs, rs, ro
    if rank != 0:
        return None
    kname = "mle_lane16_kernel" if getattr(tm, "KERNEL_LANE16", None) is not None and B <= 32768 else "mle_quad_kernel"
    traffic, traffic_src = kernel_traffic(kname)
    line = {
        "metric": METRIC, "value": world * B / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": {"workload": "2-qubit iterative_mle_state_estimate (BASELINE configs[1]), reference defaults "
                               "epsilon=.1 tol=1e-9 maxiter=10000, 1000 shots, K=15 Paulis",
                   "batch_per_gpu": B, "global_batch": world * B,
                   "l2": "flushed (256 MB memset) between timed iterations",
                   "collective": "all_gather of reconstructed states" if world > 1 else "none",
                   "iterations_mean": float(it_host.mean()), "hit_maxiter_frac": float((it_host >= 10000).mean())},
        "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(ex_host.numel() * 8),
                "d2h_bytes_per_step": int(rho_host.numel() * 16 + iters_host.numel() * 4), "ms_per_step": ms_e2e},
        "gpu_launches": args.steps * 1,
        "clocks": clocks,
        "roofline": {
            "kernel": f"{kname} (fused persistent R-rho-R loop, state in registers)",
            "bound": "fp64", "achieved": flops / (ms_kernel * 1e-3) / 1e12, "peak": fp64_peak,
            "unit": "TFLOP/s", "frac": flops / (ms_kernel * 1e-3) / 1e12 / fp64_peak,
            "peak_source": fp64_src, "peak_candidates_tflops": fp64_cands,
            "flop_model": "SURVEY.md 8(d): iters*(16d^3+10Kd+10K+6d^2) = 1870/iteration at n=2,K=15, actual iters",
            "kernel_ms": ms_kernel, "traffic": traffic, "traffic_source": traffic_src,
            "hbm_view": {"bound": "hbm", "achieved": B * bytes_item / (ms_kernel * 1e-3) / 1e9, "peak": ctx.hbm_peak,
                         "unit": "GB/s", "frac": B * bytes_item / (ms_kernel * 1e-3) / 1e9 / ctx.hbm_peak,
                         "note": "compulsory bytes only (380 B/item); the loop state never leaves registers, "
                                 "so this kernel is FP64/latency-bound, not HBM-bound"},
        },
        "roofline_streaming": {
            "kernel": "mle_step_herm_kernel (ONE R-rho-R update, rho HBM->HBM)", "bound": "hbm",
            "achieved": stream_bytes / (ms_stream * 1e-3) / 1e9, "peak": ctx.hbm_peak, "unit": "GB/s",
            "frac": stream_bytes / (ms_stream * 1e-3) / 1e9 / ctx.hbm_peak, "peak_source": ctx.hbm_src,
            "bytes_per_item": 632, "items": Bs, "traffic": kernel_traffic("mle_step_herm_kernel")[0]},
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
    return line


def part_mle_sweep(ctx):
    """The headline kernel family at larger batches (the BASELINE batch leaves the GPU under-occupied)."""
    torch = ctx.torch
    from forest_benchmarking_b200 import synthetic as sy, tomography as tm
    rows = []
    pidx, ex, _, _ = sy.state_tomography_batch(2002 + ctx.rank, 16384, 2)
    plan = tm.MlePlan(2, pidx)
    for B in (4096, 16384, 65536, 262144):
        e = torch.from_numpy(np.tile(ex, (max(1, B // 16384), 1))[:B]).cuda()
        rho = torch.empty((B, 4, 4), dtype=torch.complex128, device="cuda")
        it = torch.empty((B,), dtype=torch.int32, device="cuda")
        ms, _ = ctx.timed(lambda: tm.iterative_mle_state_estimate_batch(plan, e, None, out=rho, iters_out=it,
                                                                        **MLE_DEFAULTS), 2, 1)
        rows.append({"batch_per_gpu": B, "ms_per_step": ms, "value": ctx.world * B / (ms * 1e-3), "unit": UNIT})
        del e, rho, it
    return {"workload": "2-qubit iterative_mle_state_estimate, reference defaults, kernel chosen by AUTO per batch",
            "rows": rows} if ctx.rank == 0 else None


def part_mle3q(ctx):
    """3-qubit MLE (the case the reference needs ~50 s per reconstruction for)."""
    torch = ctx.torch
    from forest_benchmarking_b200 import synthetic as sy, tomography as tm
    n, B = 3, 2048
    pidx, ex, _, _ = sy.state_tomography_batch(2103 + ctx.rank, B, n)
    plan = tm.MlePlan(n, pidx)
    e = torch.from_numpy(ex).cuda()
    rho = torch.empty((B, 8, 8), dtype=torch.complex128, device="cuda")
    it = torch.empty((B,), dtype=torch.int32, device="cuda")
    ms, _ = ctx.timed(lambda: tm.iterative_mle_state_estimate_batch(plan, e, None, out=rho, iters_out=it,
                                                                    **MLE_DEFAULTS), 2, 1)
    if ctx.rank != 0:
        return None
    ith = it.cpu().numpy()
    flops = float(flop_model_mle(n, len(pidx), ith).sum())
    peak, src, _ = ctx.fp64_peak()
    return {"workload": "3-qubit iterative_mle_state_estimate, reference defaults, 1000 shots, K=63 Paulis",
            "batch_per_gpu": B, "ms_per_step": ms, "value": ctx.world * B / (ms * 1e-3), "unit": UNIT,
            "iterations_mean": float(ith.mean()), "hit_maxiter_frac": float((ith >= 10000).mean()),
            "roofline": {"kernel": "mle_warp_kernel<3> (one experiment per warp, state in shared memory)", "bound": "fp64",
                         "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": flops / (ms * 1e-3) / 1e12 / peak, "peak_source": src,
                         "flop_model": "SURVEY.md 8(d) F_mle_iter(3, 63) = 13 862 per iteration x actual iterations"},
            "reference_note": "SURVEY section 6 probe: 47-59 s per reconstruction on the reference CPU path"}


def part_pgdb(ctx, n=3, global_batch=1024, in_basis="pauli"):
    """BASELINE configs[2]: n-qubit pgdb_process_estimate, GLOBAL batch sharded B/N per rank (strong scaling), one
    all-gather of the reconstructed Choi matrices inside the timed region."""
    torch, dist, args = ctx.torch, ctx.dist, ctx.args
    from forest_benchmarking_b200 import synthetic as sy, tomography as tm
    from forest_benchmarking_b200.sharding import shard_range, all_gather_states
    rank, world = ctx.rank, ctx.world
    codes, pidx, ex, cnt, _ = sy.process_tomography_batch(3003, global_batch, n, in_basis=in_basis)  # same on every rank
    lo, hi = shard_range(global_batch, world, rank)
    B = hi - lo
    plan = tm.PgdbPlan(n, codes, pidx)
    m = 4 ** n
    ex_host = torch.from_numpy(np.ascontiguousarray(ex[lo:hi])).pin_memory()
    cnt_host = torch.from_numpy(np.ascontiguousarray(cnt[lo:hi])).pin_memory()
    ex_dev, cnt_dev = ex_host.cuda(), cnt_host.cuda()
    out = torch.empty((B, m, m), dtype=torch.complex128, device="cuda")
    full_host = torch.empty((global_batch, m, m), dtype=torch.complex128).pin_memory()
    nbytes = int(ctx.lib.qt_pgdb_workspace_bytes(plan._h, B))
    ws = torch.empty((max(nbytes, 8) // 8,), dtype=torch.float64, device="cuda")
    last = {}
    tol = args.eigh_tol  # None = library default (1e-8); printed below

    def kernel_only():
        _, c = tm.pgdb_process_estimate_batch(plan, ex_dev, cnt_dev, True, out=out, return_counters=True, workspace=ws,
                                              eigh_rel_tol=tol)
        last["c"] = c

    def step_resident():
        kernel_only()
        last["full"] = all_gather_states(out, global_batch)

    def step_e2e():
        e, c = ex_host.cuda(non_blocking=True), cnt_host.cuda(non_blocking=True)
        tm.pgdb_process_estimate_batch(plan, e, c, True, out=out, workspace=ws, eigh_rel_tol=tol)
        full = all_gather_states(out, global_batch)
        full_host.copy_(full, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    steps = max(1, min(args.steps, 3)) if n == 3 else args.steps  # ~1-2 s per step at n = 3
    sampler = ClockSampler(ctx.local).start() if rank == 0 else None
    ms_res, _ = ctx.timed(step_resident, steps, args.warmup)
    ms_e2e, _ = ctx.timed(step_e2e, steps, 1)
    clocks = sampler.stop() if sampler else None
    ms_kernel, mine = ctx.timed(kernel_only, 1, 0)
    per_rank = ctx.gather_scalar(mine)
    counters = last["c"].to(torch.float64)
    if world > 1:  # FLOP model over ALL items: gather the per-item counters
        counters = all_gather_states(counters.contiguous(), global_batch)
    counters = counters.cpu().numpy()
    fp64_peak, fp64_src, _ = ctx.fp64_peak()
    flops = float(flop_model_pgdb(n, plan.n_in, counters).sum())
    if rank != 0:
        return None
    kname = f"pgdb_kernel<{n}>"
    traffic, traffic_src = kernel_traffic(kname)
    bytes_item = 2 * 8 * plan.S + 16 * m * m
    obj = {
        "metric": METRIC, "value": global_batch / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True,
        "scaling": "strong", "dtype": "f64 (complex128)", "data": "synthetic",
        "config": {"workload": f"{n}-qubit pgdb_process_estimate (BASELINE configs[2]), {in_basis} inputs, "
                               f"{plan.S} settings x 1000 shots, Haar-random unitary truth",
                   "global_batch": global_batch, "batch_per_gpu": B, "sharding": "contiguous B/N slices, no data-path "
                   "collective; ONE all_gather of the Choi matrices inside the timed region",
                   "l2": "flushed between timed iterations",
                   "eigh_rel_tol": tol if tol is not None else ("library default: 1e-5 + first-order correction of the PSD "
                                                                "projection" if n >= 3 else "library default: 1e-8"),
                   "outer_mean": float(counters[:, 0].mean()), "cost_evals_mean": float(counters[:, 1].mean()),
                   "eigh_calls_mean": float(counters[:, 2].mean()),
                   "jacobi_sweeps_per_eigh": float(counters[:, 3].sum() / max(1, counters[:, 2].sum()))},
        "e2e": {"value": global_batch / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(2 * ex_host.numel() * 8), "d2h_bytes_per_step": int(full_host.numel() * 16)},
        "gpu_launches": steps, "clocks": clocks,
        "kernel_ms_per_rank": {"min": min(per_rank), "max": max(per_rank), "all": [round(x, 2) for x in per_rank]},
        "roofline": {"kernel": f"{kname} (fused PGD + Dykstra + Jacobi eigh, one experiment per "
                               f"{'block' if n == 3 else 'warp'})",
                     "bound": "fp64", "achieved": flops / (ms_kernel * 1e-3) / 1e12, "peak": fp64_peak * world,
                     "peak_note": f"{world} GPU(s) x {fp64_peak:.2f} TFLOP/s",
                     "unit": "TFLOP/s", "frac": flops / (ms_kernel * 1e-3) / 1e12 / (fp64_peak * world),
                     "kernel_ms": ms_kernel,
                     "flop_model": "SURVEY.md 8(d): eigh*44m^3 + (cost_evals+2*outer)*F_A with the items' actual counters "
                                   "(all ranks' items / slowest rank's kernel time)",
                     "peak_source": fp64_src, "traffic": traffic, "traffic_source": traffic_src,
                     "hbm_view": {"bound": "hbm", "achieved": global_batch * bytes_item / (ms_kernel * 1e-3) / 1e9,
                                  "peak": ctx.hbm_peak * world, "unit": "GB/s",
                                  "frac": global_batch * bytes_item / (ms_kernel * 1e-3) / 1e9 / (ctx.hbm_peak * world),
                                  "note": "compulsory bytes only"}},
    }
    want_cpu = world == 1 and not args.no_cpu_baseline and (n <= 2 or not args.no_pgdb_cpu)
    if want_cpu:
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        cores = os.cpu_count() or 1
        items = min(cores, 8) if n <= 2 else 2
        v, per_item, cs, choi_cpu = cpu_pgdb_sample(n, codes, pidx, ex[:items], cnt[:items], items)
        obj["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": items, "kind": "port",
                               "sec_per_item_one_core": float(np.mean(per_item)),
                               "sample": f"the first {items} experiments of the GPU batch, one per process (host has "
                                         f"{cores} cores), oracle port with the reference's dense design matrix"}
        choi_gpu = out[:items].cpu().numpy()
        errs = [float(np.linalg.norm(choi_gpu[i] - choi_cpu[i]) / np.linalg.norm(choi_cpu[i])) for i in range(items)]
        obj["parity"] = {"max_rel_frobenius_err": max(errs), "tolerance": 1e-6, "items": items,
                         "outer_iteration_mismatches": int(sum(int(counters[i, 0]) != int(cs[i]["outer"])
                                                               for i in range(items))),
                         "eigh_call_mismatches": int(sum(int(counters[i, 2]) != int(cs[i]["eighs"])
                                                         for i in range(items))),
                         "against": "oracle port (pinned to the reference) on the same experiments"}
    return obj


def part_distances(ctx, pairs_total=1_000_000, n=4):
    """BASELINE configs[4]: fidelity + trace_distance over 10^6 random n-qubit pairs split over the ranks, then one
    all-gather of the two result vectors (inside the timed region)."""
    torch, dist = ctx.torch, ctx.dist
    import bench_kernels as bk
    from forest_benchmarking_b200 import distance_measures as dm
    from forest_benchmarking_b200.sharding import shard_range, all_gather_states
    d = 2 ** n
    lo, hi = shard_range(pairs_total, ctx.world, ctx.rank)
    B = hi - lo
    # generated on the device, per shard (8.2 GB of states for the whole job)
    rho, sig = bk._rand_states(torch, B, d, 5005 + 2 * ctx.rank), bk._rand_states(torch, B, d, 5006 + 2 * ctx.rank)
    res = torch.empty((B, 2), dtype=torch.float64, device="cuda")
    f, t = torch.empty((B,), dtype=torch.float64, device="cuda"), torch.empty((B,), dtype=torch.float64, device="cuda")
    full_host = torch.empty((pairs_total, 2), dtype=torch.float64).pin_memory()
    rho_h = sig_h = None
    ms_k = {}

    def step():
        dm.fidelity_batch(rho, sig, out=f)
        dm.trace_distance_batch(rho, sig, out=t)
        res[:, 0], res[:, 1] = f, t
        return all_gather_states(res, pairs_total)

    sampler = ClockSampler(ctx.local).start() if ctx.rank == 0 else None
    ms, _ = ctx.timed(step, ctx.args.steps, ctx.args.warmup)
    clocks = sampler.stop() if sampler else None
    ms_k["fidelity"], _ = ctx.timed(lambda: dm.fidelity_batch(rho, sig, out=f), 3, 1)
    ms_k["trace_distance"], _ = ctx.timed(lambda: dm.trace_distance_batch(rho, sig, out=t), 3, 1)
    # end to end from pinned host memory: only at a size the host can hold comfortably (2 x B x 4 KB)
    Be = min(B, 1 << 17)
    rho_h, sig_h = rho[:Be].cpu().pin_memory(), sig[:Be].cpu().pin_memory()

    def step_e2e():
        r, s = rho_h.cuda(non_blocking=True), sig_h.cuda(non_blocking=True)
        o = torch.stack([dm.fidelity_batch(r, s), dm.trace_distance_batch(r, s)], dim=1)
        full_host[:Be].copy_(o, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms_e2e, _ = ctx.timed(step_e2e, 3, 1)
    if ctx.rank != 0:
        return None
    bytes_pair = 32 * d * d + 8
    obj = {"metric": "state-pair distance evaluations/sec (fidelity + trace_distance per pair)",
           "value": pairs_total / (ms * 1e-3), "unit": "pairs/s", "n_gpus": ctx.world, "ms_per_step": ms,
           "steps": ctx.args.steps, "warmup": ctx.args.warmup, "scaling": "strong", "higher_is_better": True,
           "dtype": "f64 (complex128)", "data": "synthetic (Ginibre states generated on the device per shard)",
           "config": {"workload": f"fidelity + trace_distance over {pairs_total} random {n}-qubit density-matrix pairs "
                                  "(BASELINE configs[4])", "pairs_per_gpu": B,
                      "collective": "one all_gather of the [pairs, 2] results" if ctx.world > 1 else "none",
                      "l2": "inputs (2 x 4 KB per pair) far larger than the 126 MB L2, and flushed between steps"},
           "e2e": {"value": ctx.world * Be / (ms_e2e * 1e-3), "unit": "pairs/s", "pairs_per_gpu": Be, "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": int(2 * Be * 16 * d * d), "d2h_bytes_per_step": int(Be * 16),
                   "note": "bounded by the host->device copy of 8 KB per pair"},
           "gpu_launches": 2 * ctx.args.steps + 2 * ctx.args.steps, "clocks": clocks,
           "kernels": {k: {"ms": v, "pairs_per_s": ctx.world * B / (v * 1e-3),
                           "hbm_gbs": B * bytes_pair / (v * 1e-3) / 1e9,
                           "frac_of_hbm_peak": B * bytes_pair / (v * 1e-3) / 1e9 / ctx.hbm_peak} for k, v in ms_k.items()},
           "roofline": {"kernel": "fidelity_tri_kernel<16> (16 lanes per pair: Cholesky, L^dagger sigma L, Householder "
                                   "tridiagonalisation; then one lane per pair: square-root-free QL)",
                        "bound": "hbm", "achieved": B * bytes_pair / (ms_k["fidelity"] * 1e-3) / 1e9, "peak": ctx.hbm_peak,
                        "unit": "GB/s", "frac": B * bytes_pair / (ms_k["fidelity"] * 1e-3) / 1e9 / ctx.hbm_peak,
                        "peak_source": ctx.hbm_src, "traffic": kernel_traffic("fidelity_tri_kernel<16>")[0],
                        "traffic_source": kernel_traffic("fidelity_tri_kernel<16>")[1],
                        "note": "the fidelity kernel is instruction-issue / instruction-cache bound (~4.4 k warp "
                                "instructions per pair, profiles/r02_ubench_fidelity_tri.txt), the trace-distance "
                                "kernel is the HBM-bound one: see kernels.trace_distance"}}
    if ctx.world == 1 and not ctx.args.no_cpu_baseline:
        obj["cpu_baseline"] = bk.cpu_distance_baseline(torch)
    return obj


def run_ours(args):
    ctx = Ctx(args)
    parts = DEFAULT_PARTS if args.workload == "all" else (args.workload,)
    line, extra = None, {}
    for p in parts:
        if p == "mle2q":
            line = part_mle2q(ctx)
        elif p == "mle_batch_sweep":
            extra[p] = part_mle_sweep(ctx)
        elif p == "mle3q":
            extra[p] = part_mle3q(ctx)
        elif p.startswith("pgdb"):
            extra[p] = part_pgdb(ctx, n={"pgdb3q": 3, "pgdb2q": 2, "pgdb1q": 1}[p],
                                 global_batch=args.pgdb_batch, in_basis=args.in_basis)
        elif p == "distances":
            extra[p] = part_distances(ctx)
        ctx.torch.cuda.empty_cache()
    if ctx.rank == 0:
        if line is None:  # single-part run: that part is the headline
            (key, line), = extra.items()
            extra = {}
            line.setdefault("vs_baseline", None)
        line.update(extra)
        print(json.dumps(line))
    if ctx.dist:
        ctx.dist.destroy_process_group()


def run_kernels(args):
    """--workload streaming | convert | next: per-kernel HBM-roofline tables (bench_kernels.py)."""
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
    else:
        rows = bk.convert_rows(torch, peak)
        headline = next(r for r in rows if r["kernel"].startswith("superop2pauli_liouville n=3"))
        workload = "BASELINE configs[3]: conversion sweep kraus<->choi<->pauli_liouville, n=1..5, batch 16384 (chunked)"
    clocks = sampler.stop()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "convert":
        cpu = bk.cpu_convert_baseline(torch)
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
    ap.add_argument("--workload", default="all", choices=["all", "mle2q", "pgdb3q", "pgdb2q", "pgdb1q", "distances",
                                                          "mle_batch_sweep", "mle3q", "streaming", "convert", "next"])
    ap.add_argument("--in-basis", default="pauli", choices=["pauli", "sic"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="MLE experiments per GPU")
    ap.add_argument("--pgdb-batch", type=int, default=1024, help="GLOBAL process-tomography batch (sharded over the ranks)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pgdb-cpu", action="store_true", help="skip the 2-experiment CPU port of 3-qubit PGDB (~80 s)")
    ap.add_argument("--eigh-tol", type=float, default=None,
                    help="eigh_rel_tol passed to pgdb_process_estimate_batch (default: library default 1e-8); reported")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload in ("streaming", "convert", "next"):
        run_kernels(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
