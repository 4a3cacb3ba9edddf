"""State and process tomography estimators -- same call signatures as
forest/benchmarking/tomography.py, executed by the sm_100a kernels in csrc/.

``iterative_mle_state_estimate`` (reference :168-270), ``pgdb_process_estimate`` (:542-594) and the
linear-inversion estimators (:130-165, :459-491) keep their signatures; ``*_batch`` twins take a whole
batch of independent experiments (the data-parallel axis) as dense arrays / device tensors.
"""
import ctypes
import warnings
from typing import List, Sequence

import numpy as np

from . import _lib
from .utils import flatten_process_results, flatten_state_results

MAXITER = "maxiter"
OPTIMAL = "optimal"
FRO = "fro"

KERNEL_AUTO, KERNEL_REGISTER, KERNEL_WARP, KERNEL_QUAD = 0, 1, 2, 3
DEFAULT_MAXITER = 10_000  # iterative_mle_state_estimate's default (reference tomography.py:170)


class MlePlan:
    """Device-side description of one list of observables (shared by every experiment of a batch)."""

    def __init__(self, n_qubits: int, pauli_idx, coeffs=None):
        torch = _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device())  # the plan's tables live here
        self.n = int(n_qubits)
        idx = np.ascontiguousarray(pauli_idx, dtype=np.int32)
        cf = np.ones(len(idx)) if coeffs is None else np.ascontiguousarray(coeffs, dtype=np.float64)
        if idx.ndim != 1 or cf.shape != idx.shape:
            raise ValueError("pauli_idx and coeffs must be 1-D with one entry per result")
        self.K = len(idx)
        self.pauli_idx, self.coeffs = idx, cf
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().qt_mle_plan_create(
            self.n, self.K, idx.ctypes.data_as(ctypes.c_void_p), cf.ctypes.data_as(ctypes.c_void_p),
            ctypes.byref(self._h)), "qt_mle_plan_create")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().qt_mle_plan_destroy(h)
            except Exception:
                pass


def iterative_mle_state_estimate_batch(plan: MlePlan, expectations, counts=None, epsilon=.1,
                                       entropy_penalty=0.0, beta=0.0, tol=1e-9, maxiter=DEFAULT_MAXITER,
                                       kernel=KERNEL_AUTO, out=None, iters_out=None):
    """Batched diluted MLE.  ``expectations`` / ``counts``: CUDA float64 tensors [B, K] (K = plan.K).
    Returns (rho [B, d, d] complex128 CUDA tensor, iterations [B] int32 CUDA tensor); iterations equals
    ``maxiter`` for experiments that hit the cap (where the reference warns, tomography.py:244-246)."""
    torch = _lib.require_cuda()
    if (entropy_penalty != 0.0) and (beta != 0.0):
        raise ValueError("One can't sensibly do entropy penalty and hedging. Do one or the other"
                         " but not both.")
    if expectations.dtype != torch.float64 or not expectations.is_cuda or expectations.dim() != 2:
        raise ValueError("expectations must be a CUDA float64 tensor of shape [B, K]")
    if expectations.shape[1] != plan.K:
        raise ValueError(f"expectations has {expectations.shape[1]} columns, plan has {plan.K}")
    expectations = expectations.contiguous()
    b = expectations.shape[0]
    d = 2 ** plan.n
    if counts is not None:
        if not counts.is_cuda or tuple(counts.shape) != (b, plan.K):
            raise ValueError(f"counts must be a CUDA tensor of shape [{b}, {plan.K}]")
        counts = counts.to(torch.float64).contiguous()
    elif beta != 0.0:
        raise ValueError("hedged MLE (beta != 0) needs the shot counts")
    dev = _lib.common_device(expectations, counts, out, iters_out, plan=plan)
    with _lib.on_device(dev):
        if out is None:
            out = torch.empty((b, d, d), dtype=torch.complex128, device=dev)
        else:
            _lib.check_tensor("out", out, torch.complex128, (b, d, d))
        if iters_out is None:
            iters_out = torch.empty((b,), dtype=torch.int32, device=dev)
        else:
            _lib.check_tensor("iters_out", iters_out, torch.int32, (b,))
        _lib.check(_lib.lib().qt_mle_state_batch(
            plan._h, ctypes.c_int64(b), _lib.ptr(expectations), _lib.ptr(counts), ctypes.c_double(epsilon),
            ctypes.c_double(entropy_penalty), ctypes.c_double(beta), ctypes.c_double(tol), ctypes.c_int(maxiter),
            ctypes.c_int(kernel), _lib.ptr(out), _lib.ptr(iters_out), _lib.current_stream_ptr()),
            "qt_mle_state_batch")
    return out, iters_out


def linear_inv_state_estimate_batch(plan: MlePlan, expectations, out=None):
    """Batched linear inversion.  expectations: CUDA float64 [B, K] -> rho [B, d, d] complex128 (CUDA)."""
    torch = _lib.require_cuda()
    if expectations.dtype != torch.float64 or not expectations.is_cuda or expectations.dim() != 2 \
            or expectations.shape[1] != plan.K:
        raise ValueError(f"expectations must be a CUDA float64 tensor of shape [B, {plan.K}]")
    expectations = expectations.contiguous()
    b, d = expectations.shape[0], 2 ** plan.n
    dev = _lib.common_device(expectations, out, plan=plan)
    with _lib.on_device(dev):
        if out is None:
            out = torch.empty((b, d, d), dtype=torch.complex128, device=dev)
        else:
            _lib.check_tensor("out", out, torch.complex128, (b, d, d))
        _lib.check(_lib.lib().qt_linear_inv_state_batch(plan._h, ctypes.c_int64(b), _lib.ptr(expectations),
                                                        _lib.ptr(out), _lib.current_stream_ptr()),
                   "qt_linear_inv_state_batch")
    return out


def linear_inv_state_estimate(results: List, qubits: List[int]) -> np.ndarray:
    """Drop-in for reference tomography.py:130-165."""
    torch = _lib.require_cuda()
    idx, cf, ex, _ = flatten_state_results(results, qubits)
    plan = MlePlan(len(qubits), idx, cf)
    dev = torch.device("cuda", torch.cuda.current_device())
    return linear_inv_state_estimate_batch(plan, torch.from_numpy(ex[None, :]).to(dev))[0].cpu().numpy()


def state_log_likelihood_batch(plan: MlePlan, rho, expectations, counts, out=None):
    """Batched log10 likelihood.  rho [B, d, d] complex128, expectations / counts [B, K] float64 (CUDA) -> [B]."""
    torch = _lib.require_cuda()
    b, d = rho.shape[0], 2 ** plan.n
    if rho.dtype != torch.complex128 or not rho.is_cuda or tuple(rho.shape) != (b, d, d):
        raise ValueError(f"rho must be a CUDA complex128 tensor of shape [B, {d}, {d}]")
    for name, t in (("expectations", expectations), ("counts", counts)):
        if t.dtype != torch.float64 or not t.is_cuda or tuple(t.shape) != (b, plan.K):
            raise ValueError(f"{name} must be a CUDA float64 tensor of shape [{b}, {plan.K}]")
    rho, expectations, counts = rho.contiguous(), expectations.contiguous(), counts.contiguous()
    dev = _lib.common_device(rho, expectations, counts, out, plan=plan)
    with _lib.on_device(dev):
        if out is None:
            out = torch.empty((b,), dtype=torch.float64, device=dev)
        else:
            _lib.check_tensor("out", out, torch.float64, (b,))
        _lib.check(_lib.lib().qt_state_log_likelihood_batch(plan._h, ctypes.c_int64(b), _lib.ptr(rho),
                                                            _lib.ptr(expectations), _lib.ptr(counts), _lib.ptr(out),
                                                            _lib.current_stream_ptr()),
                   "qt_state_log_likelihood_batch")
    return out


def state_log_likelihood(state: np.ndarray, results: List, qubits: List[int]) -> float:
    """Drop-in for reference tomography.py:341-375."""
    torch = _lib.require_cuda()
    idx, cf, ex, cnt = flatten_state_results(results, qubits)
    plan = MlePlan(len(qubits), idx, cf)
    dev = torch.device("cuda", torch.cuda.current_device())
    rho = torch.from_numpy(np.ascontiguousarray(state, dtype=np.complex128)[None]).to(dev)
    return float(state_log_likelihood_batch(plan, rho, torch.from_numpy(ex[None, :]).to(dev),
                                            torch.from_numpy(cnt[None, :]).to(dev)).item())


def mle_step_batch(n_qubits: int, expect_canon, rho, epsilon=.1, out=None):
    """ONE R-rho-R update streamed through HBM (n = 1, 2; complete canonical Pauli set).
    expect_canon: [4^n - 1, B] float64 CUDA (item-minor); rho: [B, d, d] complex128 CUDA."""
    torch = _lib.require_cuda()
    n = int(n_qubits)
    if n not in (1, 2):
        raise ValueError("mle_step_batch supports n_qubits = 1, 2")
    d, k = 2 ** n, 4 ** n - 1
    if rho.dim() != 3:
        raise ValueError(f"rho must be a CUDA complex128 tensor of shape [B, {d}, {d}]")
    b = rho.shape[0]
    _lib.check_tensor("rho", rho, torch.complex128, (b, d, d))
    _lib.check_tensor("expect_canon", expect_canon, torch.float64, (k, b))
    dev = _lib.common_device(rho, expect_canon, out)
    with _lib.on_device(dev):
        if out is None:
            out = torch.empty_like(rho)
        else:
            _lib.check_tensor("out", out, torch.complex128, (b, d, d))
        _lib.check(_lib.lib().qt_mle_step_batch(ctypes.c_int(n), ctypes.c_int64(b), _lib.ptr(expect_canon),
                                                _lib.ptr(rho), ctypes.c_double(epsilon), _lib.ptr(out),
                                                _lib.current_stream_ptr()), "qt_mle_step_batch")
    return out


def iterative_mle_state_estimate(results: List, qubits: List[int], epsilon=.1, entropy_penalty=0.0,
                                 beta=0.0, tol=1e-9, maxiter=DEFAULT_MAXITER) -> np.ndarray:
    """Drop-in for reference tomography.py:168-270 (one experiment = a batch of one)."""
    torch = _lib.require_cuda()
    if (entropy_penalty != 0.0) and (beta != 0.0):
        raise ValueError("One can't sensibly do entropy penalty and hedging. Do one or the other"
                         " but not both.")
    idx, cf, ex, cnt = flatten_state_results(results, qubits)
    plan = MlePlan(len(qubits), idx, cf)
    dev = torch.device("cuda", torch.cuda.current_device())
    rho, iters = iterative_mle_state_estimate_batch(
        plan, torch.from_numpy(ex[None, :]).to(dev), torch.from_numpy(cnt[None, :]).to(dev),
        epsilon, entropy_penalty, beta, tol, maxiter)
    if int(iters.item()) >= maxiter:
        warnings.warn('Maximum number of iterations reached before convergence.')
    return rho[0].cpu().numpy()


def _resample_expectations_with_beta(expectations, counts, n_resamples, prior_counts=1):
    """reference tomography.py:378-409 for n_resamples bootstrap replicas at once: expectation -> (+1, -1)
    counts -> Beta(n+ + prior, n- + prior) -> expectation.  Draws come from NumPy's global RNG in the order the
    reference consumes them (replica-major, result-minor), so a seeded run reproduces the reference's samples."""
    num_plus = ((expectations + 1) / 2) * counts
    num_minus = counts - num_plus
    alpha = np.broadcast_to(num_plus + prior_counts, (n_resamples, len(expectations)))
    beta = np.broadcast_to(num_minus + prior_counts, (n_resamples, len(expectations)))
    return 2 * np.random.beta(alpha, beta) - 1


def estimate_variance(results: List, qubits: List[int], tomo_estimator, functional, target_state=None,
                      n_resamples: int = 40, project_to_physical: bool = False):
    """Drop-in for reference tomography.py:412-453 (bootstrap error bar on a functional of the state).
    The n_resamples replicas are ONE batch: when ``tomo_estimator`` is this module's
    ``iterative_mle_state_estimate`` / ``linear_inv_state_estimate`` and ``functional`` one of this package's
    distance measures, everything after the host-side resampling runs as batched kernels; any other callable
    is applied replica by replica like the reference does."""
    from . import distance_measures as dm
    from .operator_tools.project_state_matrix import project_state_matrix_to_physical_batch
    torch = _lib.require_cuda()
    if functional != dm.purity and target_state is None:
        raise ValueError("You're not using the `purity` functional. Please specify a target state.")
    idx, cf, ex, cnt = flatten_state_results(results, qubits)
    resampled = _resample_expectations_with_beta(ex, cnt, n_resamples)
    dev = torch.device("cuda", torch.cuda.current_device())
    if tomo_estimator in (iterative_mle_state_estimate, linear_inv_state_estimate):
        plan = MlePlan(len(qubits), idx, cf)
        e = torch.from_numpy(np.ascontiguousarray(resampled)).to(dev)
        if tomo_estimator is linear_inv_state_estimate:
            rho = linear_inv_state_estimate_batch(plan, e)
        else:
            c = torch.from_numpy(np.tile(cnt, (n_resamples, 1))).to(dev)
            maxiter = DEFAULT_MAXITER  # the reference calls the estimator with its defaults (tomography.py:442)
            rho, iters = iterative_mle_state_estimate_batch(plan, e, c, maxiter=maxiter)
            if bool((iters >= maxiter).any().item()):
                warnings.warn('Maximum number of iterations reached before convergence.')
    else:
        from .observable_estimation import ExperimentResult
        rhos = []
        for row in resampled:
            rs = [ExperimentResult(r.setting, float(x), r.total_counts, getattr(r, "std_err", None))
                  for r, x in zip(results, row)]
            rhos.append(np.asarray(tomo_estimator(rs, qubits), dtype=np.complex128))
        rho = torch.from_numpy(np.stack(rhos)).to(dev)
    if project_to_physical:
        rho = project_state_matrix_to_physical_batch(rho)
    if functional == dm.purity:
        sample = dm.purity_batch(rho).cpu().numpy()
    elif functional in (dm.fidelity, dm.infidelity, dm.trace_distance):
        tgt = torch.from_numpy(np.ascontiguousarray(np.asarray(target_state, dtype=np.complex128))).to(dev)
        tgt = tgt.expand(rho.shape[0], -1, -1).contiguous()
        if functional == dm.trace_distance:
            sample = dm.trace_distance_batch(tgt, rho).cpu().numpy()
        else:
            sample = dm.fidelity_batch(tgt, rho).cpu().numpy()
            if functional == dm.infidelity:
                sample = 1 - sample
    else:
        sample = np.array([np.real(functional(target_state, r)) for r in rho.cpu().numpy()])
    return np.mean(sample), np.var(sample)


# ==================================================================================================
# PROCESS tomography
# ==================================================================================================
class PgdbPlan:
    """Device-side description of one list of process-tomography settings (shared by the batch)."""

    def __init__(self, n_qubits: int, state_codes, pauli_idx, coeffs=None):
        torch = _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device())  # the plan's tables live here
        self.n = int(n_qubits)
        codes = np.ascontiguousarray(state_codes, dtype=np.int32).reshape(-1, self.n)
        idx = np.ascontiguousarray(pauli_idx, dtype=np.int32)
        cf = np.ones(len(idx)) if coeffs is None else np.ascontiguousarray(coeffs, dtype=np.float64)
        if len(codes) != len(idx) or len(cf) != len(idx):
            raise ValueError("state_codes, pauli_idx and coeffs must have one entry per setting")
        self.S = len(idx)
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().qt_pgdb_plan_create(
            self.n, self.S, codes.ctypes.data_as(ctypes.c_void_p), idx.ctypes.data_as(ctypes.c_void_p),
            cf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(self._h)), "qt_pgdb_plan_create")
        n_in, canon = ctypes.c_int32(), ctypes.c_int32()
        _lib.check(_lib.lib().qt_pgdb_plan_info(self._h, ctypes.byref(n_in), ctypes.byref(canon)), "qt_pgdb_plan_info")
        self.n_in, self.canonical = n_in.value, bool(canon.value)

    @classmethod
    def complete(cls, n_qubits: int, in_basis="pauli"):
        """The tomographically complete settings of generate_process_tomography_experiment
        (reference tomography.py:71-123): product(input states) x all traceless Paulis."""
        import itertools
        codes = range(0, 6) if in_basis.lower() == "pauli" else range(6, 10)
        k = 4 ** n_qubits - 1
        states = np.array(list(itertools.product(codes, repeat=n_qubits)), dtype=np.int32)
        return cls(n_qubits, np.repeat(states, k, axis=0), np.tile(np.arange(1, k + 1, dtype=np.int32), len(states)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().qt_pgdb_plan_destroy(h)
            except Exception:
                pass


STATUS_DYKSTRA_CAP, STATUS_JACOBI_CAP, STATUS_PGDB_CAP = 1, 2, 4  # QT_STATUS_* of include/qtomo.h


def warn_on_status(status, what):
    """The reference's loops are unbounded; the kernels carry safety caps (10000 Dykstra trips, 30 Jacobi sweeps,
    100000 PGD steps).  A capped item is returned but NOT converged: say so instead of passing it off as valid."""
    bits = int(np.bitwise_or.reduce(np.asarray(status, dtype=np.int64).ravel(), initial=0))
    if bits & STATUS_PGDB_CAP:
        warnings.warn(f"{what}: projected gradient descent stopped at its iteration cap before convergence.")
    if bits & STATUS_DYKSTRA_CAP:
        warnings.warn(f"{what}: a proj_choi_to_physical call stopped at its iteration cap before convergence.")
    if bits & STATUS_JACOBI_CAP:
        warnings.warn(f"{what}: an eigendecomposition used all its Jacobi sweeps; the result may be inaccurate.")
    return bits


def pgdb_process_estimate_batch(plan: PgdbPlan, expectations, counts, trace_preserving=True, out=None,
                                return_counters=False, workspace=None, eigh_rel_tol=None, return_status=False):
    """Batched PGDB.  expectations / counts: CUDA float64 [B, S].  Returns choi [B, 4^n, 4^n] complex128
    (and, optionally, int32 [B, 4] counters: outer iterations, cost evaluations, eigh calls, Jacobi sweeps; and
    int32 [B] status words, non-zero where a safety cap was hit -- see ``warn_on_status``).
    ``eigh_rel_tol``: stopping tolerance of the eigensolver behind the CP projection (None = library default 1e-8,
    0 = tight); a per-call argument, nothing global."""
    torch = _lib.require_cuda()
    for t in (expectations, counts):
        if t.dtype != torch.float64 or not t.is_cuda or t.dim() != 2 or t.shape[1] != plan.S:
            raise ValueError(f"expectations and counts must be CUDA float64 tensors of shape [B, {plan.S}]")
    if expectations.shape != counts.shape:
        raise ValueError("expectations and counts must have the same shape")
    expectations, counts = expectations.contiguous(), counts.contiguous()
    b, m = expectations.shape[0], 4 ** plan.n
    lib = _lib.lib()
    dev = _lib.common_device(expectations, counts, out, workspace, plan=plan)
    with _lib.on_device(dev):
        if out is None:
            out = torch.empty((b, m, m), dtype=torch.complex128, device=dev)
        else:
            _lib.check_tensor("out", out, torch.complex128, (b, m, m))
        counters = torch.zeros((b, 4), dtype=torch.int32, device=dev)
        status = torch.zeros((b,), dtype=torch.int32, device=dev)
        nbytes = int(lib.qt_pgdb_workspace_bytes(plan._h, ctypes.c_int64(b)))
        if workspace is None or workspace.numel() * workspace.element_size() < nbytes:
            workspace = torch.empty((max(nbytes, 8) // 8,), dtype=torch.float64, device=dev)
        _lib.check(lib.qt_pgdb_process_batch(plan._h, ctypes.c_int64(b), _lib.ptr(expectations), _lib.ptr(counts),
                                             ctypes.c_int(1 if trace_preserving else 0),
                                             ctypes.c_double(-1.0 if eigh_rel_tol is None else eigh_rel_tol),
                                             _lib.ptr(out), _lib.ptr(counters), _lib.ptr(status), _lib.ptr(workspace),
                                             ctypes.c_int64(workspace.numel() * workspace.element_size()),
                                             _lib.current_stream_ptr()), "qt_pgdb_process_batch")
    res = (out,)
    if return_counters:
        res += (counters,)
    if return_status:
        res += (status,)
    return res if len(res) > 1 else out


def pgdb_process_estimate(results: List, qubits: List[int], trace_preserving=True) -> np.ndarray:
    """Drop-in for reference tomography.py:542-594."""
    torch = _lib.require_cuda()
    codes, idx, cf, ex, cnt = flatten_process_results(results, qubits)
    plan = PgdbPlan(len(qubits), codes, idx, cf)
    dev = torch.device("cuda", torch.cuda.current_device())
    choi, status = pgdb_process_estimate_batch(plan, torch.from_numpy(ex[None, :]).to(dev),
                                               torch.from_numpy(cnt[None, :]).to(dev), trace_preserving,
                                               return_status=True)
    warn_on_status(status.cpu().numpy(), "pgdb_process_estimate")
    return choi[0].cpu().numpy()


def linear_inv_process_estimate_batch(plan: PgdbPlan, expectations, out=None):
    """Batched linear-inversion process tomography.  expectations: CUDA float64 [B, S] -> choi [B, 4^n, 4^n]."""
    torch = _lib.require_cuda()
    if expectations.dtype != torch.float64 or not expectations.is_cuda or expectations.dim() != 2 \
            or expectations.shape[1] != plan.S:
        raise ValueError(f"expectations must be a CUDA float64 tensor of shape [B, {plan.S}]")
    expectations = expectations.contiguous()
    b, m = expectations.shape[0], 4 ** plan.n
    dev = _lib.common_device(expectations, out, plan=plan)
    with _lib.on_device(dev):
        if out is None:
            out = torch.empty((b, m, m), dtype=torch.complex128, device=dev)
        else:
            _lib.check_tensor("out", out, torch.complex128, (b, m, m))
        _lib.check(_lib.lib().qt_linear_inv_process_batch(plan._h, ctypes.c_int64(b), _lib.ptr(expectations),
                                                          _lib.ptr(out), _lib.current_stream_ptr()),
                   "qt_linear_inv_process_batch")
    return out


def linear_inv_process_estimate(results: List, qubits: List[int]) -> np.ndarray:
    """Drop-in for reference tomography.py:459-491."""
    torch = _lib.require_cuda()
    codes, idx, cf, ex, _ = flatten_process_results(results, qubits)
    plan = PgdbPlan(len(qubits), codes, idx, cf)
    dev = torch.device("cuda", torch.cuda.current_device())
    return linear_inv_process_estimate_batch(plan, torch.from_numpy(ex[None, :]).to(dev))[0].cpu().numpy()
