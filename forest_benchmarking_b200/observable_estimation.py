"""Input record types of the tomography estimators (the argument types of the drop-in boundary).

Mirrors the fields the reference estimators read from forest/benchmarking/observable_estimation.py:
``_OneQState`` (:36-74), ``TensorProductState`` (:77-128), the state factories (:131-172),
``ExperimentSetting`` (:175-213) and ``ExperimentResult`` (:694-733).  Data acquisition, grouping and
calibration (the rest of that module) drive a QVM/QPU and are out of scope.  The reference's own
objects are accepted wherever these are (duck typing).
"""
from dataclasses import dataclass

import numpy as np
from typing import Iterable, Tuple, Union


@dataclass(frozen=True)
class _OneQState:
    label: str   # 'X', 'Y', 'Z' or 'SIC'
    index: int   # 0 = plus eigenstate, 1 = minus (X/Y/Z); 0..3 for SIC
    qubit: int

    def __str__(self):
        if self.label in ("X", "Y", "Z"):
            return f"{self.label}{'+-'[self.index]}_{self.qubit}"
        return f"{self.label}{self.index}_{self.qubit}"


class TensorProductState:
    def __init__(self, states: Iterable[_OneQState] = ()):
        self.states = tuple(states)

    def __mul__(self, other):
        return TensorProductState(self.states + other.states)

    def __getitem__(self, qubit):
        for s in self.states:
            if s.qubit == qubit:
                return s
        raise IndexError(qubit)

    def __iter__(self):
        return iter(self.states)

    def __len__(self):
        return len(self.states)

    def __eq__(self, other):
        return isinstance(other, TensorProductState) and frozenset(self.states) == frozenset(other.states)

    def __hash__(self):
        return hash(frozenset(self.states))

    def __repr__(self):
        return "TensorProductState[" + " * ".join(str(s) for s in self.states) + "]"


def _factory(label, index):
    def make(q):
        return TensorProductState((_OneQState(label, index, q),))
    return make


plusX, minusX = _factory("X", 0), _factory("X", 1)
plusY, minusY = _factory("Y", 0), _factory("Y", 1)
plusZ, minusZ = _factory("Z", 0), _factory("Z", 1)
SIC0, SIC1, SIC2, SIC3 = (_factory("SIC", i) for i in range(4))


def zeros_state(qubits: Iterable[int]):
    return TensorProductState(_OneQState("Z", 0, q) for q in qubits)


@dataclass(frozen=True)
class ExperimentSetting:
    in_state: TensorProductState
    observable: object  # PauliTerm (ours or pyquil's)


@dataclass(frozen=True)
class ExperimentResult:
    setting: ExperimentSetting
    expectation: Union[float, complex]
    total_counts: int
    std_err: Union[float, complex] = None
    raw_expectation: Union[float, complex] = None
    raw_std_err: float = None
    calibration_expectation: Union[float, complex] = None
    calibration_std_err: Union[float, complex] = None
    calibration_counts: int = None


# --------------------------------------------------------------------------------------------------
# raw shots -> moments (the producer of ExperimentResult.expectation / std_err)
# --------------------------------------------------------------------------------------------------
def shots_to_obs_moments_batch(bitarrays, col_masks, coeffs=None, use_beta_dist_unbiased_prior: bool = False):
    """``shots_to_obs_moments`` for B settings in one launch.

    bitarrays: CUDA uint8 [B, n_shots, n_qubits] of 0/1; col_masks: CUDA int32 [B], bit q set when column q of the
    bitarray belongs to the setting's observable (0 = identity term); coeffs: CUDA float64 [B] (default 1).
    Returns (mean [B], var [B]) float64 CUDA tensors (var = variance of the mean, as in the reference).
    """
    from . import _lib
    import ctypes
    torch = _lib.require_cuda()
    if bitarrays.dtype != torch.uint8 or not bitarrays.is_cuda or bitarrays.dim() != 3:
        raise ValueError("bitarrays must be a CUDA uint8 tensor of shape [B, n_shots, n_qubits]")
    b, s, q = bitarrays.shape
    if col_masks.dtype != torch.int32 or not col_masks.is_cuda or tuple(col_masks.shape) != (b,):
        raise ValueError(f"col_masks must be a CUDA int32 tensor of shape [{b}]")
    if coeffs is None:
        coeffs = torch.ones((b,), dtype=torch.float64, device=bitarrays.device)
    if coeffs.dtype != torch.float64 or not coeffs.is_cuda or tuple(coeffs.shape) != (b,):
        raise ValueError(f"coeffs must be a CUDA float64 tensor of shape [{b}]")
    bitarrays, col_masks, coeffs = bitarrays.contiguous(), col_masks.contiguous(), coeffs.contiguous()
    with _lib.on_device(_lib.common_device(bitarrays, col_masks, coeffs)):
        mean = torch.empty((b,), dtype=torch.float64, device=bitarrays.device)
        var = torch.empty_like(mean)
        _lib.check(_lib.lib().qt_shots_to_obs_moments_batch(
            ctypes.c_int64(b), ctypes.c_int64(s), ctypes.c_int(q), _lib.ptr(bitarrays), _lib.ptr(col_masks),
            _lib.ptr(coeffs), ctypes.c_int(1 if use_beta_dist_unbiased_prior else 0), _lib.ptr(mean), _lib.ptr(var),
            _lib.current_stream_ptr()), "qt_shots_to_obs_moments_batch")
    return mean, var


def shots_to_obs_moments(bitarray: np.ndarray, qubits, observable, use_beta_dist_unbiased_prior: bool = False):
    """Drop-in for reference observable_estimation.py:804-853 (one setting = a batch of one)."""
    from . import _lib
    torch = _lib.require_cuda()
    coeff = complex(observable.coefficient)
    if not np.isclose(coeff.imag, 0):
        raise ValueError(f"The coefficient of an observable should not be complex.")
    coeff = coeff.real
    obs_qubits = [q for q, _ in observable]
    idxs = [idx for idx, q in enumerate(qubits) if q in obs_qubits]
    if len(idxs) == 0:  # identity term
        return coeff, 0
    bitarray = np.asarray(bitarray)
    assert bitarray.shape[1] == len(qubits), 'qubits should label each column of the bitarray'
    if len(qubits) > 32:  # the kernel addresses columns through a 32-bit mask: keep the observable's columns only
        bitarray, idxs = bitarray[:, idxs], list(range(len(idxs)))
        if len(idxs) > 32:
            raise ValueError("observables on more than 32 qubits are not supported")
    dev = torch.device("cuda", torch.cuda.current_device())
    mask = sum(1 << i for i in idxs)
    mask = mask - (1 << 32) if mask >= (1 << 31) else mask
    mean, var = shots_to_obs_moments_batch(
        torch.from_numpy(np.ascontiguousarray(bitarray, dtype=np.uint8)[None]).to(dev),
        torch.tensor([mask], dtype=torch.int32, device=dev),
        torch.tensor([coeff], dtype=torch.float64, device=dev), use_beta_dist_unbiased_prior)
    return float(mean.item()), float(var.item())


def ratio_variance(a, var_a, b, var_b):
    """reference observable_estimation.py:1052-1090."""
    return var_a / b ** 2 + (a ** 2 * var_b) / b ** 4


def calibrate_estimates_batch(mean, var, cal_mean, cal_var):
    """The arithmetic of ``calibrate_observable_estimates`` (reference :1033-1049) for B results at once:
    corrected mean = mean / cal_mean, corrected variance = ratio_variance(mean, var, cal_mean, cal_var).
    All arguments CUDA float64 [B]; returns (corrected_mean, corrected_var)."""
    from . import _lib
    import ctypes
    torch = _lib.require_cuda()
    ts = [t.contiguous() for t in (mean, var, cal_mean, cal_var)]
    b = ts[0].shape[0]
    for t in ts:
        if t.dtype != torch.float64 or not t.is_cuda or tuple(t.shape) != (b,):
            raise ValueError(f"all arguments must be CUDA float64 tensors of shape [{b}]")
    with _lib.on_device(_lib.common_device(*ts)):
        om, ov = torch.empty_like(ts[0]), torch.empty_like(ts[0])
        _lib.check(_lib.lib().qt_calibrate_estimates_batch(ctypes.c_int64(b), *[_lib.ptr(t) for t in ts], _lib.ptr(om),
                                                           _lib.ptr(ov), _lib.current_stream_ptr()),
                   "qt_calibrate_estimates_batch")
    return om, ov
