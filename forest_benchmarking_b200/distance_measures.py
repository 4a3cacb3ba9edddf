"""Distance measures between quantum states -- signatures of forest/benchmarking/distance_measures.py,
computed by the batched kernels of csrc/qt_distance.cu (one pair per warp)."""
import ctypes

import numpy as np

from . import _lib


def _n_qubits(dim):
    n = int(round(np.log2(dim)))
    if 2 ** n != dim or not 1 <= n <= 5:
        raise ValueError(f"dimension {dim} is not 2^n with 1 <= n <= 5")
    return n


def _pair_call(fn_name, rho, sigma, out=None):
    torch = _lib.require_cuda()
    if rho.shape != sigma.shape or rho.dim() != 3 or rho.dtype != torch.complex128:
        raise ValueError("rho and sigma must be complex128 CUDA tensors of identical shape [B, d, d]")
    rho, sigma = rho.contiguous(), sigma.contiguous()
    b, d = rho.shape[0], rho.shape[1]
    if out is None:
        out = torch.empty((b,), dtype=torch.float64, device=rho.device)
    fn = getattr(_lib.lib(), fn_name)
    _lib.check(fn(ctypes.c_int(_n_qubits(d)), ctypes.c_int64(b), _lib.ptr(rho), _lib.ptr(sigma), _lib.ptr(out),
                  _lib.current_stream_ptr()), fn_name)
    return out


def fidelity_batch(rho, sigma, out=None):
    """[B,d,d] x [B,d,d] -> [B] fidelities (reference distance_measures.py:64-84 per pair)."""
    return _pair_call("qt_fidelity_batch", rho, sigma, out)


def trace_distance_batch(rho, sigma, out=None):
    """Reference semantics (distance_measures.py:114): 0.5 * induced 1-norm = 0.5 max_j sum_i |d_ij|."""
    return _pair_call("qt_trace_distance_batch", rho, sigma, out)


def trace_distance_nuclear_batch(rho, sigma, out=None):
    """Textbook trace distance 0.5 * sum |eig(rho - sigma)| (extra; NOT what the reference computes)."""
    return _pair_call("qt_trace_distance_nuclear_batch", rho, sigma, out)


def purity_batch(rho, out=None):
    torch = _lib.require_cuda()
    rho = rho.contiguous()
    b, d = rho.shape[0], rho.shape[1]
    if out is None:
        out = torch.empty((b,), dtype=torch.float64, device=rho.device)
    _lib.check(_lib.lib().qt_purity_batch(ctypes.c_int(_n_qubits(d)), ctypes.c_int64(b), _lib.ptr(rho),
                                          _lib.ptr(out), _lib.current_stream_ptr()), "qt_purity_batch")
    return out


def _one(x):
    torch = _lib.require_cuda()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.complex128))[None]).cuda()


def fidelity(rho: np.ndarray, sigma: np.ndarray, tol: float = 1000) -> float:
    """Drop-in for reference distance_measures.py:64-84."""
    return float(fidelity_batch(_one(rho), _one(sigma)).item())


def infidelity(rho: np.ndarray, sigma: np.ndarray, tol: float = 1000) -> float:
    """reference distance_measures.py:87-97."""
    return 1 - fidelity(rho, sigma, tol)


def trace_distance(rho: np.ndarray, sigma: np.ndarray) -> float:
    """Drop-in for reference distance_measures.py:100-114 (induced 1-norm, sic)."""
    return float(trace_distance_batch(_one(rho), _one(sigma)).item())


def purity(rho: np.ndarray, dim_renorm=True, tol: float = 1000) -> float:
    """reference distance_measures.py:14-37."""
    p = float(purity_batch(_one(rho)).item())
    if dim_renorm:
        d = np.asarray(rho).shape[0]
        p = (d / (d - 1.0)) * (p - 1.0 / d)
    return p


def hilbert_schmidt_ip_batch(a, b, out=None):
    """tr(A^dagger B) for every pair of [B, r, c] complex128 CUDA tensors -> complex128 [B]."""
    torch = _lib.require_cuda()
    if a.shape != b.shape or a.dim() != 3 or a.dtype != torch.complex128 or b.dtype != torch.complex128:
        raise ValueError("a and b must be complex128 CUDA tensors of identical shape [B, r, c]")
    a, b = a.contiguous(), b.contiguous()
    if out is None:
        out = torch.empty((a.shape[0],), dtype=torch.complex128, device=a.device)
    _lib.check(_lib.lib().qt_hs_inner_batch(ctypes.c_int64(a.shape[1]), ctypes.c_int64(a.shape[2]),
                                            ctypes.c_int64(a.shape[0]), _lib.ptr(a), _lib.ptr(b), _lib.ptr(out),
                                            _lib.current_stream_ptr()), "qt_hs_inner_batch")
    return out


def _real_if_close_item(z, tol=1000):
    return np.ndarray.item(np.real_if_close(np.asarray(z), tol))


def hilbert_schmidt_ip(A: np.ndarray, B: np.ndarray, tol: float = 1000) -> float:
    """reference distance_measures.py:198-216."""
    return _real_if_close_item(hilbert_schmidt_ip_batch(_one(A), _one(B)).cpu().numpy()[0], tol)


def entanglement_fidelity_batch(pl0, pl1):
    """F_e = tr(E^dagger F) / dim^2 for Pauli-Liouville matrices [B, dim^2, dim^2] -> complex128 [B]."""
    return hilbert_schmidt_ip_batch(pl0, pl1) / pl0.shape[1]


def entanglement_fidelity(pauli_lio0: np.ndarray, pauli_lio1: np.ndarray, tol: float = 1000) -> float:
    """reference distance_measures.py:271-303."""
    assert pauli_lio0.shape == pauli_lio1.shape
    assert pauli_lio0.shape[0] == pauli_lio1.shape[1]
    dim = int(np.sqrt(pauli_lio0.shape[0]))
    fe = hilbert_schmidt_ip_batch(_one(pauli_lio0), _one(pauli_lio1)).cpu().numpy()[0] / (dim ** 2)
    return _real_if_close_item(fe, tol)


def process_fidelity(pauli_lio0: np.ndarray, pauli_lio1: np.ndarray) -> float:
    """reference distance_measures.py:306-361: (dim F_e + 1) / (dim + 1)."""
    assert pauli_lio0.shape == pauli_lio1.shape
    assert pauli_lio0.shape[0] == pauli_lio1.shape[1]
    dim = int(np.sqrt(pauli_lio0.shape[0]))
    return (dim * entanglement_fidelity(pauli_lio0, pauli_lio1) + 1) / (dim + 1)


def process_infidelity(pauli_lio0: np.ndarray, pauli_lio1: np.ndarray) -> float:
    """reference distance_measures.py:364-375."""
    return 1 - process_fidelity(pauli_lio0, pauli_lio1)
