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
    if rho.shape[1] != rho.shape[2]:
        raise ValueError("rho and sigma must be square matrices [B, d, d]")
    rho, sigma = rho.contiguous(), sigma.contiguous()
    b, d = rho.shape[0], rho.shape[1]
    n = _n_qubits(d)
    dev = _lib.common_device(rho, sigma, out)
    with _lib.on_device(dev):
        if out is None:
            out = torch.empty((b,), dtype=torch.float64, device=dev)
        else:
            _lib.check_tensor("out", out, torch.float64, (b,))
        fn = getattr(_lib.lib(), fn_name)
        _lib.check(fn(ctypes.c_int(n), ctypes.c_int64(b), _lib.ptr(rho), _lib.ptr(sigma), _lib.ptr(out),
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
    """tr(rho rho) (un-renormalised, the reference's default) for [B, d, d] complex128 CUDA states -> [B]."""
    torch = _lib.require_cuda()
    if rho.dim() != 3 or rho.shape[1] != rho.shape[2] or rho.dtype != torch.complex128 or not rho.is_cuda:
        raise ValueError("rho must be a complex128 CUDA tensor of shape [B, d, d]")
    rho = rho.contiguous()
    b, d = rho.shape[0], rho.shape[1]
    n = _n_qubits(d)
    dev = _lib.common_device(rho, out)
    with _lib.on_device(dev):
        if out is None:
            out = torch.empty((b,), dtype=torch.float64, device=dev)
        else:
            _lib.check_tensor("out", out, torch.float64, (b,))
        _lib.check(_lib.lib().qt_purity_batch(ctypes.c_int(n), ctypes.c_int64(b), _lib.ptr(rho),
                                              _lib.ptr(out), _lib.current_stream_ptr()), "qt_purity_batch")
    return out


def _one(x):
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.complex128))[None]).to(dev)


def _real_if_close_item(z, tol=1000):
    """The reference's return idiom (distance_measures.py:37, 61, 84, 216, 303)."""
    return np.ndarray.item(np.real_if_close(np.asarray(z), tol))


def fidelity(rho: np.ndarray, sigma: np.ndarray, tol: float = 1000) -> float:
    """Drop-in for reference distance_measures.py:64-84: ``np.ndarray.item(np.real_if_close(fid, tol))``.

    The reference's ``fid`` is ``trace(sqrtm_psd(M))**2``; sqrtm_psd (calculational.py:77-91) rebuilds the root from
    the REAL eigenvalues of ``eigh``, so the trace is ``sum_i sqrt(max(w_i, 0))`` plus an imaginary part that is
    pure rounding noise of the ``(v * w) @ v^dagger`` product (~1e-17).  The kernel returns that real sum squared, so
    the value handed to ``real_if_close`` has an exactly-zero imaginary part and the result is a float for every
    ``tol``; the reference can only return a complex number when ``tol`` is set below its own rounding noise
    (tol << 1), in which case the two agree in the real part."""
    fid = np.complex128(fidelity_batch(_one(rho), _one(sigma)).item())
    return _real_if_close_item(fid, tol)


def infidelity(rho: np.ndarray, sigma: np.ndarray, tol: float = 1000) -> float:
    """reference distance_measures.py:87-97."""
    return 1 - fidelity(rho, sigma, tol)


def trace_distance(rho: np.ndarray, sigma: np.ndarray) -> float:
    """Drop-in for reference distance_measures.py:100-114 (induced 1-norm, sic)."""
    return float(trace_distance_batch(_one(rho), _one(sigma)).item())


def purity(rho: np.ndarray, dim_renorm=False, tol: float = 1000) -> float:
    """reference distance_measures.py:14-37: tr(rho rho), optionally renormalised to [0, 1] (default: NOT)."""
    p = np.complex128(hilbert_schmidt_ip_batch(_one(np.asarray(rho).conj().T), _one(rho)).item())
    if dim_renorm:
        dim = np.asarray(rho).shape[0]
        p = (dim / (dim - 1.0)) * (p - 1.0 / dim)
    return _real_if_close_item(p, tol)


def impurity(rho: np.ndarray, dim_renorm=False, tol: float = 1000) -> float:
    """reference distance_measures.py:40-61: 1 - tr(rho rho), optionally renormalised to [0, 1] (default: NOT)."""
    imp = 1 - np.complex128(hilbert_schmidt_ip_batch(_one(np.asarray(rho).conj().T), _one(rho)).item())
    if dim_renorm:
        dim = np.asarray(rho).shape[0]
        imp = (dim / (dim - 1.0)) * imp
    return _real_if_close_item(imp, tol)


def hilbert_schmidt_ip_batch(a, b, out=None):
    """tr(A^dagger B) for every pair of [B, r, c] complex128 CUDA tensors -> complex128 [B]."""
    torch = _lib.require_cuda()
    if a.shape != b.shape or a.dim() != 3 or a.dtype != torch.complex128 or b.dtype != torch.complex128:
        raise ValueError("a and b must be complex128 CUDA tensors of identical shape [B, r, c]")
    a, b = a.contiguous(), b.contiguous()
    dev = _lib.common_device(a, b, out)
    with _lib.on_device(dev):
        if out is None:
            out = torch.empty((a.shape[0],), dtype=torch.complex128, device=dev)
        else:
            _lib.check_tensor("out", out, torch.complex128, (a.shape[0],))
        _lib.check(_lib.lib().qt_hs_inner_batch(ctypes.c_int64(a.shape[1]), ctypes.c_int64(a.shape[2]),
                                                ctypes.c_int64(a.shape[0]), _lib.ptr(a), _lib.ptr(b), _lib.ptr(out),
                                                _lib.current_stream_ptr()), "qt_hs_inner_batch")
    return out


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
