"""Superoperator representation changes -- signatures of
forest/benchmarking/operator_tools/superoperator_transformations.py, computed by csrc/qt_convert.cu.

``*_batch`` functions take/return complex128 CUDA tensors [B, rows, cols]; the reference-named
functions take/return numpy arrays (one matrix = a batch of one).  Column-stacking ``vec`` throughout.
"""
import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .. import _lib

__all__ = ["vec", "unvec", "kraus2choi", "kraus2superop", "kraus2pauli_liouville", "choi2superop",
           "superop2choi", "superop2pauli_liouville", "pauli_liouville2superop", "choi2pauli_liouville",
           "pauli_liouville2choi", "choi2kraus", "choi2kraus_batch", "kraus2chi", "chi2choi", "chi2pauli_liouville",
           "chi2superop", "chi2kraus", "choi2chi", "superop2chi", "pauli_liouville2chi", "superop2kraus",
           "pauli_liouville2kraus", "kraus2chi_batch", "chi2choi_batch", "kraus2choi_batch", "kraus2superop_batch", "reshuffle_batch",
           "superop2pauli_liouville_batch", "pauli_liouville2superop_batch",
           "choi2pauli_liouville_batch", "pauli_liouville2choi_batch", "pauli2computational_basis_matrix",
           "computational2pauli_basis_matrix"]


def vec(matrix: np.ndarray) -> np.ndarray:
    """Column-stacking vectorisation (reference :33-51); host-side index bookkeeping."""
    return np.asarray(matrix).T.reshape((-1, 1))


def unvec(vector: np.ndarray, shape: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """Inverse of vec (reference :54-79)."""
    vector = np.asarray(vector)
    if shape is None:
        dim = int(np.sqrt(vector.size))
        shape = dim, dim
    return vector.reshape(*shape).T


def pauli2computational_basis_matrix(dim) -> np.ndarray:
    """sum_k |sigma_k>> <k|: column k is vec(sigma_k), unnormalised Paulis in canonical order (reference :374-408).
    Host-side constant; the device kernels never materialise it (a Pauli is a pair of bit masks there)."""
    from ..synthetic import pauli_stack
    n = int(round(np.log2(dim)))
    if 2 ** n != dim or n < 1:
        raise ValueError("dim must be a power of two")
    ops = pauli_stack(n)
    return np.ascontiguousarray(ops.transpose(0, 2, 1).reshape(4 ** n, dim * dim).T.astype(complex))


def computational2pauli_basis_matrix(dim) -> np.ndarray:
    """(1/dim) sum_k |k> <<sigma_k|: the inverse of pauli2computational_basis_matrix (reference :411-438)."""
    return pauli2computational_basis_matrix(dim).conj().T / dim


def _check_c128(t, ndim):
    torch = _lib.require_cuda()
    if not t.is_cuda or t.dtype != torch.complex128 or t.dim() != ndim:
        raise ValueError(f"expected a complex128 CUDA tensor with {ndim} dimensions")
    return t.contiguous()


def _nq(d2):
    n = int(round(np.log2(d2) / 2))
    if 4 ** n != d2 or not 1 <= n <= 5:
        raise ValueError(f"superoperator dimension {d2} is not 4^n with 1 <= n <= 5")
    return n


def _kraus_call(name, kraus):
    torch = _lib.require_cuda()
    kraus = _check_c128(kraus, 4)
    b, nk, d, d_ = kraus.shape
    if d != d_:
        raise ValueError("batched Kraus conversion needs square Kraus operators")
    with _lib.on_device(kraus.device):
        out = torch.empty((b, d * d, d * d), dtype=torch.complex128, device=kraus.device)
        _lib.check(getattr(_lib.lib(), name)(ctypes.c_int(d), ctypes.c_int(nk), ctypes.c_int64(b), _lib.ptr(kraus),
                                             _lib.ptr(out), _lib.current_stream_ptr()), name)
    return out


def kraus2choi_batch(kraus):
    """kraus [B, K, d, d] -> choi [B, d^2, d^2]."""
    return _kraus_call("qt_kraus2choi_batch", kraus)


def kraus2superop_batch(kraus):
    """kraus [B, K, d, d] -> superop [B, d^2, d^2]."""
    return _kraus_call("qt_kraus2superop_batch", kraus)


def reshuffle_batch(mat, out=None):
    """choi2superop == superop2choi on a batch [B, d^2, d^2] (out-of-place)."""
    torch = _lib.require_cuda()
    mat = _check_c128(mat, 3)
    b, d2, _ = mat.shape
    d = 2 ** _nq(d2)
    with _lib.on_device(_lib.common_device(mat, out)):
        if out is None:
            out = torch.empty_like(mat)
        else:
            _lib.check_tensor("out", out, torch.complex128, mat.shape)
            if mat.numel() and out.data_ptr() == mat.data_ptr():
                raise ValueError("reshuffle_batch is out-of-place")
        _lib.check(_lib.lib().qt_choi_superop_reshuffle_batch(ctypes.c_int(d), ctypes.c_int64(b), _lib.ptr(mat),
                                                              _lib.ptr(out), _lib.current_stream_ptr()),
                   "qt_choi_superop_reshuffle_batch")
    return out


PL_VARIANTS = {"butterfly": 0, "dense_mma": 1}


def _pl_call(name, mat, out, workspace, variant="butterfly"):
    torch = _lib.require_cuda()
    mat = _check_c128(mat, 3)
    b, d2, _ = mat.shape
    n = _nq(d2)
    if variant != "butterfly":
        # the reference's dense formulation on the FP64 tensor path (n = 2, 3): kept for the measured comparison
        if variant not in PL_VARIANTS:
            raise ValueError(f"unknown variant {variant!r}; choose from {sorted(PL_VARIANTS)}")
        with _lib.on_device(_lib.common_device(mat, out)):
            if out is None:
                out = torch.empty_like(mat)
            else:
                _lib.check_tensor("out", out, torch.complex128, mat.shape)
            _lib.check(_lib.lib().qt_superop_pl_batch_variant(
                ctypes.c_int(n), ctypes.c_int64(b), _lib.ptr(mat), _lib.ptr(out), ctypes.c_void_p(0),
                ctypes.c_int(1 if name == "qt_superop2pl_batch" else 0), ctypes.c_int(PL_VARIANTS[variant]),
                _lib.current_stream_ptr()), "qt_superop_pl_batch_variant")
        return out
    with _lib.on_device(_lib.common_device(mat, out, workspace)):
        if out is None:
            out = torch.empty_like(mat)
        else:
            _lib.check_tensor("out", out, torch.complex128, mat.shape)
        if n >= 4 and workspace is None:
            workspace = torch.empty_like(mat)
        if workspace is not None and workspace.numel() * workspace.element_size() < mat.numel() * 16:
            raise ValueError("workspace is smaller than the batch")
        _lib.check(getattr(_lib.lib(), name)(ctypes.c_int(n), ctypes.c_int64(b), _lib.ptr(mat), _lib.ptr(out),
                                             _lib.ptr(workspace), _lib.current_stream_ptr()), name)
    return out


def superop2pauli_liouville_batch(superop, out=None, workspace=None, variant="butterfly"):
    return _pl_call("qt_superop2pl_batch", superop, out, workspace, variant)


def pauli_liouville2superop_batch(pl, out=None, workspace=None, variant="butterfly"):
    return _pl_call("qt_pl2superop_batch", pl, out, workspace, variant)


def choi2pauli_liouville_batch(choi):
    return superop2pauli_liouville_batch(reshuffle_batch(choi))


def pauli_liouville2choi_batch(pl):
    return reshuffle_batch(pauli_liouville2superop_batch(pl))


def choi2kraus_batch(choi, tol: float = 1e-9):
    """choi [B, d^2, d^2] (n <= 5) -> (kraus [B, d^2, d, d], counts [B] int32, evals [B, d^2] ascending).
    kraus[b, :counts[b]] are the operators sqrt(lambda) unvec(v) with |lambda| > tol in ascending-eigenvalue
    order (reference :325-336); the remaining slots are zero."""
    torch = _lib.require_cuda()
    choi = _check_c128(choi, 3)
    b, d2, _ = choi.shape
    n = _nq(d2)
    d = 2 ** n
    with _lib.on_device(choi.device):
        kraus = torch.empty((b, d2, d, d), dtype=torch.complex128, device=choi.device)
        evals = torch.empty((b, d2), dtype=torch.float64, device=choi.device)
        counts = torch.empty((b,), dtype=torch.int32, device=choi.device)
        if n >= 4:  # the eigenproblem does not fit shared memory: one-sided Jacobi out of an L2-resident workspace
            lib = _lib.lib()
            nbytes = int(lib.qt_choi2kraus_large_workspace_bytes(ctypes.c_int(n), ctypes.c_int64(b)))
            ws = torch.empty((nbytes // 16,), dtype=torch.complex128, device=choi.device)
            _lib.check(lib.qt_choi2kraus_large_batch(ctypes.c_int(n), ctypes.c_int64(b), _lib.ptr(choi), ctypes.c_double(tol),
                                                     _lib.ptr(evals), _lib.ptr(kraus), _lib.ptr(counts), _lib.ptr(ws),
                                                     ctypes.c_int64(nbytes), ctypes.c_void_p(0), _lib.current_stream_ptr()),
                       "qt_choi2kraus_large_batch")
            return kraus, counts, evals
        _lib.check(_lib.lib().qt_choi2kraus_batch(ctypes.c_int(n), ctypes.c_int64(b), _lib.ptr(choi), ctypes.c_double(tol),
                                                  _lib.ptr(evals), _lib.ptr(kraus), _lib.ptr(counts),
                                                  _lib.current_stream_ptr()), "qt_choi2kraus_batch")
        return kraus, counts, evals


def kraus2chi_batch(kraus):
    """kraus [B, K, d, d] -> chi [B, d^2, d^2] = c2p (sum_k vec K vec K^dagger) c2p^dagger (reference :82-97).
    c2p X c2p^dagger is the PTM butterfly kernel's (1/d) F X F^dagger divided by d."""
    d = kraus.shape[-1]
    return superop2pauli_liouville_batch(kraus2choi_batch(kraus)) / d


def chi2choi_batch(chi):
    """chi [B, d^2, d^2] -> choi = p2c chi p2c^dagger (reference :217-226) = d * (1/d) F^dagger chi F."""
    d = int(round(np.sqrt(chi.shape[-1])))
    return pauli_liouville2superop_batch(chi) * d


# ---- reference-named single-matrix functions -------------------------------------------------
def _to_dev(x):
    torch = _lib.require_cuda()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.complex128))).cuda()


def _kraus_stack(kraus_ops):
    if isinstance(kraus_ops, np.ndarray) and kraus_ops.ndim == 2:
        kraus_ops = [kraus_ops]
    return np.stack([np.asarray(k, dtype=np.complex128) for k in kraus_ops])


def kraus2choi(kraus_ops: Sequence[np.ndarray]) -> np.ndarray:
    """reference :159-182."""
    return kraus2choi_batch(_to_dev(_kraus_stack(kraus_ops)[None]))[0].cpu().numpy()


def kraus2superop(kraus_ops: Sequence[np.ndarray]) -> np.ndarray:
    """reference :100-145 (square Kraus operators; the reference also allows non-square ones)."""
    return kraus2superop_batch(_to_dev(_kraus_stack(kraus_ops)[None]))[0].cpu().numpy()


def kraus2pauli_liouville(kraus_ops: Sequence[np.ndarray]) -> np.ndarray:
    """reference :148-156."""
    s = kraus2superop_batch(_to_dev(_kraus_stack(kraus_ops)[None]))
    return superop2pauli_liouville_batch(s)[0].cpu().numpy()


def choi2superop(choi: np.ndarray) -> np.ndarray:
    """reference :351-361."""
    return reshuffle_batch(_to_dev(choi)[None])[0].cpu().numpy()


def superop2choi(superop: np.ndarray) -> np.ndarray:
    """reference :267-277."""
    return reshuffle_batch(_to_dev(superop)[None])[0].cpu().numpy()


def superop2pauli_liouville(superop: np.ndarray) -> np.ndarray:
    """reference :253-264."""
    return superop2pauli_liouville_batch(_to_dev(superop)[None])[0].cpu().numpy()


def pauli_liouville2superop(pl_matrix: np.ndarray) -> np.ndarray:
    """reference :301-312."""
    return pauli_liouville2superop_batch(_to_dev(pl_matrix)[None])[0].cpu().numpy()


def choi2kraus(choi: np.ndarray, tol: float = 1e-9) -> List[np.ndarray]:
    """reference :325-336 (ragged list: only operators whose |eigenvalue| exceeds tol)."""
    kraus, counts, _ = choi2kraus_batch(_to_dev(choi)[None], tol)
    k = int(counts[0].item())
    return [m for m in kraus[0, :k].cpu().numpy()]


def superop2kraus(superop: np.ndarray) -> List[np.ndarray]:
    """reference :229-238."""
    return choi2kraus(superop2choi(superop))


def pauli_liouville2kraus(pl_matrix: np.ndarray) -> List[np.ndarray]:
    """reference :280-288."""
    return choi2kraus(pauli_liouville2choi(pl_matrix))


def kraus2chi(kraus_ops: Sequence[np.ndarray]) -> np.ndarray:
    """reference :82-97."""
    return kraus2chi_batch(_to_dev(_kraus_stack(kraus_ops)[None]))[0].cpu().numpy()


def chi2choi(chi_matrix: np.ndarray) -> np.ndarray:
    """reference :217-226."""
    return chi2choi_batch(_to_dev(chi_matrix)[None])[0].cpu().numpy()


def chi2pauli_liouville(chi_matrix: np.ndarray) -> np.ndarray:
    """reference :185-192."""
    return choi2pauli_liouville_batch(chi2choi_batch(_to_dev(chi_matrix)[None]))[0].cpu().numpy()


def chi2superop(chi_matrix: np.ndarray) -> np.ndarray:
    """reference :207-214."""
    return reshuffle_batch(chi2choi_batch(_to_dev(chi_matrix)[None]))[0].cpu().numpy()


def chi2kraus(chi_matrix: np.ndarray) -> List[np.ndarray]:
    """reference :195-204."""
    return choi2kraus(chi2choi(chi_matrix))


def choi2chi(choi: np.ndarray) -> np.ndarray:
    """reference :339-348: kraus2chi(choi2kraus(choi)) -- eigenvalues below 1e-9 in magnitude are dropped and
    negative ones enter with their absolute value, exactly like the reference's round trip through Kraus form."""
    return kraus2chi(choi2kraus(choi))


def superop2chi(superop: np.ndarray) -> np.ndarray:
    """reference :241-250."""
    return choi2chi(superop2choi(superop))


def pauli_liouville2chi(pl_matrix: np.ndarray) -> np.ndarray:
    """reference :291-298."""
    return choi2chi(pauli_liouville2choi(pl_matrix))


def choi2pauli_liouville(choi: np.ndarray) -> np.ndarray:
    """reference :364-371."""
    return choi2pauli_liouville_batch(_to_dev(choi)[None])[0].cpu().numpy()


def pauli_liouville2choi(pl_matrix: np.ndarray) -> np.ndarray:
    """reference :315-322."""
    return pauli_liouville2choi_batch(_to_dev(pl_matrix)[None])[0].cpu().numpy()
