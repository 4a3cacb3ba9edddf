"""Projections of Choi matrices onto CP / TP / TNI / physical maps -- signatures of
forest/benchmarking/operator_tools/project_superoperators.py, computed by csrc/qt_project.cu
(batched Jacobi eigensolver + fused Dykstra loop, one matrix per warp or block)."""
import ctypes

import numpy as np

from .. import _lib

__all__ = ["proj_choi_to_completely_positive", "proj_choi_to_trace_preserving",
           "proj_choi_to_trace_non_increasing", "proj_choi_to_physical",
           "proj_choi_to_completely_positive_batch", "proj_choi_to_trace_preserving_batch",
           "proj_choi_to_trace_non_increasing_batch", "proj_choi_to_physical_batch"]


def _prep(choi, max_n=5):
    torch = _lib.require_cuda()
    if not choi.is_cuda or choi.dtype != torch.complex128 or choi.dim() != 3 or choi.shape[1] != choi.shape[2]:
        raise ValueError("expected a complex128 CUDA tensor [B, 4^n, 4^n]")
    n = int(round(np.log2(choi.shape[1]) / 2))
    if 4 ** n != choi.shape[1] or not 1 <= n <= max_n:
        raise ValueError(f"this Choi operation supports n = 1..{max_n} qubits")
    return torch, choi.contiguous(), n


def _simple(name, choi, out, max_n=5):
    torch, choi, n = _prep(choi, max_n)
    dev = _lib.common_device(choi, out)
    with _lib.on_device(dev):
        if out is None:
            out = torch.empty_like(choi)
        else:
            _lib.check_tensor("out", out, torch.complex128, choi.shape)
        if n >= 4 and choi.numel() and out.data_ptr() == choi.data_ptr():
            raise ValueError("projections of 4- and 5-qubit Choi matrices are out-of-place")
        _lib.check(getattr(_lib.lib(), name)(ctypes.c_int(n), ctypes.c_int64(choi.shape[0]), _lib.ptr(choi),
                                             _lib.ptr(out), _lib.current_stream_ptr()), name)
    return out


def proj_choi_to_completely_positive_batch(choi, out=None):
    """(C + C^dagger)/2 -> eigh -> clamp -> V L V^dagger for every matrix of the batch, n = 1..5 (n >= 4: the eigenproblem
    runs out of an L2-resident workspace, one matrix per SM)."""
    torch, choi, n = _prep(choi)
    if n <= 3:
        return _simple("qt_proj_cp_batch", choi, out)
    lib = _lib.lib()
    dev = _lib.common_device(choi, out)
    with _lib.on_device(dev):
        if out is None:
            out = torch.empty_like(choi)
        else:
            _lib.check_tensor("out", out, torch.complex128, choi.shape)
            if choi.numel() and out.data_ptr() == choi.data_ptr():
                raise ValueError("projections of 4- and 5-qubit Choi matrices are out-of-place")
        nbytes = int(lib.qt_proj_cp_workspace_bytes(ctypes.c_int(n), ctypes.c_int64(choi.shape[0])))
        ws = torch.empty((max(nbytes, 16) // 16,), dtype=torch.complex128, device=dev)
        _lib.check(lib.qt_proj_cp_ws_batch(ctypes.c_int(n), ctypes.c_int64(choi.shape[0]), _lib.ptr(choi), _lib.ptr(out),
                                           _lib.ptr(ws), ctypes.c_int64(nbytes), _lib.current_stream_ptr()),
                   "qt_proj_cp_ws_batch")
    return out


def proj_choi_to_trace_preserving_batch(choi, out=None):
    return _simple("qt_proj_tp_batch", choi, out)


def proj_choi_to_trace_non_increasing_batch(choi, out=None):
    return _simple("qt_proj_tni_batch", choi, out)


def proj_choi_to_unitary_batch(choi, out=None):
    return _simple("qt_proj_unitary_batch", choi, out, max_n=3)


def proj_choi_to_physical_batch(choi, make_trace_preserving=True, out=None, return_counts=False,
                                eigh_rel_tol=None, return_status=False):
    """Dykstra CP + TP (or TNI) projection of every matrix of the batch (reference project_superoperators.py:87-144).

    A non-Hermitian input is treated exactly like the reference treats it: its CP step Hermitises (:30), so the
    iterates and the result only depend on the Hermitian part, while the anti-Hermitian part stays inside
    ``old_CP_change`` and enters the stopping rule (:132-136) -- the kernel adds those terms, so the number of
    Dykstra trips (``return_counts``) matches the reference for such inputs too.
    ``eigh_rel_tol``: eigensolver stopping tolerance (None = library default 1e-8, 0 = tight), per call.
    ``return_status``: also return the int32 [B] status words (non-zero where a safety cap was hit)."""
    torch, choi, n = _prep(choi)
    b = choi.shape[0]
    lib = _lib.lib()
    dev = _lib.common_device(choi, out)
    with _lib.on_device(dev):
        if out is None:
            out = torch.empty_like(choi)
        else:
            _lib.check_tensor("out", out, torch.complex128, choi.shape)
            if choi.numel() and out.data_ptr() == choi.data_ptr():
                raise ValueError("proj_choi_to_physical_batch cannot run in place")
        nbytes = int(lib.qt_proj_physical_workspace_bytes(ctypes.c_int(n), ctypes.c_int64(b)))
        ws = torch.empty((max(nbytes, 16) // 16,), dtype=torch.complex128, device=dev)
        counts = torch.empty((b,), dtype=torch.int32, device=dev)
        status = torch.zeros((b,), dtype=torch.int32, device=dev)
        _lib.check(lib.qt_proj_physical_batch(ctypes.c_int(n), ctypes.c_int64(b), _lib.ptr(choi), _lib.ptr(out),
                                              ctypes.c_int(1 if make_trace_preserving else 0),
                                              ctypes.c_double(-1.0 if eigh_rel_tol is None else eigh_rel_tol),
                                              _lib.ptr(ws), ctypes.c_int64(nbytes), _lib.ptr(counts), _lib.ptr(status),
                                              _lib.current_stream_ptr()), "qt_proj_physical_batch")
    res = (out,)
    if return_counts:
        res += (counts,)
    if return_status:
        res += (status,)
    return res if len(res) > 1 else out


def _one(x):
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.complex128))[None]).to(dev)


def proj_choi_to_completely_positive(choi: np.ndarray, check_finite: bool = True) -> np.ndarray:
    """reference project_superoperators.py:19-34."""
    choi = np.asarray(choi)
    if check_finite and not np.isfinite(choi).all():
        raise ValueError("array must not contain infs or NaNs")
    return proj_choi_to_completely_positive_batch(_one(choi))[0].cpu().numpy()


def proj_choi_to_trace_non_increasing(choi: np.ndarray) -> np.ndarray:
    """reference project_superoperators.py:37-59."""
    return proj_choi_to_trace_non_increasing_batch(_one(choi))[0].cpu().numpy()


def proj_choi_to_trace_preserving(choi: np.ndarray) -> np.ndarray:
    """reference project_superoperators.py:62-84."""
    return proj_choi_to_trace_preserving_batch(_one(choi))[0].cpu().numpy()


def proj_choi_to_physical(choi: np.ndarray, make_trace_preserving: bool = True) -> np.ndarray:
    """reference project_superoperators.py:87-144."""
    from ..tomography import warn_on_status
    out, status = proj_choi_to_physical_batch(_one(choi), make_trace_preserving, return_status=True)
    warn_on_status(status.cpu().numpy(), "proj_choi_to_physical")
    return out[0].cpu().numpy()


def proj_choi_to_unitary(choi: np.ndarray, check_finite: bool = True) -> np.ndarray:
    """reference project_superoperators.py:147-175."""
    choi = np.asarray(choi)
    if check_finite and not np.isfinite(choi).all():
        raise ValueError("array must not contain infs or NaNs")
    return proj_choi_to_unitary_batch(_one(choi))[0].cpu().numpy()
