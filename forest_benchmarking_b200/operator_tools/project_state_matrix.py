"""project_state_matrix_to_physical -- signature of
forest/benchmarking/operator_tools/project_state_matrix.py:6-52, computed by csrc/qt_distance.cu
(one state per warp: Jacobi eigendecomposition + water filling)."""
import ctypes

import numpy as np

from .. import _lib

__all__ = ["project_state_matrix_to_physical", "project_state_matrix_to_physical_batch"]


def project_state_matrix_to_physical_batch(rho, out=None):
    """rho [B, d, d] complex128 CUDA -> closest trace-one PSD matrices [B, d, d]."""
    torch = _lib.require_cuda()
    if not rho.is_cuda or rho.dtype != torch.complex128 or rho.dim() != 3 or rho.shape[1] != rho.shape[2]:
        raise ValueError("expected a complex128 CUDA tensor [B, d, d]")
    rho = rho.contiguous()
    d = rho.shape[1]
    n = int(round(np.log2(d)))
    if 2 ** n != d or not 1 <= n <= 5:
        raise ValueError(f"dimension {d} is not 2^n with 1 <= n <= 5")
    with _lib.on_device(_lib.common_device(rho, out)):
        if out is None:
            out = torch.empty_like(rho)
        else:
            _lib.check_tensor("out", out, torch.complex128, rho.shape)
        _lib.check(_lib.lib().qt_project_state_batch(ctypes.c_int(n), ctypes.c_int64(rho.shape[0]), _lib.ptr(rho),
                                                     _lib.ptr(out), _lib.current_stream_ptr()),
                   "qt_project_state_batch")
    return out


def project_state_matrix_to_physical(rho: np.ndarray) -> np.ndarray:
    """Drop-in for reference project_state_matrix.py:6-52."""
    torch = _lib.require_cuda()
    x = torch.from_numpy(np.ascontiguousarray(np.asarray(rho, dtype=np.complex128))[None]).cuda()
    return project_state_matrix_to_physical_batch(x)[0].cpu().numpy()
