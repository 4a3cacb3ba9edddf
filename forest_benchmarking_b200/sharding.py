"""Batch sharding across the GPUs of one box.

Every experiment of a batch is independent, so the only multi-GPU structure on this path is: rank r
reconstructs a contiguous slice of the batch, then ONE all-gather leaves every rank with all the results.
One process per GPU (torchrun); backend "nccl" on GPUs ("gloo" in the CPU tests of this host logic).
"""
from typing import Tuple


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the contiguous slice of `total` items owned by `rank`; the first total % world ranks
    take one extra item, so shard sizes differ by at most one and concatenate in rank order."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of world {world}")
    base, extra = divmod(int(total), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_states(local, total: int, group=None):
    """All-gather per-rank result slices (first axis = batch slice given by shard_range) into the full
    [total, ...] tensor on every rank.  complex128 travels as float64 pairs.  Single collective."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(total, world, rank)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} items, expected {hi - lo}")
    is_c = local.is_complex()
    flat = torch.view_as_real(local.contiguous()) if is_c else local.contiguous()
    per = -(-total // world)  # padded shard size: one collective even when total % world != 0
    item_shape = tuple(flat.shape[1:])
    send = flat
    if hi - lo != per:
        send = torch.zeros((per,) + item_shape, dtype=flat.dtype, device=flat.device)
        send[: hi - lo] = flat
    recv = torch.empty((world * per,) + item_shape, dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    if total % world:
        parts = []
        for r in range(world):
            a, b = shard_range(total, world, r)
            parts.append(recv[r * per: r * per + (b - a)])
        recv = torch.cat(parts, dim=0)
    return torch.view_as_complex(recv) if is_c else recv


class SingleProcessComm:
    """One host process driving several GPUs (no torch.distributed): NCCL communicators created through the C ABI
    (``qt_comm_init_all``), and the path's single collective, an all-gather of per-device result slices
    (``qt_allgather_bytes``).  ``devices``: CUDA device indices, one rank each, in rank order."""

    def __init__(self, devices):
        import ctypes
        from . import _lib
        self._lib, self._ct = _lib, ctypes
        self.devices = [int(d) for d in devices]
        arr = (ctypes.c_int32 * len(self.devices))(*self.devices)
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().qt_comm_init_all(len(self.devices), arr, ctypes.byref(self._h)), "qt_comm_init_all")

    def all_gather(self, slices):
        """slices[r]: contiguous tensor on cuda:devices[r], identical shape / dtype on every rank.  Returns one tensor per
        rank holding all slices concatenated along the first axis in rank order."""
        import torch
        ct = self._ct
        n = len(self.devices)
        if len(slices) != n:
            raise ValueError(f"expected {n} slices, one per device")
        shape, dtype = tuple(slices[0].shape), slices[0].dtype
        for r, t in enumerate(slices):
            if tuple(t.shape) != shape or t.dtype != dtype or not t.is_contiguous() or t.device.index != self.devices[r]:
                raise ValueError(f"slice {r} must be a contiguous {dtype} tensor of shape {list(shape)} on cuda:{self.devices[r]}")
        outs = [torch.empty((n * shape[0],) + shape[1:], dtype=dtype, device=t.device) for t in slices]
        nbytes = slices[0].numel() * slices[0].element_size()
        send = (ct.c_void_p * n)(*[t.data_ptr() for t in slices])
        recv = (ct.c_void_p * n)(*[t.data_ptr() for t in outs])
        streams = (ct.c_void_p * n)(*[torch.cuda.current_stream(t.device).cuda_stream for t in slices])
        self._lib.check(self._lib.lib().qt_allgather_bytes(self._h, send, recv, ct.c_int64(nbytes), streams),
                        "qt_allgather_bytes")
        return outs

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.lib().qt_comm_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
