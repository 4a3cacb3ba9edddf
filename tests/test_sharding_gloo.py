"""N>1 host logic on CPU: world_size-2 (and 3) gloo runs of the batch sharding + single all-gather
(SURVEY.md 8e).  The compute is a stand-in (the kernels need a GPU); what is checked is the partition and
the collective plumbing bench.py / users rely on."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from forest_benchmarking_b200.sharding import all_gather_states, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 1024, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        full = rng.standard_normal((total, 4, 4)) + 1j * rng.standard_normal((total, 4, 4))
        lo, hi = shard_range(total, world, rank)
        local = torch.from_numpy(full[lo:hi] * 2.0)  # stand-in for "reconstruct my slice"
        got = all_gather_states(local, total)
        iters = all_gather_states(torch.arange(lo, hi, dtype=torch.int32), total)
        ok = np.array_equal(got.numpy(), full * 2.0) and np.array_equal(iters.numpy(), np.arange(total))
        q.put((rank, bool(ok), tuple(got.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 64), (2, 37), (3, 10)])
def test_all_gather_states_gloo(world, total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (total, 4, 4) for _, _, shape in res)


def test_single_process_is_identity():
    t = torch.ones(3, 2, 2, dtype=torch.complex128)
    assert all_gather_states(t, 3) is t
