"""CPU (gloo, world_size 2): the N>1 host logic -- contiguous frame sharding + one all-gather of the
probability blocks reproduces the single-process matrix, including uneven and empty shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from timed_design_b200.dist import gather_rows, shard_range, shard_samples


def _fake_probs(idx: np.ndarray, c: int) -> np.ndarray:
    z = np.sin(np.outer(idx + 1, np.arange(1, c + 1)) * 0.37)
    e = np.exp(z)
    return (e / e.sum(1, keepdims=True)).astype(np.float32)


def _worker(rank, world, port, n_total, c, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, stop = shard_range(n_total, rank, world)
        local = torch.from_numpy(_fake_probs(np.arange(start, stop), c)) if stop > start else torch.zeros((0, c))
        full = gather_rows(local, n_total)
        q.put((rank, full.numpy()))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n_total,c", [(76, 20), (7, 338), (1, 20)])
def test_gather_rows_world2_matches_single_process(n_total, c):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, c, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _fake_probs(np.arange(n_total), c)
    for r in range(2):
        np.testing.assert_array_equal(results[r], ref)


def test_shard_ranges_partition_everything():
    for n in (0, 1, 7, 76, 4096, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(0 <= s <= e for s, e in spans)
            assert max(e - s for s, e in spans) == (-(-n // world) if n else 0)
    assert shard_samples(10, 1, 4) == (3, 3) and shard_samples(10, 3, 4) == (9, 1)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_gather_rows_without_process_group_is_identity():
    x = torch.arange(12, dtype=torch.float32).reshape(6, 2)
    assert torch.equal(gather_rows(x, 6), x)
