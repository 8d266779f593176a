"""Multi-GPU plumbing for the hot path: frames shard embarrassingly by flat frame
index -- contiguous ranges, so each chain's rows stay adjacent as ``extract_sequence_from_pred_matrix``
(design_utils/utils.py:679-692) requires -- with ONE all-gather of the per-rank probability blocks.
``torch.distributed`` is used for the plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of this rank's contiguous share of ``n_total`` rows: ceil(N/G) rows per rank,
    the last ranks possibly short or empty."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    per = -(-n_total // world) if n_total > 0 else 0
    start = min(rank * per, n_total)
    return start, min(start + per, n_total)


def shard_samples(n_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """(first_sample, count) of a rank's block of a chain's ``n_samples`` draws; passing
    ``first_sample`` to ``sample_block`` makes the union of the shards identical to one device's draw."""
    start, stop = shard_range(n_samples, rank, world)
    return start, stop - start


def gather_rows(local, n_total: int, group=None):
    """All-gather row blocks produced under ``shard_range`` into the full (n_total, C) matrix on every
    rank.  Uneven shards are padded to ceil(N/G) rows for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return local[:n_total]
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-n_total // world)
    start, stop = shard_range(n_total, rank, world)
    if local.shape[0] != stop - start:
        raise ValueError(f"rank {rank}: expected {stop - start} local rows, got {local.shape[0]}")
    padded = local
    if local.shape[0] != per:
        padded = torch.zeros((per, *local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    out = torch.empty((world * per, *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    return out[:n_total]
