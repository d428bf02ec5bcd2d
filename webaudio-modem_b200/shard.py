"""Multi-GPU sharding of independent streams: one process per GPU, contiguous stream ranges,
no collective on the data path.  torch.distributed is used only for the
barrier around timing and for gathering the small results (decoded bytes, counters)."""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of `n_items` owned by `rank` (sizes differ by at most one)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_stream_results(local_bytes: list[bytes], n_streams: int, rank: int, world: int, dist=None) -> list[bytes] | None:
    """Collect per-stream decoded bytes on rank 0 in global stream order (None on other ranks).
    `dist` is torch.distributed (already initialised: nccl on GPUs, gloo in the CPU tests)."""
    if world == 1 or dist is None:
        return list(local_bytes)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local_bytes, gathered, dst=0)
    if rank != 0:
        return None
    out: list[bytes] = []
    for r in range(world):
        lo, hi = shard_range(n_streams, r, world)
        assert len(gathered[r]) == hi - lo
        out.extend(gathered[r])
    return out


def reduce_counters(values: np.ndarray, dist=None, device=None) -> np.ndarray:
    """Sum small integer counters over all ranks (reporting only, off the data path)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(values).copy()
    import torch

    t = torch.as_tensor(np.asarray(values, dtype=np.int64), device=device)
    dist.all_reduce(t)
    return t.cpu().numpy()
