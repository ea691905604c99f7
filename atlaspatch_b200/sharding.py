"""Slide / patch sharding for one-process-per-GPU runs.

The reference scales out by launching one OS process per GPU and letting lock files arbitrate
(README.md:527,628; orchestration/runner.py:154-168).  Here ranks are `torch.distributed` processes and the
assignment is computed, not raced for:

* inter-slide (default): longest-processing-time-first by level-0 pixel count -> every rank gets a list of slides; no
  data-path collective at all (the only collective is the weight broadcast at start-up);
* intra-slide: contiguous row ranges of the (N, 5) coordinate array, so concatenating the per-rank feature blocks in rank
  order reproduces the reference's row order.
"""
from __future__ import annotations

from typing import Sequence


def assign_slides(sizes: Sequence[int], world_size: int) -> list[list[int]]:
    """LPT: slide indices per rank; deterministic (ties broken by index)."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    loads = [0] * world_size
    out: list[list[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += int(sizes[i])
    for lst in out:
        lst.sort()
    return out


def row_range(n_rows: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous [begin, end) slice of the coordinate rows owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_rows), world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_rows(local, n_rows: int, *, group=None):
    """Intra-slide mode (BASELINE.json configs[4]): every rank holds the feature rows of its `row_range`; one all_gather
    (NCCL over NVLink for CUDA tensors, gloo for CPU tensors) returns the full (n_rows, D) matrix on every rank, in the reference's
    row order.  This is the path's only data-path collective: (N, D) fp32 once per slide (100 k x 1536 x 4 B = 614 MB)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    b, e = row_range(n_rows, rank, world)
    if local.shape[0] != e - b:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, its range [{b}, {e}) has {e - b}")
    pad = row_range(n_rows, 0, world)[1]                          # rank 0 owns the longest range
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: e - b] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = []
    for r, p in enumerate(parts):
        rb, re_ = row_range(n_rows, r, world)
        out.append(p[: re_ - rb])
    return torch.cat(out)
