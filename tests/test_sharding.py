"""CPU tests of the multi-GPU host logic, including a world_size-2 gloo run (no GPU needed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from atlaspatch_b200.sharding import assign_slides, gather_rows, row_range


def test_assign_slides_lpt_covers_everything_once():
    sizes = [80000 * 60000, 40000 * 40000, 40000 * 40000, 8192 * 8192, 100, 40000 * 40000, 7, 7]
    for ws in (1, 2, 3, 8, 16):
        parts = assign_slides(sizes, ws)
        assert len(parts) == ws
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in p) for p in parts]
        assert max(loads) <= sum(sizes) / ws + max(sizes)          # LPT bound
    two = assign_slides(sizes, 2)
    assert two[0][0] == 0 and set(two[1]) >= {1, 2, 5}            # the big slide alone vs the three 40k^2 slides
    assert assign_slides([], 4) == [[], [], [], []]


def test_row_ranges_partition_in_order():
    for n in (0, 1, 7, 25396, 100000):
        for ws in (1, 2, 4, 8):
            rs = [row_range(n, r, ws) for r in range(ws)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(ws - 1))
            lens = [e - b for b, e in rs]
            assert max(lens) - min(lens) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rows, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 0 owns the "weights"; every rank must end up with identical bytes (stands in for the NCCL broadcast)
    w = torch.arange(1000, dtype=torch.float32) * (1.0 if rank == 0 else 0.0)
    dist.broadcast(w, src=0)
    assert torch.equal(w, torch.arange(1000, dtype=torch.float32))
    # intra-slide mode: each rank "embeds" its contiguous row range; gather reproduces the reference row order
    coords = torch.arange(n_rows * 5, dtype=torch.int32).reshape(n_rows, 5)
    b, e = row_range(n_rows, rank, world)
    feats = coords[b:e, :1].to(torch.float32) * 2.0 + 1.0          # a function of the row only
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([e - b], dtype=torch.int64))
    total = int(sum(c.item() for c in counts))
    assert total == n_rows
    pad = max(int(c.item()) for c in counts)
    buf = torch.zeros(pad, 1)
    buf[: e - b] = feats
    parts = [torch.zeros(pad, 1) for _ in range(world)]
    dist.all_gather(parts, buf)
    full = torch.cat([p[: int(c.item())] for p, c in zip(parts, counts)])
    assert torch.equal(full, coords[:, :1].to(torch.float32) * 2.0 + 1.0)
    # the product function for the same exchange, (N, D) features
    f2 = torch.stack([coords[b:e, 0].float(), coords[b:e, 1].float() * 0.5, torch.full((e - b,), float(rank))], 1)
    g2 = gather_rows(f2, n_rows)
    assert g2.shape == (n_rows, 3) and torch.equal(g2[:, 0], coords[:, 0].float()) and torch.equal(g2[:, 1], coords[:, 1].float() * 0.5)
    owners = torch.cat([torch.full((row_range(n_rows, r, world)[1] - row_range(n_rows, r, world)[0],), float(r)) for r in range(world)])
    assert torch.equal(g2[:, 2], owners)
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([b, e]))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(tmp_path):
    world, n_rows = 2, 1001
    mp.spawn(_worker, args=(world, _free_port(), n_rows, str(tmp_path)), nprocs=world, join=True)
    rs = [tuple(np.load(tmp_path / f"ok_{r}.npy")) for r in range(world)]
    assert rs == [(0, 501), (501, 1001)]
