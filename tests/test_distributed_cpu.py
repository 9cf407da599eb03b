"""world_size-2 gloo tests (CPU) of the host-side logic of the multi-GPU paths: the single all-gather's layout and
the rank -> residue mapping of the superbandwidth stitch, and the replica sharding arithmetic of bench.py."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tempestsdr_b200 import superband
    n, nd = 8, 4
    block = torch.arange(2 * (n + nd), dtype=torch.float32) + 1000.0 * rank       # rank-tagged [X | D]
    g = superband.gather_blocks(block)
    ok = g.numel() == world * block.numel()
    for r in range(world):
        ok = ok and torch.equal(g[r * block.numel():(r + 1) * block.numel()], torch.arange(2 * (n + nd), dtype=torch.float32) + 1000.0 * r)
    # every rank owns a distinct residue and together they tile the output
    res = torch.zeros(world, dtype=torch.int64); res[superband.residue_of_rank(rank, world)] = 1
    dist.all_reduce(res)
    ok = ok and bool((res == 1).all())
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and t.item() == float(world)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gloo_world2_gather_layout_and_residues():
    world, port = 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]
