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


def _shard_model_worker(rank, world, port, q):
    """The dataflow of csrc/superb_mgpu.cu as a float64 model, one hop per rank, exchanges over gloo:
    phase 1 local spectrum + alignment ramp + re-order, all-to-all #1 (rank t collects the decimated sub-sequence t of the
    concatenated spectrum), N-point inverse per rank, all-to-all #2 + the last log2 H radix-2 stages, |.| gathered on the root --
    against superb_ondataready's own order of operations (rotate, FFT/N, concatenate, one H*N-point inverse; superbandwidth.c:121-152)."""
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, N = world, 64
    rng = np.random.default_rng(7)
    hops = rng.standard_normal((H, N)) + 1j * rng.standard_normal((H, N))          # the same on every rank; rank q only USES hop q
    lags = [0, 5, 17, 40][:H]
    # ---- what the reference computes (every rank can check against it)
    cat = np.concatenate([np.fft.fft(np.roll(hops[qq], -lags[qq])) / N for qq in range(H)])
    want = np.abs(np.fft.ifft(cat) * (H * N))                                       # fft_perform's inverse does not scale
    # ---- phase 1 (local): X_q, ramp e^{2 pi i m lag / N} == rotation by lag samples, re-order so that run t goes to rank t
    m = np.arange(N)
    Xq = np.fft.fft(hops[rank]) / N * np.exp(2j * np.pi * m * lags[rank] / N)
    per = N // H
    Xp = np.stack([Xq[t::H] for t in range(H)])                                     # Xp[t][u'] = X~_q[t + H u']
    # ---- all-to-all #1: rank t receives run t of every rank q -> B_t[u], u = q * per + u'  (k = H u + t with k = q N + m ... m = t + H u')
    def all_to_all(rows):                              # gloo has no all-to-all: everybody gathers everything and keeps its column
        mine_t = torch.from_numpy(np.ascontiguousarray(np.stack([np.stack([r.real, r.imag]) for r in rows])))      # [H][2][len]
        everyone = [torch.empty_like(mine_t) for _ in range(H)]
        dist.all_gather(everyone, mine_t)
        return [everyone[src][rank].numpy()[0] + 1j * everyone[src][rank].numpy()[1] for src in range(H)]
    B = np.concatenate(all_to_all([Xp[t] for t in range(H)]))                         # length N: sub-sequence t = rank of the big spectrum
    A = np.fft.ifft(B) * N                                                           # unscaled N-point inverse
    # ---- all-to-all #2: rank r owns positions v in [r per, (r+1) per) and needs A_t[v] from every t
    At = np.stack(all_to_all([A[r * per:(r + 1) * per] for r in range(H)]))         # At[t][j], v = rank * per + j
    v = rank * per + np.arange(per)
    # the last log2 H decimation-in-time stages across the H blocks (block b holds A_{bitrev(b)})
    log2h = H.bit_length() - 1
    rev = lambda b: int(format(b, f"0{log2h}b")[::-1], 2) if log2h else 0
    a = [At[rev(b)].copy() for b in range(H)]
    for s in range(log2h):
        for b in range(H):
            if b & (1 << s):
                continue
            beta = b & ((1 << s) - 1)
            w = np.exp(1j * np.pi * (v + N * beta) / (N * (1 << s)))
            lo, hi = a[b], a[b | (1 << s)] * w
            a[b], a[b | (1 << s)] = lo + hi, lo - hi
    mine = np.stack([np.abs(a[c]) for c in range(H)])                                # mine[c][j] = |y[v + N c]|
    # ---- the root collects the time-contiguous stream
    parts = [torch.empty(H, per, dtype=torch.float64) for _ in range(H)] if rank == 0 else None
    dist.gather(torch.from_numpy(mine), parts, dst=0)
    ok = True
    if rank == 0:
        stream = np.zeros(H * N)
        for r in range(H):
            for c in range(H):
                stream[N * c + r * per: N * c + (r + 1) * per] = parts[r][c].numpy()
        ok = bool(np.max(np.abs(stream - want)) <= 1e-9 * np.max(want))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_sharded_stitch_model_matches_the_reference_order(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_model_worker, args=(r, world, 29750 + world, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]
