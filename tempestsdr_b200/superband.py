"""Superbandwidth stitch with one hop per GPU.

``stitch_distributed`` is what rank r runs: local spectra -> ONE NCCL all-gather (``torch.distributed``) -> integer
alignment lags (exact) -> this rank's strided residue of the H*N-point inverse.  ``stitch_simulated`` replays the same
dataflow for all ranks inside one process (no process group) -- used by the single-GPU parity test.

Replaces superb_ondataready (superbandwidth.c:121-152) for the multi-GPU configuration; the single-GPU form is
``Context.superb_stitch``.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import api


def local_block(ctx: api.Context, hop: torch.Tensor, samples_in_frame: int) -> Tuple[torch.Tensor, int, int]:
    """[FFT_N(raw hop) | FFT_nd(first difference of |hop|)] as interleaved float32, plus (N, nd)."""
    pairs = hop.numel() // 2
    n = ctx.fft_getrealsize(pairs)
    block = torch.empty(2 * (2 * n), dtype=torch.float32, device=hop.device)       # room for N + nd <= 2N complex
    hn, hnd = C.c_uint32(0), C.c_uint32(0)
    ctx.chk(ctx._lib.tsdrgpu_superb_local_spectra(ctx._h, ctx.stream, hop.data_ptr(), pairs, samples_in_frame, block.data_ptr(),
                                                  C.byref(hn), C.byref(hnd)))
    return block[: 2 * (hn.value + hnd.value)], hn.value, hnd.value


def lags_from_gathered(ctx: api.Context, gathered: torch.Tensor, nhops: int, n: int, nd: int) -> List[int]:
    lags = (C.c_int * nhops)()
    ctx.chk(ctx._lib.tsdrgpu_superb_lags(ctx._h, ctx.stream, gathered.data_ptr(), nhops, n + nd, n, nd, lags))
    return list(lags)


def residue(ctx: api.Context, gathered: torch.Tensor, nhops: int, n: int, nd: int, s: int, lags: Sequence[int]) -> torch.Tensor:
    out = torch.empty(2 * n, dtype=torch.float32, device=gathered.device)
    arr = (C.c_int * nhops)(*lags)
    ctx.chk(ctx._lib.tsdrgpu_superb_residue_ifft_lag(ctx._h, ctx.stream, gathered.data_ptr(), nhops, n + nd, n, s, arr, out.data_ptr()))
    return out


def gather_blocks(block: torch.Tensor, group=None) -> torch.Tensor:
    """The single collective of the path: all-gather of every rank's block, rank-major (works on NCCL and gloo)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty(world * block.numel(), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(out, block.contiguous(), group=group)
    return out


def residue_of_rank(rank: int, world: int) -> int:
    """Rank r produces output samples y[world*p + r]."""
    return rank % world


def stitch_distributed(ctx: api.Context, hop: torch.Tensor, samples_in_frame: int, group=None):
    """Rank-local view of the stitch: returns (this rank's residue as interleaved IQ, lags in pairs, N)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    block, n, nd = local_block(ctx, hop, samples_in_frame)
    gathered = gather_blocks(block, group)
    lags = lags_from_gathered(ctx, gathered, world, n, nd)
    return residue(ctx, gathered, world, n, nd, residue_of_rank(rank, world), lags), lags, n


def stitch_simulated(ctx: api.Context, hops: Sequence[torch.Tensor], samples_in_frame: int):
    """All ranks in one process: returns (full stitched IQ interleaved, lags in FLOATS like the reference's best_offset)."""
    blocks = [local_block(ctx, h, samples_in_frame) for h in hops]
    n, nd = blocks[0][1], blocks[0][2]
    gathered = torch.cat([b[0] for b in blocks])
    H = len(hops)
    lags = lags_from_gathered(ctx, gathered, H, n, nd)
    full = torch.empty(H * n, 2, dtype=torch.float32, device=hops[0].device)
    for s in range(H):
        full[s::H] = residue(ctx, gathered, H, n, nd, s, lags).view(n, 2)
    return full.reshape(-1), [2 * l for l in lags]
