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


def own_lag_from_gathered(ctx: api.Context, gathered: torch.Tensor, rank: int, n: int, nd: int) -> int:
    """The alignment lag of hop `rank` against hop 0 only: tsdrgpu_superb_lags on the two-block view {block 0, block rank}."""
    if rank == 0:
        return 0
    lags = (C.c_int * 2)()
    ctx.chk(ctx._lib.tsdrgpu_superb_lags(ctx._h, ctx.stream, gathered.data_ptr(), 2, rank * (n + nd), n, nd, lags))
    return int(lags[1])


def stitch_distributed(ctx: api.Context, hop: torch.Tensor, samples_in_frame: int, group=None, distributed_lags: bool = False):
    """Rank-local view of the stitch: returns (this rank's residue as interleaved IQ, lags in pairs, N).

    distributed_lags=False: every rank derives all H-1 lags itself from the gathered difference spectra (one collective in
    total, redundant cross-correlations).  True: rank q computes only ITS lag and the H integers are exchanged with a second,
    4-byte-per-rank all-gather -- the lag work per rank no longer grows with H (not yet measured: DESIGN.md section 9)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    block, n, nd = local_block(ctx, hop, samples_in_frame)
    gathered = gather_blocks(block, group)
    if distributed_lags:
        mine = torch.tensor([own_lag_from_gathered(ctx, gathered, rank, n, nd)], dtype=torch.int32, device=gathered.device)
        every = torch.empty(world, dtype=torch.int32, device=gathered.device)
        dist.all_gather_into_tensor(every, mine, group=group)
        lags = [int(v) for v in every.cpu()]
    else:
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


class PeerExchange:
    """Gather buffers of all ranks mapped into every rank (CUDA IPC over NVLink), for the fused transform + all-gather.

    Every rank allocates one buffer of ``world * block_stride`` complex values with ``tsdrgpu_malloc`` and publishes its
    IPC handle (one ``all_gather_object`` at set-up, never on the data path).  ``stitch_distributed_fused`` then makes each
    rank's forward transforms store their result into all buffers directly (``tsdrgpu_superb_local_spectra_scatter``); what
    is left of the exchange is a barrier."""

    def __init__(self, ctx: api.Context, n_max_complex: int, group=None):
        import torch.distributed as dist
        self.ctx, self.group = ctx, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.block_stride = int(n_max_complex)
        self.nbytes = 8 * self.block_stride * self.world
        lib = ctx._lib
        own = C.c_void_p()
        ctx.chk(lib.tsdrgpu_malloc(ctx._h, self.nbytes, C.byref(own)))
        self.own = own.value
        handle = (C.c_uint8 * 64)()
        ctx.chk(lib.tsdrgpu_ipc_export(ctx._h, C.c_void_p(self.own), handle))
        handles: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.peers: List[int] = []
        self._opened: List[int] = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.peers.append(self.own)
                continue
            p = C.c_void_p()
            ctx.chk(lib.tsdrgpu_ipc_import(ctx._h, (C.c_uint8 * 64).from_buffer_copy(h), C.byref(p)))
            self.peers.append(p.value); self._opened.append(p.value)
        self._arr = (C.c_void_p * self.world)(*self.peers)
        self._flag = torch.zeros(1, dtype=torch.int32, device=f"cuda:{torch.cuda.current_device()}")

    def gathered(self) -> torch.Tensor:
        """This rank's gather buffer as a float32 tensor view (valid after ``barrier``)."""
        return _device_view(self.own, self.nbytes // 4)

    def barrier(self) -> None:
        """All ranks have finished storing into all buffers: a 4-byte all-reduce ordered behind the transforms on the stream."""
        import torch.distributed as dist
        dist.all_reduce(self._flag, group=self.group)

    def close(self) -> None:
        import torch.distributed as dist
        lib = self.ctx._lib
        torch.cuda.synchronize()
        dist.barrier(group=self.group)             # nobody still stores into a buffer that is about to be unmapped / freed
        for p in self._opened:
            lib.tsdrgpu_ipc_release(self.ctx._h, C.c_void_p(p))
        self._opened = []
        dist.barrier(group=self.group)
        if self.own:
            lib.tsdrgpu_free(self.ctx._h, C.c_void_p(self.own)); self.own = 0


def _device_view(ptr: int, nfloats: int) -> torch.Tensor:
    """A float32 torch tensor over device memory this package allocated itself (``__cuda_array_interface__``)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(h, device=f"cuda:{torch.cuda.current_device()}")


def stitch_distributed_fused(ctx: api.Context, hop: torch.Tensor, samples_in_frame: int, ex: PeerExchange):
    """``stitch_distributed`` with the all-gather fused into the forward transforms (peer stores over NVLink)."""
    pairs = hop.numel() // 2
    hn, hnd = C.c_uint32(0), C.c_uint32(0)
    ex.barrier()                                   # every rank is done reading its buffer from the previous stitch
    ctx.chk(ctx._lib.tsdrgpu_superb_local_spectra_scatter(ctx._h, ctx.stream, hop.data_ptr(), pairs, samples_in_frame, ex._arr,
                                                          ex.world, ex.rank, ex.block_stride, C.byref(hn), C.byref(hnd)))
    ex.barrier()
    n, nd = hn.value, hnd.value
    gathered = ex.gathered()
    lags = (C.c_int * ex.world)()
    ctx.chk(ctx._lib.tsdrgpu_superb_lags(ctx._h, ctx.stream, gathered.data_ptr(), ex.world, ex.block_stride, n, nd, lags))
    out = torch.empty(2 * n, dtype=torch.float32, device=hop.device)
    ctx.chk(ctx._lib.tsdrgpu_superb_residue_ifft_lag(ctx._h, ctx.stream, gathered.data_ptr(), ex.world, ex.block_stride, n,
                                                     residue_of_rank(ex.rank, ex.world), lags, out.data_ptr()))
    return out, list(lags), n


# ---------------------------------------------------------------------------------------------------------------------
# The product path (csrc/superb_mgpu.cu, tsdrgpu_superb_mgpu_*): everything above the C-ABI is bookkeeping.
class SuperbGroup:
    """One rank's handle on a superbandwidth group (tsdrgpu_superb_mgpu_t): rank q owns hop q, the root ends up with the
    time-contiguous magnitude stream.  No collective library is involved in a stitch: the kernels exchange data and flags
    through windows mapped into every rank.

    * one process per GPU (torchrun): ``SuperbGroup.for_process_group(ctx, max_pairs)`` -- the 64-byte CUDA IPC handles of the
      windows travel once, at set-up, through ``torch.distributed.all_gather_object``;
    * all ranks in one process: ``SuperbGroup.local(ctxs, max_pairs)`` (one context per device; several contexts on ONE
      device also work and let a single GPU replay the whole dataflow).
    """

    def __init__(self, ctx: api.Context, nranks: int, rank: int, root: int, max_pairs: int):
        self.ctx, self.nranks, self.rank, self.root = ctx, nranks, rank, root
        h = C.c_void_p()
        ctx.chk(ctx._lib.tsdrgpu_superb_mgpu_create(ctx._h, nranks, rank, root, max_pairs, C.byref(h)))
        self._h = h

    @classmethod
    def for_process_group(cls, ctx: api.Context, max_pairs: int, root: int = 0, group=None) -> "SuperbGroup":
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        g = cls(ctx, world, rank, root, max_pairs)
        handle = (C.c_uint8 * 64)()
        ctx.chk(ctx._lib.tsdrgpu_superb_mgpu_export(g._h, handle))
        handles: List[Optional[bytes]] = [None] * world
        dist.all_gather_object(handles, bytes(handle), group=group)
        blob = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(handles))
        ctx.chk(ctx._lib.tsdrgpu_superb_mgpu_connect_ipc(g._h, blob))
        g._spans_processes = True
        dist.barrier(group=group)
        return g

    @classmethod
    def local(cls, ctxs: Sequence[api.Context], max_pairs: int, root: int = 0) -> List["SuperbGroup"]:
        gs = [cls(c, len(ctxs), r, root, max_pairs) for r, c in enumerate(ctxs)]
        arr = (C.c_void_p * len(gs))(*[g._h for g in gs])
        gs[0].ctx.chk(gs[0].ctx._lib.tsdrgpu_superb_mgpu_connect_local(arr, len(gs)))
        return gs

    def stitch(self, hop: torch.Tensor, samples_in_frame: int, out: Optional[torch.Tensor] = None, hop0: Optional[torch.Tensor] = None,
               in_place: bool = False):
        """This rank's share of one stitch, asynchronous on the current stream of the context's device.  On the root returns
        the magnitude stream (nranks * N floats, time-contiguous); elsewhere None.  hop0: this device's copy of hop 0 (the
        alignment reference), passed by every rank or by none.  in_place (root): no copy out of the group's window -- the
        returned tensor IS the window and is valid until this rank's next stitch."""
        pairs = hop.numel() // 2
        n = self.ctx.fft_getrealsize(pairs)
        if in_place:
            hn = C.c_uint32(0)
            self.ctx.chk(self.ctx._lib.tsdrgpu_superb_mgpu_stitch(self._h, self.ctx.stream, hop.data_ptr(), hop0.data_ptr() if hop0 is not None else None, pairs,
                                                                  samples_in_frame, None, C.byref(hn)))
            if self.rank != self.root:
                return None
            p = C.c_void_p()
            self.ctx.chk(self.ctx._lib.tsdrgpu_superb_mgpu_stream_window(self._h, C.byref(p)))
            return _device_view(p.value, self.nranks * hn.value)
        if self.rank == self.root and out is None:
            out = torch.empty(self.nranks * n, dtype=torch.float32, device=hop.device)
        hn = C.c_uint32(0)
        self.ctx.chk(self.ctx._lib.tsdrgpu_superb_mgpu_stitch(self._h, self.ctx.stream, hop.data_ptr(), hop0.data_ptr() if hop0 is not None else None, pairs, samples_in_frame,
                                                              out.data_ptr() if out is not None else None, C.byref(hn)))
        return (out[: self.nranks * hn.value] if self.rank == self.root else None)

    def lags(self) -> List[int]:
        """Alignment lags (complex samples) of the last stitch; synchronises and raises if a rank went missing."""
        lags = (C.c_int * self.nranks)()
        status = C.c_uint32(0)
        self.ctx.chk(self.ctx._lib.tsdrgpu_superb_mgpu_lags(self._h, self.ctx.stream, lags, C.byref(status)))
        return list(lags)

    def close(self, group=None):
        """Collective when the group spans processes: everybody unmaps, a barrier, everybody frees."""
        if getattr(self, "_h", None):
            self.ctx._lib.tsdrgpu_superb_mgpu_disconnect(self._h)
            if getattr(self, "_spans_processes", False):
                import torch.distributed as dist
                dist.barrier(group=group)
            self.ctx._lib.tsdrgpu_superb_mgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.ctx._lib.tsdrgpu_superb_mgpu_destroy(self._h)
                self._h = None
        except Exception:
            pass
