"""Multi-GPU host logic (one process per GPU, torch.distributed): screen-tile sharding with a gather of
the bands to rank 0, and z-slab sharding of the distance-field regeneration with one boundary-plane
exchange.  The reference is single-GPU; this is the B200 scale-out of its pass interface.

Everything here is plumbing over `torch.distributed` (NCCL on GPUs, gloo in the CPU tests); the arithmetic
lives in libvxrt_cuda.so (`vxrt_cuda_df_slab_phase_a/b`, the `vxrt_tile` field of every pass).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .pipeline import band_rows


def slab_bounds(nz: int, world: int) -> list[int]:
    """Plane ranges [z0[s], z0[s+1]) of the z-slabs; equal sizes when world divides nz."""
    base, extra = divmod(nz, world)
    z0 = [0]
    for s in range(world):
        z0.append(z0[-1] + base + (1 if s < extra else 0))
    return z0


def bind_streams(ctx, device=None) -> None:
    """Stream contract of everything in this module: torch's collectives are ordered against torch's CURRENT stream only, while an
    engine.Context launches on its own non-blocking stream by default.  Zero-copy views of the context's memory
    (Context.df_device_array, Context.attachment_as_device_array) handed to a collective are therefore only ordered against the
    passes that wrote them when both use one stream.  This puts the context on torch's current stream of its device; the sharded
    entry points below call it themselves, so a caller in the default state cannot get a collective that overtakes a kernel."""
    dev = torch.device("cuda", ctx.device) if device is None else torch.device(device)
    if dev.type != "cuda":
        return
    # torch's default stream is the NULL stream; set_stream(NULL) means "the context's own stream", so name it by the
    # runtime's explicit handle for the legacy default stream (cudaStreamLegacy == 0x1)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream or 0x1)


def gather_bands(full: torch.Tensor, height: int, rank: int, world: int, dst: int = 0, band: int = 8, group=None, ctx=None):
    """`full` is a full-frame attachment [H, ...] of which this rank rendered rows band_rows(height, rank, world).
    After the call rank `dst` holds every band (rows are contiguous in memory, so the receives land in place).
    Pass the `ctx` that rendered `full` so the transfer is ordered behind its passes (see bind_streams)."""
    if ctx is not None and full.is_cuda:
        bind_streams(ctx, full.device)
    if world == 1:
        return full
    result = full
    if full.dtype not in (torch.uint8, torch.float32, torch.float16, torch.int32):
        full = full.view(torch.uint8)  # NCCL moves bytes; not every dtype (e.g. int16) is accepted
    ops = []
    if rank == dst:
        for r in range(world):
            if r == dst:
                continue
            row0, rows = band_rows(height, r, world, band)
            if rows > 0:
                ops.append(dist.P2POp(dist.irecv, full[row0:row0 + rows], r, group=group))
    else:
        row0, rows = band_rows(height, rank, world, band)
        if rows > 0:
            ops.append(dist.P2POp(dist.isend, full[row0:row0 + rows], dst, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return result


class CudaSlabBackend:
    """Adapter from an engine.Context to the operations regenerate_distance_field_sharded needs."""

    def __init__(self, ctx, device):
        self.ctx = ctx
        self.device = device
        nx, ny, nz = ctx.dims
        self.nz = nz
        self.df = torch.as_tensor(ctx.df_device_array(), device=device)  # zero-copy [nz, ny, nx] view

    # every phase re-binds the context to torch's current stream: the all-gathers between the phases read / write the
    # context's field through zero-copy views and are ordered against that stream only (bind_streams)
    def phase_a(self, slab, z0):
        bind_streams(self.ctx, self.device)
        self.ctx.df_slab_phase_a(slab, z0)

    def plane(self, z):
        return self.df[z]

    def phase_b(self, slab, z0, firsts: torch.Tensor, lasts: torch.Tensor):
        bind_streams(self.ctx, self.device)
        self.ctx.df_slab_phase_b(slab, z0, firsts.data_ptr(), lasts.data_ptr())

    def commit(self):
        bind_streams(self.ctx, self.device)
        self.ctx.df_commit()


def regenerate_distance_field_sharded(backend, rank: int, world: int, group=None):
    """Distance-field regeneration sharded by z-slabs:
       phase A (slab-local X, Y, Z sweeps) -> all-gather of each rank's first / last plane ->
       phase B (apply the other slabs' carries) -> all-gather of the slabs (restores the replicated field)."""
    z0 = slab_bounds(backend.nz, world)
    backend.phase_a(rank, z0)
    if world == 1:
        backend.commit()
        return z0
    first = backend.plane(z0[rank]).contiguous()
    last = backend.plane(z0[rank + 1] - 1).contiguous()
    firsts = torch.empty((world,) + tuple(first.shape), dtype=first.dtype, device=first.device)
    lasts = torch.empty_like(firsts)
    # the path's one real exchange step: 2 * nx*ny bytes per rank
    dist.all_gather_into_tensor(firsts.view(-1), first.view(-1), group=group)
    dist.all_gather_into_tensor(lasts.view(-1), last.view(-1), group=group)
    backend.phase_b(rank, z0, firsts, lasts)
    sizes = {z0[s + 1] - z0[s] for s in range(world)}
    if len(sizes) == 1:
        dist.all_gather_into_tensor(backend.df.view(-1), backend.df[z0[rank]:z0[rank + 1]].clone().view(-1), group=group)
    else:  # uneven slabs: one broadcast per slab
        for s in range(world):
            dist.broadcast(backend.df[z0[s]:z0[s + 1]], src=s, group=group)
    backend.commit()
    return z0


class NumpySlabBackend:
    """CPU stand-in with the same algebra (used by the gloo tests of the orchestration): the slab-local
    sweeps and the carry application are written with numpy on a [nz, ny, nx] uint8 array."""

    def __init__(self, blocks: np.ndarray):
        self.blocks = blocks
        self.nz = blocks.shape[0]
        self.df = torch.zeros(blocks.shape, dtype=torch.uint8)

    @staticmethod
    def _sweep(a: np.ndarray, axis: int):
        a = np.moveaxis(a, axis, 0)
        for i in range(1, a.shape[0]):
            np.minimum(a[i], a[i - 1] + 1, out=a[i])
        for i in range(a.shape[0] - 2, -1, -1):
            np.minimum(a[i], a[i + 1] + 1, out=a[i])

    def phase_a(self, slab, z0):
        nz, ny, nx = self.blocks.shape
        maxd = min(254, nx + ny + nz)
        sl = slice(z0[slab], z0[slab + 1])
        d = np.where(self.blocks[sl] > 0, 0, maxd).astype(np.int32)
        self._sweep(d, 2)
        self._sweep(d, 1)
        self._sweep(d, 0)
        self.df[sl] = torch.from_numpy(d.astype(np.uint8))

    def plane(self, z):
        return self.df[z]

    def phase_b(self, slab, z0, firsts, lasts):
        z = np.arange(z0[slab], z0[slab + 1])[:, None, None]
        d = self.df[z0[slab]:z0[slab + 1]].numpy().astype(np.int32)
        for t in range(len(z0) - 1):
            if t < slab:
                d = np.minimum(d, lasts[t].numpy().astype(np.int32)[None] + (z - (z0[t + 1] - 1)))
            elif t > slab:
                d = np.minimum(d, firsts[t].numpy().astype(np.int32)[None] + (z0[t] - z))
        self.df[z0[slab]:z0[slab + 1]] = torch.from_numpy(d.astype(np.uint8))

    def commit(self):
        pass
