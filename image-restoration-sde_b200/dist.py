"""Batch sharding across the GPUs of one box (one process per GPU, torch.distributed plumbing).

The sampler has no cross-image dependency (every op of the network is per image), so
rank r simply owns images [r*B/R, (r+1)*B/R): weights are broadcast once, the T steps run with zero
communication, and x0 is gathered once at the end.  This replaces the reference's
``DataParallel`` (models/denoising_model.py:41-42), which re-broadcasts all parameters and
scatters/gathers the batch on every one of the T forwards.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


def shard_range(B, rank, world):
    """Contiguous, order-preserving split; the first B % world ranks get one extra image."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_weights(model, src=0):
    """One broadcast of every parameter from ``src`` (packed into a single flat buffer)."""
    params = [p for p in model.parameters()]
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    dist.broadcast(flat, src=src)
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))
            off += n
    return flat.numel()


def sharded_reverse(run, xT, mu=None, zs=None, group=None, gather=True, sde=None):
    """Run ``run(xT_slice, mu_slice, zs_slice)`` on this rank's slice of the batch and gather x0.

    ``run`` is e.g. ``lambda x, m, z: (sde.set_mu(m), sde.reverse_sde(x, zs=z))[1]``.  ``zs`` is
    [T,B,...] drawn for the FULL batch, so that every image sees the same noise regardless of the
    number of ranks (bit-identical to the single-GPU result).  With the in-kernel Philox (``zs`` None,
    ``sde.rng == "philox"``) pass ``sde``: its ``image_base`` is moved to this rank's first image for the call, which
    gives the same bit-identical guarantee without materialising any noise.  Returns the full [B,...] tensor on
    every rank (all_gather), or the local slice if ``gather`` is False.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = xT.shape[0]
    lo, hi = shard_range(B, rank, world)
    base = None
    if sde is not None:
        base = sde.image_base
        sde.image_base = base + lo
    try:
        out = run(xT[lo:hi], None if mu is None else mu[lo:hi], None if zs is None else zs[:, lo:hi])
    finally:
        if sde is not None:
            sde.image_base = base
    if not gather or world == 1:
        return out
    sizes = [shard_range(B, r, world) for r in range(world)]
    maxn = max(h - l for l, h in sizes)
    pad = torch.zeros((maxn,) + tuple(out.shape[1:]), dtype=out.dtype, device=out.device)
    pad[:hi - lo] = out
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][:sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)


# ---- NCCL behind the C ABI (include/irsde_b200.h: irsde_comm_*) -------------------------------------------------------
def comm_unique_id():
    """128-byte NCCL id (bytes); rank 0 creates it and hands it to the other ranks out of band (file, pipe, TCP store)."""
    buf = ctypes.create_string_buffer(128)
    _lib.check(_lib.load().irsde_comm_unique_id(buf))
    return bytes(buf.raw)


class NativeComm:
    """The path's two collectives through the library's own NCCL communicator, for hosts that do not use
    torch.distributed: ``broadcast_weights`` once, T steps with no communication, one ``gather`` of x0."""

    def __init__(self, model, uid, rank, nranks, device=None):
        self.model, self.rank, self.nranks = model, rank, nranks
        self.ctx = model.sync_weights(device)
        _lib.check(self.ctx.L.irsde_comm_init(self.ctx.h, uid, rank, nranks), self.ctx.h)

    def broadcast_weights(self, src=0):
        """Rank ``src``'s parameters replace every rank's inside the native context (one NCCL group); the nn.Module's
        own tensors are refreshed from it lazily only if the caller asks (``pull=True`` is not needed for sampling)."""
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self.ctx.L.irsde_broadcast_weights(self.ctx.h, src, ctypes.c_void_p(st)), self.ctx.h)

    def gather(self, x_local, B_total):
        """All ranks receive the [B_total, ...] batch in rank order (the contiguous partition of ``shard_range``)."""
        per = x_local[0].numel() if x_local.shape[0] else 0
        sizes = [shard_range(B_total, r, self.nranks) for r in range(self.nranks)]
        if per == 0:
            raise ValueError("gather needs at least one image per rank to learn the image shape")
        counts = (ctypes.c_int64 * self.nranks)(*[(hi - lo) * per for lo, hi in sizes])
        x_local = x_local.contiguous().float()
        out = torch.empty((B_total,) + tuple(x_local.shape[1:]), device=x_local.device, dtype=torch.float32)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self.ctx.L.irsde_gather(self.ctx.h, ctypes.c_void_p(x_local.data_ptr()), ctypes.c_void_p(out.data_ptr()), counts,
                                           ctypes.c_void_p(st)), self.ctx.h)
        return out
