"""Batch sharding across the GPUs of one box (one process per GPU, torch.distributed plumbing).

The sampler has no cross-image dependency (every op of the network is per image), so
rank r simply owns images [r*B/R, (r+1)*B/R): weights are broadcast once, the T steps run with zero
communication, and x0 is gathered once at the end.  This replaces the reference's
``DataParallel`` (models/denoising_model.py:41-42), which re-broadcasts all parameters and
scatters/gathers the batch on every one of the T forwards.
"""
import torch
import torch.distributed as dist


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
