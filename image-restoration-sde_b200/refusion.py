"""Refusion latent restoration with the chain sharded over independent latent TILES (BASELINE config 4, SURVEY 8 e).

The reference loop (codes/config/latent-dehazing/test.py:88-96):

    latent_LQ, hidden = latent_model.encode(LQ)          # UNet.encode          (UNet_arch.py:59-76)
    noisy_state = sde.noise_state(latent_LQ)
    model.feed_data(noisy_state, latent_LQ, GT); model.test(sde, hidden=hidden)   # reverse_sde on the latent + decode

runs the whole latent of one image through ConditionalNAFNet.  Tile mode cuts the latent into a grid of tiles and treats
every tile as an independent image of the chain: the result is, by definition, THE REFERENCE APPLIED TO EACH TILE (NAFNet's
global SCA pooling then pools over the tile, not the image; seams are a quality question, SURVEY 8 e option (i)) - which
is what the parity test checks, tile by tile, against the oracle.  Tiles are the sharding unit: the global list of
(image, tile) units is cut contiguously over the ranks, each rank runs its units as one batch per tile shape with no
communication during the T steps, and two small all-gathers (latents in, restored tiles out) frame the chain.  The
in-kernel Philox is keyed by the unit's GLOBAL index, so the result does not depend on the number of ranks.

Encode / decode are per image (the autoencoder has a global LinearAttention at its deepest level): images are sharded
contiguously over the ranks; a rank may own no image (more ranks than images) and still process tiles.
"""
import torch
import torch.distributed as dist

from .dist import shard_range


def tile_boxes(h, w, tile):
    """Row-major grid of (y0, y1, x0, x1) boxes covering an h x w latent with tile x tile tiles (edge tiles are smaller)."""
    if tile is None or tile <= 0:
        return [(0, h, 0, w)]
    return [(y, min(y + tile, h), x, min(x + tile, w)) for y in range(0, h, tile) for x in range(0, w, tile)]


def plan_units(B, h, w, tile):
    """Global unit list [(uid, image, box)] in (image, tile) order and its grouping by tile shape:
    {(th, tw): [uid, ...]} (at most four shapes: interior, right edge, bottom edge, corner)."""
    boxes = tile_boxes(h, w, tile)
    units = [(b * len(boxes) + k, b, box) for b in range(B) for k, box in enumerate(boxes)]
    groups = {}
    for uid, b, (y0, y1, x0, x1) in units:
        groups.setdefault((y1 - y0, x1 - x0), []).append(uid)
    return units, groups


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allgather_rows(t, counts, group=None):
    """Concatenate per-rank row blocks (rank r holds counts[r] rows of identical trailing shape) on every rank."""
    rank, world = _world(group)
    if world == 1:
        return t
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][:counts[r]] for r in range(world)], dim=0)


class TiledRefusion:
    """encode -> tile-sharded reverse chain on the latent -> decode.

    ``ae``: latent autoencoder with ``encode(x) -> (z, h)`` / ``decode(z, h)`` (irsde_b200.UNet);  ``sde``: IRSDE whose
    model is the latent ConditionalNAFNet;  ``tile``: tile edge in latent pixels (None = whole latent, i.e. the
    reference's own per-image behaviour);  ``mode``: "sde" | "posterior" | "ode".
    ``chain`` may be replaced by any callable (xT_tiles, mu_tiles, uids) -> x0_tiles (the CPU gloo test does)."""

    def __init__(self, ae, sde, tile=None, mode="sde", seed=0, group=None, chain=None):
        self.ae, self.sde, self.tile, self.mode, self.seed, self.group = ae, sde, tile, mode, seed, group
        self.chain = chain or self._native_chain

    # -- the chain on a batch of same-shaped tiles: noise_state + reverse_* with per-unit Philox uids
    def _native_chain(self, mu_tiles, uids):
        sde = self.sde
        keep = (sde.rng, sde.seed, sde.seed_auto_increment, sde.image_uids)
        try:
            sde.rng, sde.seed, sde.seed_auto_increment, sde.image_uids = "philox", self.seed, False, list(uids)
            sde.set_mu(mu_tiles)
            xT = sde.noise_state(mu_tiles)
            return getattr(sde, "reverse_" + self.mode)(xT)
        finally:
            sde.rng, sde.seed, sde.seed_auto_increment, sde.image_uids = keep

    def restore_latent(self, z_all):
        """z_all: the FULL latent batch [B, C, h, w] (same on every rank) -> restored latents [B, C, h, w] on every rank."""
        rank, world = _world(self.group)
        B, C, h, w = z_all.shape
        units, groups = plan_units(B, h, w, self.tile)
        out = torch.empty_like(z_all)
        for (th, tw), uids in sorted(groups.items()):
            lo, hi = shard_range(len(uids), rank, world)
            mine = uids[lo:hi]
            if mine:
                mu = torch.stack([z_all[units[u][1], :, units[u][2][0]:units[u][2][1], units[u][2][2]:units[u][2][3]] for u in mine])
                x0 = self.chain(mu.contiguous(), mine)
            else:
                x0 = z_all.new_zeros((0, C, th, tw))
            counts = [shard_range(len(uids), r, world)[1] - shard_range(len(uids), r, world)[0] for r in range(world)]
            allx = allgather_rows(x0, counts, self.group)
            for row, u in enumerate(uids):
                _, b, (y0, y1, x0_, x1) = units[u]
                out[b, :, y0:y1, x0_:x1] = allx[row]
        return out

    @torch.no_grad()
    def restore(self, LQ):
        """LQ: the FULL image batch [B, 3, H, W] on this rank's device (every rank passes the same batch; only the images
        this rank owns are encoded / decoded).  Returns (restored images of this rank's slice, (lo, hi))."""
        rank, world = _world(self.group)
        B = LQ.shape[0]
        lo, hi = shard_range(B, rank, world)
        counts = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
        h = None
        if hi > lo:
            z, h = self.ae.encode(LQ[lo:hi])
        if world > 1:
            # every rank needs every latent: learn the latent shape from whoever owns an image, then all-gather
            shp = torch.tensor(list(z.shape[1:]) if hi > lo else [0, 0, 0], device=LQ.device)
            dist.all_reduce(shp, op=dist.ReduceOp.MAX, group=self.group)
            if hi == lo:
                z = LQ.new_zeros((0,) + tuple(int(v) for v in shp.tolist()))
            z_all = allgather_rows(z.contiguous(), counts, self.group)
        else:
            z_all = z
        x0 = self.restore_latent(z_all)
        if hi > lo:
            return self.ae.decode(x0[lo:hi].contiguous(), h), (lo, hi)
        return LQ.new_zeros((0,) + tuple(LQ.shape[1:])), (lo, hi)
