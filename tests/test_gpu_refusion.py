"""Refusion tile mode (BASELINE config 4, SURVEY 8 e option (i)): tile mode is DEFINED as "the reference applied to each
latent tile", so the oracle is run tile by tile - encode (whole image) -> per-tile reverse_sde through ConditionalNAFNet
(zero-padded to the network's own multiple, NAFNet_arch.py:183-188) -> stitch -> decode - and the CUDA pipeline
(irsde_b200.TiledRefusion, tiles batched by shape) must reproduce it.  The per-tile noise comes from a pre-drawn table
keyed by the tile's global uid, so the oracle sees the same x_T and z as the device run."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import irsde_oracle as O


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("precision,tile", [("fp32", 8), ("fp32", None), ("bf16", 8)])
def test_tiled_refusion_vs_oracle_per_tile(precision, tile):
    import irsde_b200
    dev = _dev()
    ae_cfg = dict(in_ch=3, out_ch=3, ch=8, ch_mult=[1, 2], embed_dim=4)
    naf_cfg = dict(img_channel=4, width=16, middle_blk_num=1, enc_blk_nums=[1, 1], dec_blk_nums=[1, 1])
    Pae = O.make_latent_unet_weights(3, 3, 8, [1, 2], 4, seed=2)
    Pn = O.make_nafnet_weights(4, 16, 1, [1, 1], [1, 1], seed=3)
    g = torch.Generator().manual_seed(9)
    x = torch.rand(2, 3, 40, 56, generator=g)               # latent 20 x 28: tile 8 -> 3 x 4 grid with ragged edges
    T = 5
    sc = O.Schedule(50, T, "cosine", 0.005)
    zo, ho = O.latent_unet_encode(Pae, x, ae_cfg["ch_mult"])
    units, groups = irsde_b200.plan_units(2, zo.shape[2], zo.shape[3], tile)
    noise = {u: (torch.randn(4, b[1] - b[0], b[3] - b[2], generator=g), torch.randn(T, 4, b[1] - b[0], b[3] - b[2], generator=g))
             for u, _, b in units}
    # ---- oracle: the reference pipeline on every tile independently
    ref_lat = torch.empty_like(zo)
    for u, b, (y0, y1, x0, x1) in units:
        mu = zo[b:b + 1, :, y0:y1, x0:x1]
        xT = mu + noise[u][0][None] * sc.max_sigma
        fn = lambda xx, t: O.nafnet_forward(Pn, xx, mu, t, 16, [1, 1], 1, [1, 1], latent=True)
        ref_lat[b:b + 1, :, y0:y1, x0:x1] = O.reverse_chain(sc, fn, xT, mu, noise[u][1][:, None], "sde")
    ref_img = O.latent_unet_decode(Pae, ref_lat, ho, ae_cfg["ch_mult"], 40, 56)
    # ---- device
    ae = irsde_b200.UNet(precision=precision, **ae_cfg)
    ae.load_state_dict(Pae, strict=True)
    ae = ae.to(dev)
    net = irsde_b200.ConditionalNAFNet(latent=True, precision=precision, **naf_cfg)
    net.load_state_dict(Pn, strict=True)
    net = net.to(dev)
    sde = irsde_b200.IRSDE(50, T, schedule="cosine", eps=0.005, device=dev)
    sde.set_model(net)

    def chain(mu_tiles, uids):
        xT = mu_tiles + torch.stack([noise[u][0] for u in uids]).to(dev) * sde.max_sigma
        zs = torch.stack([noise[u][1] for u in uids], dim=1).to(dev)      # [T, n, C, th, tw]
        sde.set_mu(mu_tiles)
        return sde.reverse_sde(xT, zs=zs)

    pipe = irsde_b200.TiledRefusion(ae, sde, tile=tile, chain=chain)
    lat = pipe.restore_latent(ae.encode(x.to(dev))[0])
    out, (lo, hi) = pipe.restore(x.to(dev))
    assert (lo, hi) == (0, 2) and out.shape == ref_img.shape
    dl = (lat.cpu() - ref_lat).abs().max().item()
    di = (out.cpu() - ref_img).abs().max().item()
    if precision == "fp32":
        assert dl < 1e-3 and di < 1e-3 * max(1.0, ref_img.abs().max().item()), (dl, di)
    else:
        assert dl < 4e-2 * ref_lat.abs().max().item() and di < 6e-2 * ref_img.abs().max().item(), (dl, di)


def test_tiled_refusion_philox_rank_independent():
    """With the in-kernel Philox keyed by the tile's global uid, processing the tiles in one batch or one by one (= any
    sharding over ranks) gives bit-identical latents."""
    import irsde_b200
    dev = _dev()
    torch.manual_seed(0)
    net = irsde_b200.ConditionalNAFNet(img_channel=4, width=16, middle_blk_num=1, enc_blk_nums=[1, 1], dec_blk_nums=[1, 1],
                                       latent=True, precision="fp32").to(dev)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("beta") or n.endswith("gamma"):
                p.fill_(0.3)
    sde = irsde_b200.IRSDE(50, 6, eps=0.005, device=dev)
    sde.set_model(net)
    z = torch.rand(2, 4, 16, 24, device=dev)
    pipe = irsde_b200.TiledRefusion(None, sde, tile=8, mode="sde", seed=5)
    full = pipe.restore_latent(z)
    units, _ = irsde_b200.plan_units(2, 16, 24, 8)
    one = torch.empty_like(z)
    for u, b, (y0, y1, x0, x1) in units:
        one[b:b + 1, :, y0:y1, x0:x1] = pipe.chain(z[b:b + 1, :, y0:y1, x0:x1].contiguous(), [u])
    assert torch.equal(full, one)
    assert not torch.equal(full, irsde_b200.TiledRefusion(None, sde, tile=8, mode="sde", seed=6).restore_latent(z))
