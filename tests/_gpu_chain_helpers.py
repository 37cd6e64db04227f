"""Shared by the GPU chain-drift tests (tests/ is on sys.path under pytest's default import mode)."""
import math

import pytest
import torch

from oracle import irsde_oracle as O


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _chain_pair(lib, precisions, assisted, T=100, nf=64, depth=4, hw=256):
    """x0 of the same T-step reverse_sde chain (same x_T, same z) under each precision.
    assisted=False: eps-hat = net(x, mu, t), random weights - the reverse SDE's own drift term expands by prod(1 + theta_t dt) =
    1/eps = 200x and a random network does not cancel it, so ANY per-step difference is amplified ~200x.
    assisted=True: eps-hat = eps_true(x, x0, t) + 0.1 * net(x, mu, t) - the analytic noise of a known clean image
    (sde_utils.py:231-232) makes the chain contract toward x0 like a trained model does, while the network (and its
    precision) still perturbs every step: the conditioning a real checkpoint gives, without a checkpoint."""
    dev = _dev()
    P = O.make_weights(3, 3, nf, depth, seed=0)
    g = torch.Generator().manual_seed(1234)
    lq = torch.rand(1, 3, hw, hw, generator=g)
    x0_true = torch.rand(1, 3, hw, hw, generator=g)
    res, xT, zs = {}, None, None
    for prec in precisions:
        net = lib.ConditionalUNet(3, 3, nf, depth=depth, precision=prec)
        net.load_state_dict(P, strict=True)
        net = net.to(dev)
        sde = lib.IRSDE(10, T, schedule="cosine", eps=0.005, device=dev)
        sde.set_model(net)
        sde.set_mu(lq.to(dev))
        if xT is None:
            xT = (lq + torch.randn(lq.shape, generator=g) * sde.max_sigma).to(dev)
            zs = torch.randn((T,) + tuple(lq.shape), generator=g).to(dev)
        if not assisted:
            res[prec] = sde.reverse_sde(xT, zs=zs).cpu()
        else:
            x, x0d = xT.clone(), x0_true.to(dev)
            for i, t in enumerate(reversed(range(1, T + 1))):
                eps_hat = sde.get_real_noise(x, x0d, t) + 0.1 * sde.noise_fn(x, t)
                x = sde._native_step(lib._lib.MODE_SDE, x, sde.mu, eps_hat, zs[i], t)
            res[prec] = x.cpu()
        del net, sde
    return res


def _drift_line(name, a, b):
    d = a - b
    peak = b.abs().max().item()
    rms = d.pow(2).mean().sqrt().item()
    psnr = 20 * math.log10(peak / rms) if rms > 0 else float("inf")
    return "%s: max|d| %.3e  rms %.3e  max|x0| %.3g  PSNR(peak=max|x0|) %.1f dB" % (name, d.abs().max().item(), rms, peak, psnr), d, peak, psnr
