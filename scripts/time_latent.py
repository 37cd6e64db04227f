"""Time the Refusion latent path (SURVEY §8 d, config C4 shape) on one GPU: UNet.encode -> NAFNet reverse_sde -> UNet.decode.
Random-init weights of the shipped architecture (latent-dehazing/options/dehazing/test/nasde.yml:7-45). Prints one JSON line."""
import ctypes, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irsde_b200
from irsde_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
T = 100
dev = torch.device("cuda:0")
torch.manual_seed(0)
ae = irsde_b200.UNet(3, 3, 8, [4, 8, 8, 16], 8, precision="bf16").to(dev)
net = irsde_b200.ConditionalNAFNet(img_channel=8, width=64, enc_blk_nums=[1, 1, 1, 28], middle_blk_num=1, dec_blk_nums=[1, 1, 1, 1],
                                   latent=True, precision="bf16").to(dev)
sde = irsde_b200.IRSDE(max_sigma=50, T=T, schedule="cosine", eps=0.005, device=dev)
sde.set_model(net)
sde.rng = "philox"
x = torch.rand(B, 3, HW, HW, device=dev)


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


ms_enc, (z, h) = timed(lambda: ae.encode(x))
sde.set_mu(z)
xT = sde.noise_state(z)
ms_chain, z0 = timed(lambda: sde.reverse_sde(xT), reps=2, warm=1)
z, h = ae.encode(x)
ms_dec, y = timed(lambda: ae.decode(z0, h))
assert torch.isfinite(y).all() and torch.isfinite(z0).all()

ctx = net._ctx
ncat = 6
ms_a, fl_a, n_a = (ctypes.c_double * ncat)(), (ctypes.c_double * ncat)(), (ctypes.c_int64 * ncat)()
_lib.check(ctx.L.irsde_profile_begin(ctx.h), ctx.h)
sde.reverse_sde(xT, T=5)
_lib.check(ctx.L.irsde_profile_end(ctx.h, ms_a, fl_a, n_a, ncat), ctx.h)
names = ["tcgen05_conv", "simt_conv", "layernorm", "attention", "misc", "update"]
bd = {names[i]: {"ms_per_step": round(ms_a[i] / 5, 4), "launches_per_step": n_a[i] / 5,
                 "tflops": round(fl_a[i] / (ms_a[i] * 1e-3) / 1e12, 1) if ms_a[i] > 0 and fl_a[i] > 0 else None}
      for i in range(ncat) if n_a[i] > 0}
print(json.dumps({"B": B, "HW": HW, "T": T, "latent": list(z.shape), "ms_encode": ms_enc, "ms_chain": ms_chain, "ms_decode": ms_dec,
                  "ms_per_step": ms_chain / T, "img_per_s": B / ((ms_enc + ms_chain + ms_dec) * 1e-3), "nafnet_step_breakdown": bd}))
