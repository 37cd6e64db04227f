"""Three NAFNet sampler steps (latent 8x128x128, the shipped latent-dehazing architecture, bf16, no graph) for an
ncu launch list: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/naf_steps_for_ncu.py`."""
import os, sys
os.environ["IRSDE_B200_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irsde_b200
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = irsde_b200.ConditionalNAFNet(img_channel=8, width=64, enc_blk_nums=[1, 1, 1, 28], middle_blk_num=1, dec_blk_nums=[1, 1, 1, 1],
                                   latent=True, precision="bf16").to(dev)
sde = irsde_b200.IRSDE(max_sigma=50, T=100, schedule="cosine", eps=0.005, device=dev)
sde.set_model(net)
sde.rng = "philox"
z = torch.rand(int(sys.argv[1]) if len(sys.argv) > 1 else 1, 8, 128, 128, device=dev)
sde.set_mu(z)
x = sde.reverse_sde(sde.noise_state(z), T=3)
torch.cuda.synchronize()
print("done", float(x.abs().mean()))
