"""Small workloads for ncu captures of (1) the tensor-core full-attention kernel (denoising-sde UNet forward, 1 024 keys) and
(2) the fp32x3 conv kernels (a few forwards of the nf=64 UNet in precision="fp32x3")."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irsde_b200
dev = torch.device("cuda:0")
torch.manual_seed(0)
which = sys.argv[1]
if which == "attn":
    net = irsde_b200.DenoisingUNet(3, 3, 64, depth=4, precision="bf16").to(dev)   # mid level of a 256^2 input: 32^2 = 1024 tokens
    x = torch.rand(8, 3, 256, 256, device=dev)
    for _ in range(3):
        y = net(x, 17)
else:
    net = irsde_b200.ConditionalUNet(3, 3, 64, depth=4, precision="fp32x3").to(dev)
    x = torch.rand(8, 3, 256, 256, device=dev)
    for _ in range(2):
        y = net(x, x, 17)
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
