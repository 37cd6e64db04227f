"""Golden fixture for ConditionalNAFNet (Refusion score network), generated from the imported reference.
    python tests/golden/make_golden_nafnet.py
"""
import importlib
import os
import sys

import torch

REF = "/root/reference/codes"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_arch(task):
    """Load module_util + DenoisingNAFNet_arch of one task by file path (the latent task's package __init__ imports
    timm, which is absent here)."""
    import importlib.util
    import types
    base = os.path.join(REF, "config", task, "models", "modules")
    pkg = types.ModuleType("refmods_" + task.replace("-", "_"))
    pkg.__path__ = [base]
    sys.modules[pkg.__name__] = pkg
    out = {}
    for name in ("module_util", "local_arch", "DenoisingNAFNet_arch"):
        f = os.path.join(base, name + ".py")
        if not os.path.exists(f):
            continue
        spec = importlib.util.spec_from_file_location(pkg.__name__ + "." + name, f)
        m = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = m
        spec.loader.exec_module(m)
        out[name] = m
    return out["DenoisingNAFNet_arch"]


def main():
    torch.set_num_threads(4)
    gold = {}
    g = torch.Generator().manual_seed(5)
    for key, task, latent in (("nafnet", "deraining", False), ("nafnet_latent", "latent-dehazing", True)):
        arch = load_arch(task)
        cfg = dict(img_channel=3 if not latent else 4, width=8, middle_blk_num=1, enc_blk_nums=[1, 2], dec_blk_nums=[1, 1])
        torch.manual_seed(2)
        net = arch.ConditionalNAFNet(**cfg).eval()
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        for k in sd:  # beta/gamma are zero-initialised in the reference: make every branch matter
            if k.endswith("beta") or k.endswith("gamma"):
                sd[k] = 0.5 * torch.randn(sd[k].shape, generator=g)
            if k.endswith(".g"):
                sd[k] = 1 + 0.2 * torch.randn(sd[k].shape, generator=g)
        net.load_state_dict(sd)
        C = cfg["img_channel"]
        x = torch.rand(2, C, 18, 23, generator=g)
        cond = torch.rand(2, C, 18, 23, generator=g)
        with torch.no_grad():
            y = net(x, cond, 13)
            yv = net(x, cond, torch.tensor([3, 40]))
        gold[key] = dict(cfg=cfg, latent=latent, state=sd, x=x, cond=cond, t_int=13, y=y, t_vec=torch.tensor([3, 40]), y_vec=yv)
    torch.save(gold, os.path.join(OUT, "reference_golden_nafnet.pt"))
    print("wrote", os.path.getsize(os.path.join(OUT, "reference_golden_nafnet.pt")), "bytes")


if __name__ == "__main__":
    main()
