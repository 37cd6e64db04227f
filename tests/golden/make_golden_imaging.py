"""Generate tests/golden/reference_golden_imaging.npz by running the UNMODIFIED reference helpers
(/root/reference/codes/utils/img_utils.py: tensor2img, img2tensor, calculate_psnr, calculate_ssim) in the build
container.  Run from the repo root: python tests/golden/make_golden_imaging.py"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference/codes/utils/img_utils.py"
spec = importlib.util.spec_from_file_location("ref_img_utils", REF)
iu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(iu)

g = torch.Generator().manual_seed(77)
out = {}
# tensor2img: values beyond [0,1], exact .5/255 ties, odd sizes (scalar tail path) and a 4-divisible size (vector path)
for name, (C, H, W) in {"a": (3, 37, 45), "b": (3, 32, 48), "c": (1, 19, 23)}.items():
    t = torch.rand(C, H, W, generator=g) * 1.3 - 0.15
    ties = (torch.arange(0, H * W) % 255).float().add(0.5).div(255.0)          # k + 0.5 over 255: rounding ties
    t.view(C, -1)[0, : H * W // 2] = ties[: H * W // 2]
    out["t_" + name] = t.numpy()
    out["img_" + name] = iu.tensor2img(t.clone())
    out["img11_" + name] = iu.tensor2img(t.clone() * 2 - 1, min_max=(-1, 1))
x = (torch.rand(41, 53, 3, generator=g) * 255).to(torch.uint8).numpy()
y = np.clip(x.astype(np.int32) + (torch.randn(41, 53, 3, generator=g) * 12).round().to(torch.int32).numpy(), 0, 255).astype(np.uint8)
out["x"], out["y"] = x, y
out["x_tensor"] = iu.img2tensor(x).numpy()
out["psnr"] = np.float64(iu.calculate_psnr(x, y))
out["psnr_crop4"] = np.float64(iu.calculate_psnr(x[4:-4, 4:-4], y[4:-4, 4:-4]))
out["ssim"] = np.float64(iu.calculate_ssim(x, y))
out["ssim_crop4"] = np.float64(iu.calculate_ssim(x[4:-4, 4:-4], y[4:-4, 4:-4]))
out["ssim_gray"] = np.float64(iu.calculate_ssim(x[:, :, 0], y[:, :, 0]))
out["psnr_same"] = np.float64(iu.calculate_psnr(x, x))
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden_imaging.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, os.path.getsize(dst), "bytes;", {k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items()})
