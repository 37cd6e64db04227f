"""TEST INFRASTRUCTURE ONLY - numpy restatement of the reference's image helpers (codes/utils/img_utils.py), the
checker for the device kernels in csrc/imaging.cu.  Pinned to the reference's own functions by
tests/golden/reference_golden_imaging.npz (made by tests/golden/make_golden_imaging.py).  Never imported by the product."""
import math

import numpy as np


def tensor2img(t, min_max=(0, 1)):
    """img_utils.py:136-163 for one image: t float32 [C,H,W] or [H,W] (numpy) -> uint8 [H,W,C] BGR / [H,W]."""
    t = np.asarray(t, dtype=np.float32)
    t = np.clip(t, np.float32(min_max[0]), np.float32(min_max[1]))                    # :142 clamp_
    t = (t - np.float32(min_max[0])) / np.float32(min_max[1] - min_max[0])              # :143
    if t.ndim == 3:
        t = np.transpose(t[[2, 1, 0], :, :], (1, 2, 0))                                 # :151 HWC, BGR
    return (t * np.float32(255.0)).round().astype(np.uint8)                             # :160-162


def img2tensor(img):
    """img_utils.py:171-180: uint8 [H,W,C] BGR -> float32 [C,H,W] RGB in [0,1]."""
    a = img.astype(np.float32) / np.float32(255.0)
    a = a[:, :, [2, 1, 0]]
    return np.ascontiguousarray(np.transpose(a, (2, 0, 1)))


def calculate_psnr(img1, img2):
    """img_utils.py:182-190."""
    mse = np.mean((img1.astype(np.float64) - img2.astype(np.float64)) ** 2)
    if mse == 0:
        return float("inf")
    return 20 * math.log10(255.0 / math.sqrt(mse))


def _gauss11():
    k = np.exp(-((np.arange(11) - 5.0) ** 2) / (2 * 1.5 * 1.5))     # cv2.getGaussianKernel(11, 1.5)
    return k / k.sum()


def _valid_filter(x, win):
    """cv2.filter2D(x, -1, win)[5:-5, 5:-5] (:202): correlation restricted to windows that lie inside the image."""
    H, W = x.shape[:2]
    out = np.zeros((H - 10, W - 10) + x.shape[2:], dtype=np.float64)
    for r in range(11):
        for q in range(11):
            out += win[r, q] * x[r:r + H - 10, q:q + W - 10]
    return out


def ssim(img1, img2):
    """img_utils.py:193-214 (2-D or HxWxC arrays; filter2D treats channels independently)."""
    C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    a, b = img1.astype(np.float64), img2.astype(np.float64)
    k = _gauss11()
    win = np.outer(k, k)
    mu1, mu2 = _valid_filter(a, win), _valid_filter(b, win)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 ** 2, mu2 ** 2, mu1 * mu2
    s1 = _valid_filter(a ** 2, win) - mu1_sq
    s2 = _valid_filter(b ** 2, win) - mu2_sq
    s12 = _valid_filter(a * b, win) - mu1_mu2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean()


def calculate_ssim(img1, img2):
    """img_utils.py:217-234: for HxWx3 the reference averages three identical ssim(img1, img2) calls."""
    if img1.shape != img2.shape:
        raise ValueError("Input images must have the same dimensions.")
    if img1.ndim == 3 and img1.shape[2] == 1:
        return ssim(np.squeeze(img1), np.squeeze(img2))
    return ssim(img1, img2)
