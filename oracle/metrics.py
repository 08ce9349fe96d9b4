"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement of the reference's image-quality metrics on the BT.601 Y channel (src/utils/metrics.py):
`rgb_to_y` (:30-52), `calculate_psnr` (:76-126) and `calculate_ssim` (:189-246) in its PyTorch branch
`calculate_ssim_torch` (:129-186) -- the branch the reference takes when scikit-image is not installed, as in this image.
Pinned by tests/golden/metrics.pt (values produced by the unmodified reference, oracle/make_golden_metrics.py).
"""
import math

import torch
import torch.nn.functional as F


def rgb_to_y(img):
    """metrics.py:30-52: Y = (65.481 R + 128.553 G + 24.966 B + 16) / 255 on [0,1] RGB, NCHW -> [B,1,H,W]."""
    r, g, b = img[:, 0:1], img[:, 1:2], img[:, 2:3]
    return (65.481 * r + 128.553 * g + 24.966 * b + 16.0) / 255.0


def _prep(a, b, crop):
    a, b = a.clamp(0, 1), b.clamp(0, 1)
    if crop > 0:
        a, b = a[:, :, crop:-crop, crop:-crop], b[:, :, crop:-crop, crop:-crop]
    return rgb_to_y(a), rgb_to_y(b)


def psnr_y(a, b, crop=4):
    """metrics.py:76-126 with test_y_channel=True, one value per sample."""
    ya, yb = _prep(a, b, crop)
    out = []
    for i in range(ya.shape[0]):
        mse = torch.mean((ya[i] - yb[i]) ** 2).item()
        out.append(float("inf") if mse < 1e-10 else 10 * math.log10(1.0 / mse))
    return out


def ssim_y(a, b, crop=4, window_size=11, sigma=1.5):
    """metrics.py:189-246 (crop, clamp, Y) + :129-186 (11x11 Gaussian window, sigma 1.5, zero padding, C1 = 0.01^2,
    C2 = 0.03^2, mean of the full map), one value per sample."""
    ya, yb = _prep(a, b, crop)
    gauss = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    gauss = gauss / gauss.sum()
    window = (gauss.unsqueeze(1) @ gauss.unsqueeze(0)).float()[None, None]
    pad = window_size // 2
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    mu1, mu2 = F.conv2d(ya, window, padding=pad), F.conv2d(yb, window, padding=pad)
    s11 = F.conv2d(ya * ya, window, padding=pad) - mu1 ** 2
    s22 = F.conv2d(yb * yb, window, padding=pad) - mu2 ** 2
    s12 = F.conv2d(ya * yb, window, padding=pad) - mu1 * mu2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 ** 2 + mu2 ** 2 + c1) * (s11 + s22 + c2))
    return [m[i].mean().item() for i in range(m.shape[0])]
