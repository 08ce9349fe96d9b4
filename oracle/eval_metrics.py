"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (numpy + scipy) of the PSNR / SSIM the reference's evaluation harness reports: `cal_psnr_ssim`
(utils/utils_image.py:287-312, called per image from eval.py:157), i.e.

    uint8 RGB images -> crop `border` pixels -> Y = cv2.cvtColor(img, COLOR_RGB2YCrCb)[..., 0] -> float64
    PSNR = 10 log10(255^2 / mean((Ya - Yb)^2))                      (inf when identical)
    SSIM = skimage.metrics.structural_similarity(Ya, Yb, data_range=255.0)

PARITY UNPINNED: both arithmetic steps live in third-party packages that are absent from /root/reference and from this image
(requirements.txt:30 `opencv-python>=4.7.0`, :38 `scikit-image>=0.20.0`; commented pin opencv-python==4.8.0.74 at :124), so
this file restates their published algorithms and cannot be checked against the reference run here:

* OpenCV `cvtColor(..., COLOR_RGB2YCrCb)` on 8-bit data (modules/imgproc/src/color_yuv: `RGB2YCrCb_i<uchar>`): fixed point with
  yuv_shift = 14, Y = (4899 R + 9617 G + 1868 B + 2^13) >> 14  (the BT.601 weights 0.299 / 0.587 / 0.114 times 2^14; they sum to
  2^14, so Y stays in [0, 255] without saturation).
* scikit-image `structural_similarity` with its defaults (skimage/metrics/_structural_similarity.py, unchanged from 0.19 on):
  win_size 7, uniform filter (scipy.ndimage.uniform_filter, centred), K1 = 0.01, K2 = 0.03, sample covariance
  (cov_norm = 49 / 48), float64, S = ((2 ux uy + C1)(2 vxy + C2)) / ((ux^2 + uy^2 + C1)(vx + vy + C2)), and the mean of S over the
  image with a (win_size - 1) / 2 = 3 pixel border removed (only windows that lie inside the image count).

tests/test_cpu_oracle.py checks the properties the algorithms guarantee (identical images, a hand-computed 7x7 case, the
integer luma against the real-valued BT.601 formula); the GPU kernel ff_eval_psnr_ssim_u8 is held to this file.
"""
import math

import numpy as np
from scipy.ndimage import uniform_filter


def luma_u8(img):
    """uint8 [H, W, 3] RGB -> int32 [H, W]: OpenCV's 8-bit RGB2YCrCb luma (fixed point, shift 14, round half up)."""
    x = img.astype(np.int64)
    return ((4899 * x[..., 0] + 9617 * x[..., 1] + 1868 * x[..., 2] + (1 << 13)) >> 14).astype(np.int32)


def structural_similarity_u(x, y, data_range=255.0, win_size=7, k1=0.01, k2=0.03):
    """scikit-image's structural_similarity with default arguments on 2-D float64 arrays."""
    x = x.astype(np.float64)
    y = y.astype(np.float64)
    npx = win_size * win_size
    cov_norm = npx / (npx - 1.0)
    ux, uy = uniform_filter(x, size=win_size), uniform_filter(y, size=win_size)
    uxx, uyy, uxy = uniform_filter(x * x, size=win_size), uniform_filter(y * y, size=win_size), uniform_filter(x * y, size=win_size)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    pad = (win_size - 1) // 2
    return float(s[pad:-pad, pad:-pad].mean(dtype=np.float64))


def cal_psnr_ssim(output_img, target_img, border=4):
    """utils_image.py:287-312 on two uint8 [H, W, 3] RGB arrays (what `imread_uint(path, 3)` returns): (psnr, ssim)."""
    if border > 0:
        output_img = output_img[border:-border, border:-border, :]
        target_img = target_img[border:-border, border:-border, :]
    ya = luma_u8(output_img).astype(np.float64)
    yb = luma_u8(target_img).astype(np.float64)
    mse = np.mean((ya - yb) ** 2)
    psnr = float("inf") if mse == 0 else 10 * math.log10(255.0 ** 2 / mse)
    return psnr, structural_similarity_u(ya, yb, data_range=255.0)
