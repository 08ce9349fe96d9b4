"""Generates tests/golden/metrics.pt with the UNMODIFIED reference's src/utils/metrics.py (build container only):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_metrics.py

Seeded image pairs and the reference's calculate_psnr / calculate_ssim (crop_border=4, test_y_channel=True; scikit-image is
not installed here, so calculate_ssim takes its PyTorch branch).  Pins oracle/metrics.py and, through it, ff_psnr_y / ff_ssim_y.
"""
import importlib.util
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FF_REFERENCE_ROOT", "/root/reference")


def main():
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(REF, "src", "utils", "metrics.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert not m.SKIMAGE_AVAILABLE
    g = torch.Generator().manual_seed(11)
    cases = []
    for (h, w, noise) in ((40, 52, 0.02), (33, 71, 0.08), (64, 64, 0.005)):
        low = torch.rand(1, 3, h // 4, w // 4, generator=g)
        a = F.interpolate(low, size=(h, w), mode="bicubic", align_corners=False)      # overshoots [0,1]: exercises the clamp
        b = a + noise * torch.randn(1, 3, h, w, generator=g)
        cases.append({"a": a, "b": b,
                      "psnr": m.calculate_psnr(a, b, crop_border=4, test_y_channel=True),
                      "ssim": m.calculate_ssim(a, b, crop_border=4, test_y_channel=True)})
    torch.save(cases, os.path.join(ROOT, "tests", "golden", "metrics.pt"))
    for c in cases:
        print(tuple(c["a"].shape), c["psnr"], c["ssim"])


if __name__ == "__main__":
    main()
