"""Golden vector of the reference's cached-training inference path WITH expert features
(`forward_with_precomputed(lr, expert_outputs, expert_features)` -> EnhancedCollaborativeWithLKA, enhanced_fusion.py:466-496,
large_kernel_attention.py:251-419): tests/golden/head_collab_64.pt, from the UNMODIFIED reference (build container only).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_collab.py
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isr2_b200 import weights  # noqa: E402
from oracle import collab, refshim  # noqa: E402
from oracle.make_golden import GOLD, SEED, lr_image  # noqa: E402


def main():
    torch.manual_seed(0)
    ens, model, ffio = refshim.build_reference()
    ms = model.state_dict()
    ms.update(weights.make_state_dict("fusion", SEED))
    model.load_state_dict(ms, strict=False)
    model.eval()
    lr = lr_image(1, 64, 64, 107)
    g = torch.Generator().manual_seed(108)
    up = F.interpolate(lr, scale_factor=4, mode="bicubic", align_corners=False)
    ex = {k: (up + s * torch.randn(1, 3, 256, 256, generator=g)).clamp(0, 1) for k, s in (("hat", 0.01), ("dat", 0.02), ("nafnet", 0.03))}
    feats = collab.synth_features(1, 64, 64, 109)
    with torch.no_grad():
        out, inter = model.forward_with_precomputed(lr, ex, feats, return_intermediates=True)
        plain = model.forward_with_precomputed(lr, ex)
    enh = inter["enhanced_outputs"]
    mods = {k: ((enh[k] / ex[k].clamp(min=1e-3))[ex[k] > 0.2]).mean().item() for k in ex}
    torch.save({"lr": lr, "expert_seed": 108, "feature_seed": 109, "out": out, "out_without_features": plain,
                "enhanced": torch.cat([enh[k] for k in ("hat", "dat", "nafnet")], 1).half()}, os.path.join(GOLD, "head_collab_64.pt"))
    print("head_collab_64.pt", os.path.getsize(os.path.join(GOLD, "head_collab_64.pt")), "mean modulation", mods,
          "effect on the output", (out - plain).abs().max().item())
    # second golden: the feature-handling paths of EnhancedCollaborativeWithLKA.forward (large_kernel_attention.py:337-378) --
    # too many channels (truncated), too few (zero padded), a larger spatial size (aligned map resized to the smallest)
    lr2 = lr_image(1, 32, 32, 117)
    g2 = torch.Generator().manual_seed(118)
    up2 = F.interpolate(lr2, scale_factor=4, mode="bicubic", align_corners=False)
    ex2 = {k: (up2 + s * torch.randn(1, 3, 128, 128, generator=g2)).clamp(0, 1) for k, s in (("hat", 0.01), ("dat", 0.02), ("nafnet", 0.03))}
    feats2 = collab.synth_features_mixed(1, 32, 32, 119)
    with torch.no_grad():
        out2, inter2 = model.forward_with_precomputed(lr2, ex2, feats2, return_intermediates=True)
    enh2 = inter2["enhanced_outputs"]
    torch.save({"lr": lr2, "expert_seed": 118, "feature_seed": 119, "out": out2,
                "enhanced": torch.cat([enh2[k] for k in ("hat", "dat", "nafnet")], 1).half()}, os.path.join(GOLD, "head_collab_mixed_32.pt"))
    print("head_collab_mixed_32.pt", os.path.getsize(os.path.join(GOLD, "head_collab_mixed_32.pt")))


if __name__ == "__main__":
    main()
