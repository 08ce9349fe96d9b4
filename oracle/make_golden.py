"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference, build container only) on seeded
inputs with the seeded synthetic weights of isr2_b200.weights.  Re-run with:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The fixtures pin the oracle restatement (oracle/*.py) and, through it, the CUDA path.  The reference's own tests hold no
golden vectors for this path, so these outputs of the reference itself are the pin.
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isr2_b200 import weights  # noqa: E402
from oracle import refshim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SEED = 0


def lr_image(B, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(B, 3, max(h // 4, 1), max(w // 4, 1), generator=g)
    x = F.interpolate(low, size=(h, w), mode="bicubic", align_corners=False) + 0.03 * torch.randn(B, 3, h, w, generator=g)
    return (x.clamp(0, 1) * 255).round() / 255


def main():
    torch.manual_seed(0)
    ens, model, ffio = refshim.build_reference()
    state = {m: weights.make_state_dict(m, SEED) for m in ("hat", "dat", "nafnet", "fusion")}
    ens.hat.load_state_dict(state["hat"], strict=False)
    ens.dat.load_state_dict(state["dat"], strict=False)
    info = ens.nafnet.load_nafnet_weights(state["nafnet"])
    assert info["skipped"] == 0
    ms = model.state_dict()
    for k, v in state["fusion"].items():
        assert k in ms and ms[k].shape == v.shape, k
        ms[k] = v
    model.load_state_dict(ms, strict=False)
    model.eval()

    # state_dict layout contract (names / shapes / dtypes)
    man = {}
    for name, mod in (("hat", ens.hat), ("dat", ens.dat), ("nafnet", ens.nafnet.nafnet)):
        man[name] = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in mod.state_dict().items()}
    man["fusion"] = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items() if not k.startswith("expert_ensemble.")}
    json.dump(man, open(os.path.join(GOLD, "state_dict_manifest.json"), "w"), indent=0)

    with torch.no_grad():
        # experts on a small non-aligned image (exercises reflect padding to x16 and DAT's zero padding to x32)
        x = lr_image(1, 24, 40, 101)
        torch.save({"x": x, "hat": ens.forward_hat(x), "dat": ens.forward_dat(x), "nafnet": ens.forward_nafnet(x)}, os.path.join(GOLD, "experts_24x40.pt"))
        # experts on one aligned 64x64 tile (the GPU tests' shape), stored as fp16 to keep the fixture small
        x = lr_image(1, 64, 64, 102)
        torch.save({"x": x, "hat": ens.forward_hat(x).half(), "dat": ens.forward_dat(x).half(), "nafnet": ens.forward_nafnet(x).half()},
                   os.path.join(GOLD, "experts_64x64_fp16.pt"))
        # fusion head (BASELINE.json configs[0]): synthetic expert outputs on a 64x64 LR tile
        lr = lr_image(1, 64, 64, 103)
        g = torch.Generator().manual_seed(104)
        up = F.interpolate(lr, scale_factor=4, mode="bicubic", align_corners=False)
        ex = {k: (up + s * torch.randn(1, 3, 256, 256, generator=g)).clamp(0, 1) for k, s in (("hat", 0.01), ("dat", 0.02), ("nafnet", 0.03))}
        out = model.forward_with_precomputed(lr, ex)
        raw = model.multi_domain_freq.decompose(lr)
        bf, _ = model.process_frequency_bands(lr)
        fused = model.fuse_experts(lr, ex, bf)
        torch.save({"lr": lr, "expert_seed": 104, "out": out, "raw_bands": torch.cat(raw, 1), "band_features": torch.cat(bf, 1), "fused_before_refine": fused},
                   os.path.join(GOLD, "head_64.pt"))
        # full model on a 32x32 image (whole-image forward of the reference = its primary path)
        x = lr_image(1, 32, 32, 105)
        torch.save({"x": x, "out": model(x)}, os.path.join(GOLD, "full_32.pt"))
        # tiled forward index math with a nearest-upsample stand-in model
        fake = lambda t: F.interpolate(t, scale_factor=4, mode="nearest")
        til = {}
        for (h, w, ts, ov) in ((339, 510, 128, 32), (339, 510, 64, 8), (256, 300, 128, 32), (128, 128, 128, 32), (70, 90, 64, 8)):
            xi = lr_image(1, h, w, 200 + h)
            o = ffio._tiled_forward(fake, xi, ts, ov, 4, "cpu")
            step = ts - ov
            ys = list(range(0, max(h - ts + 1, 1), step))
            if ys[-1] + ts < h:
                ys.append(h - ts)
            xs = list(range(0, max(w - ts + 1, 1), step))
            if xs[-1] + ts < w:
                xs.append(w - ts)
            til[f"{h}x{w}_{ts}_{ov}"] = {"ys": ys, "xs": xs, "seed": 200 + h, "max_err_vs_nearest": (o - fake(xi)).abs().max().item(),
                                         "checksum": o.double().sum().item(), "out_small": o if h * w < 8000 else None}
        torch.save(til, os.path.join(GOLD, "tiling.pt"))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
