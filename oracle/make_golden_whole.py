"""Golden vector of the reference's WHOLE-IMAGE forward on a size nothing divides (tests/golden/full_odd_37x50.pt):
runs the UNMODIFIED reference (/root/reference, build container only) with the seeded synthetic weights, storing the final
output, the three expert outputs (fp16 to keep the fixture small) and the fp32 frequency bands.  37 x 50 exercises every
padding path of io.py:219-221 at once: reflect padding to 48 x 64 for HAT / DAT (expert_loader.py:63-91), DAT's zero padding of
q / k / v to 64 x 64 with run-time masks (dat_arch.py:505-528), NAFNet's zero padding of the 148 x 200 bicubic image to
160 x 208 (nafnet_arch.py:219-225), the DCT's reflect padding to 40 x 56, odd-length DWT and rfft2 axes, and the floor-sized
1/2 and 1/4 pyramids of fusion_network.py:594,599.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_whole.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isr2_b200 import weights  # noqa: E402
from oracle import refshim  # noqa: E402
from oracle.make_golden import GOLD, SEED, lr_image  # noqa: E402


def main():
    torch.manual_seed(0)
    ens, model, ffio = refshim.build_reference()
    state = {m: weights.make_state_dict(m, SEED) for m in ("hat", "dat", "nafnet", "fusion")}
    ens.hat.load_state_dict(state["hat"], strict=False)
    ens.dat.load_state_dict(state["dat"], strict=False)
    assert ens.nafnet.load_nafnet_weights(state["nafnet"])["skipped"] == 0
    ms = model.state_dict()
    ms.update(state["fusion"])
    model.load_state_dict(ms, strict=False)
    model.eval()
    x = lr_image(1, 37, 50, 106)
    with torch.no_grad():
        out, inter = model(x, return_intermediates=True)
        raw = model.multi_domain_freq.decompose(x)
    ex = inter["expert_outputs"]
    torch.save({"x": x, "out": out, "hat": ex["hat"].half(), "dat": ex["dat"].half(), "nafnet": ex["nafnet"].half(),
                "raw_bands": torch.cat(list(raw.values()) if isinstance(raw, dict) else list(raw), 1)}, os.path.join(GOLD, "full_odd_37x50.pt"))
    print("full_odd_37x50.pt", os.path.getsize(os.path.join(GOLD, "full_odd_37x50.pt")), tuple(out.shape))


if __name__ == "__main__":
    main()
