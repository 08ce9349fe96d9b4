"""Whole-image (un-tiled) forward for arbitrary H x W -- the reference's primary path in io.main (io.py:218-221) -- on the
CUDA kernels: the padding paths (reflect to the 16-px window for HAT / DAT, DAT's zero padding to 32 with run-time masks,
NAFNet's zero padding to 16, the DCT's reflect padding, DFTs of any length, floor-sized pyramids) against golden outputs of the
reference itself and against the oracle.  Tolerances as everywhere: max-abs <= 2e-2 on [0,1] outputs (bf16 operands), 1e-4 for
the fp32 frequency bands."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda:0"
TOL = 2e-2


@pytest.fixture(autouse=True)
def _strict_fp32():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _lr(B, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(B, 3, max(h // 4, 1), max(w // 4, 1), generator=g)
    x = F.interpolate(low, size=(h, w), mode="bicubic", align_corners=False) + 0.03 * torch.randn(B, 3, h, w, generator=g)
    return (x.clamp(0, 1) * 255).round() / 255


def _to(sd, dev):
    return {k: v.to(dev) for k, v in sd.items()}


@pytest.fixture(scope="module")
def model():
    from isr2_b200 import model as M
    return M.FreqFusionB200(DEV, init_seed=0, verbose=False)


def test_experts_vs_reference_golden_unaligned_24x40(model):
    """tests/golden/experts_24x40.pt: outputs of the reference's forward_hat / forward_dat / forward_nafnet on a 24x40 image
    (reflect padding to 32x48, DAT window padding to 32x64, NAFNet padding 96x160 -> 96x160 is already aligned)."""
    g = torch.load(os.path.join(GOLD, "experts_24x40.pt"))
    ex = model.expert_outputs_nchw(g["x"].to(DEV))
    for k in ("hat", "dat", "nafnet"):
        err = (ex[k].cpu() - g[k]).abs().max().item()
        print(f"experts 24x40 {k}: {err:.2e}")
        assert err < TOL, f"{k}: {err}"


def test_full_model_vs_reference_golden_32(model):
    g = torch.load(os.path.join(GOLD, "full_32.pt"))
    out = model.forward(g["x"].to(DEV)).cpu()
    err = (out - g["out"]).abs().max().item()
    print(f"full 32x32 vs reference golden: {err:.2e}")
    assert err < TOL


def test_full_model_vs_reference_golden_odd_37x50(model):
    """Every padding path at once (oracle/make_golden_whole.py), against the reference's own output."""
    g = torch.load(os.path.join(GOLD, "full_odd_37x50.pt"))
    inter = {}
    out = model.forward(g["x"].to(DEV), intermediates=inter).cpu()
    ex = model.expert_outputs_nchw(g["x"].to(DEV))
    for k in ("hat", "dat", "nafnet"):
        err = (ex[k].cpu() - g[k].float()).abs().max().item()
        print(f"odd 37x50 {k}: {err:.2e}")
        assert err < TOL, f"{k}: {err}"
    raw = inter["bands_raw"].cpu().view(1, 37, 50, 27).permute(0, 3, 1, 2)
    e_raw = (raw - g["raw_bands"]).abs().max().item()
    err = (out - g["out"]).abs().max().item()
    print(f"odd 37x50: raw bands {e_raw:.2e}, output {err:.2e}")
    assert e_raw < 1e-4 and err < TOL
    # replayed CUDA graph of the same shape gives the same bits
    assert torch.equal(model.forward(g["x"].to(DEV)).cpu(), model.forward(g["x"].to(DEV)).cpu())


@pytest.mark.parametrize("B,h,w", [(1, 75, 101), (2, 90, 56), (1, 339, 510)])
def test_full_model_whole_image_vs_oracle(model, B, h, w):
    """Odd sizes, a batch of equal-size images, and a DIV2K-shaped 339x510 image run WHOLE (22 x 32 HAT windows, DAT padded to
    352 x 512, NAFNet on 1360 x 2048) against the oracle in strict fp32 on this GPU."""
    from oracle import full
    lr = _lr(B, h, w, 900 + h)
    with torch.no_grad():
        ref = full.forward({k: _to(v, DEV) for k, v in model.state.items()}, lr.to(DEV)).cpu()
    out = model.forward(lr.to(DEV)).cpu()
    err = (out - ref).abs().max().item()
    print(f"whole image B={B} {h}x{w}: max-abs {err:.2e}")
    assert out.shape == (B, 3, 4 * h, 4 * w) and err < TOL


def test_too_small_images_fail_like_the_reference(model):
    from isr2_b200 import lib
    with pytest.raises(lib.FFError):
        model.forward(torch.rand(1, 3, 8, 40, device=DEV))


def test_io_main_whole_image_first(tmp_path):
    """main() runs images whole (io.py:218-221), tiles only above the size threshold: an odd-sized PNG must match the oracle's
    WHOLE-image output, not its tiled one."""
    from PIL import Image
    from isr2_b200 import io as ffio, weights
    from oracle import full, tiling as otil
    root = str(tmp_path)
    fusion = weights.save_checkpoints(root, seed=6)
    os.environ["FFB200_PRETRAINED_ROOT"] = root
    inp, outp = os.path.join(root, "in"), os.path.join(root, "out")
    os.makedirs(inp)
    imgs = {"odd.png": _lr(1, 70, 93, 71), "tiny.png": _lr(1, 20, 33, 72)}
    for n, t in imgs.items():
        Image.fromarray((t[0].permute(1, 2, 0).numpy() * 255).round().astype("uint8")).save(os.path.join(inp, n))
    ffio.main(model_dir=fusion, input_path=inp, output_path=outp, device=torch.device(DEV))
    state = {m: _to(weights.make_state_dict(m, 6), DEV) for m in ("hat", "dat", "nafnet", "fusion")}
    for n, t in imgs.items():
        got = np.array(Image.open(os.path.join(outp, n)))
        with torch.no_grad():
            ref = otil.to_uint8(full.forward(state, t.to(DEV)).cpu())
        diff = np.abs(got.astype(int) - ref.astype(int))
        print(f"main whole-image {n}: max |diff| {diff.max()} gray levels")
        assert got.shape == ref.shape and diff.max() <= 5
