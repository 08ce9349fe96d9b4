"""DAT and NAFNet-SR experts on the ffb200 kernels vs the fp32 oracles (pinned to the reference).
Tolerance: max-abs <= 2e-2 on [0,1] outputs (bf16 GEMM operands, fp32 accumulation / residual stream)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _img(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(B, 3, H // 4, W // 4, generator=g)
    x = torch.nn.functional.interpolate(low, scale_factor=4, mode="bicubic", align_corners=False) + 0.03 * torch.randn(B, 3, H, W, generator=g)
    return (x.clamp(0, 1) * 255).round() / 255


def _run_dat(groups, blocks, B, H, W):
    from isr2_b200 import dat, ops, weights
    from oracle import dat as odat
    sd = weights.make_state_dict("dat", 0)
    x = _img(B, H, W, 7)
    with torch.no_grad():
        ref = odat.forward_dat(sd, x, groups, blocks)
    dev = torch.device("cuda:0")
    r = dat.DATRunner(sd, dev, groups, blocks)
    stack = torch.zeros(B * 16 * H * W, 12, device=dev)
    r.forward(x.to(dev), stack, out_off=3)
    got = torch.zeros(B, 3, 4 * H, 4 * W, device=dev)
    ops.nhwc_to_nchw(stack, 3, 3, got)
    torch.cuda.synchronize()
    return (got.cpu() - ref).abs().max().item()


def test_dat_two_groups():
    # group 0 block 2 and group 1 blocks 0,4 are the shifted ones; blocks 1,3,5 are channel attention
    err = _run_dat(2, 6, 1, 64, 32)
    assert err < TOL, f"max abs err {err}"


def test_dat_full_64():
    err = _run_dat(6, 6, 2, 64, 64)
    assert err < TOL, f"max abs err {err}"


def _run_naf(B, h, w, **kw):
    from isr2_b200 import nafnet, ops, weights
    from oracle import nafnet as onaf
    sd = weights.make_state_dict("nafnet", 0)
    x = _img(B, h, w, 9)
    with torch.no_grad():
        ref = onaf.forward_nafnet(sd, x, **kw)
    dev = torch.device("cuda:0")
    r = nafnet.NAFNetRunner(sd, dev, **kw)
    stack = torch.zeros(B * 16 * h * w, 12, device=dev)
    r.forward(x.to(dev), stack, out_off=6)
    got = torch.zeros(B, 3, 4 * h, 4 * w, device=dev)
    ops.nhwc_to_nchw(stack, 6, 3, got)
    torch.cuda.synchronize()
    return (got.cpu() - ref).abs().max().item()


def test_nafnet_full_64():
    err = _run_naf(2, 64, 64)
    assert err < TOL, f"max abs err {err}"


def test_nafnet_rect_128x64():
    err = _run_naf(1, 128, 64)
    assert err < TOL, f"max abs err {err}"
