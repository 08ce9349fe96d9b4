"""HAT-L expert on the ffb200 kernels vs the fp32 oracle (oracle/hat.py, pinned to the reference).
Tolerance (BASELINE.json north_star): max-abs <= 2e-2 on [0,1] outputs with bf16 GEMM operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _run(depths, blocks, B, H, W, seed=0):
    from isr2_b200 import hat, ops, weights
    from oracle import hat as ohat
    sd = weights.make_state_dict("hat", seed)
    g = torch.Generator().manual_seed(100 + seed)
    x = torch.rand(B, 3, H, W, generator=g)
    with torch.no_grad():
        ref = ohat.forward_hat(sd, x, depths, blocks)
    dev = torch.device("cuda:0")
    runner = hat.HATRunner(sd, dev, depths, blocks)
    stack = torch.zeros(B * 16 * H * W, 12, device=dev)
    runner.forward(x.to(dev), stack, out_off=0)
    got = torch.zeros(B, 3, 4 * H, 4 * W, device=dev)
    ops.nhwc_to_nchw(stack, 0, 3, got)
    torch.cuda.synchronize()
    return (got.cpu() - ref).abs().max().item(), ref


def test_hat_one_group_two_blocks():
    err, ref = _run(1, 2, 1, 32, 32)
    assert err < TOL, f"max abs err {err}"


def test_hat_two_groups_full_blocks_rect():
    err, ref = _run(2, 6, 2, 32, 48)
    assert err < TOL, f"max abs err {err}"


def test_hat_l_full_64():
    err, ref = _run(12, 6, 2, 64, 64)
    assert ref.std() > 0.02
    assert err < TOL, f"max abs err {err}"
