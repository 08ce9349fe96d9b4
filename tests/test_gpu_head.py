"""Fusion head on the ffb200 kernels vs the fp32 oracle (oracle/head.py, pinned to the reference).
fp32 stages (frequency bands) are held to 1e-4; the bf16 conv chains to the 2e-2 output tolerance."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _inputs(B, S, seed):
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(B, 3, S // 4, S // 4, generator=g)
    lr = (F.interpolate(low, scale_factor=4, mode="bicubic", align_corners=False) + 0.03 * torch.randn(B, 3, S, S, generator=g)).clamp(0, 1)
    lr = (lr * 255).round() / 255
    up = F.interpolate(lr, scale_factor=4, mode="bicubic", align_corners=False)
    ex = [(up + s * torch.randn(B, 3, 4 * S, 4 * S, generator=g)).clamp(0, 1) for s in (0.01, 0.02, 0.03)]
    return lr, ex


@pytest.mark.parametrize("B,S", [(1, 64), (2, 128)])
def test_head_vs_oracle(B, S):
    from isr2_b200 import head, weights
    from oracle import head as ohead
    sd = weights.make_state_dict("fusion", 0)
    lr, ex = _inputs(B, S, 11)
    with torch.no_grad():
        ref, inter = ohead.head_forward(sd, lr, ex, True)
        raw_ref = torch.cat(ohead.decompose(sd, lr), 1)
    dev = torch.device("cuda:0")
    r = head.HeadRunner(sd, dev)
    stack = torch.zeros(B * 16 * S * S, 12)
    stack[:, :9] = torch.cat(ex, 1).permute(0, 2, 3, 1).reshape(-1, 9)
    got_i = {}
    out = r.forward(lr.to(dev), stack.to(dev), intermediates=got_i)
    torch.cuda.synchronize()
    nh = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])
    e_raw = (got_i["bands_raw"].cpu() - nh(raw_ref)).abs().max().item()
    e_bf = (got_i["band_features"].cpu() - nh(torch.cat(inter["band_features"], 1))).abs().max().item()
    e_fused = (got_i["fused_before_refine"].cpu()[:, :3] - nh(inter["fused_before_refine"])).abs().max().item()
    e_out = (out.cpu() - ref).abs().max().item()
    print(f"head B={B} S={S}: raw bands {e_raw:.2e}  band_features {e_bf:.2e}  fused {e_fused:.2e}  out {e_out:.2e}")
    assert e_raw < 1e-4, e_raw          # fp32 DCT / DWT / DFT
    assert e_bf < 2e-2, e_bf            # through the bf16 cross-band attention + LKA
    assert e_fused < 2e-2, e_fused
    assert e_out < 2e-2, e_out
