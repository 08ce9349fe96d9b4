"""Micro-benchmark of ff_window_attention (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import ops
dev = torch.device("cuda:0")
B, S = int(sys.argv[1]), int(sys.argv[2])
M = B * S * S
qkv = (torch.randn(M, 576, device=dev) * 0.3).to(torch.bfloat16)
out = torch.empty(M, 192, device=dev, dtype=torch.bfloat16)
cases = [
    ("HAT W-MSA      ", dict(bias_table=torch.randn(6, 961, device=dev), wh=16, ww=16)),
    ("HAT SW-MSA     ", dict(bias_table=torch.randn(6, 961, device=dev), wh=16, ww=16, shift=(8, 8))),
    ("HAT OCAB       ", dict(bias_table=torch.randn(6, 1521, device=dev), wh=16, ww=16, kh=24, kw=24, kpad=(4, 4), rel_sign=-1, rel_off=(-7, -7), rel_stride=39)),
    ("DAT 8x32       ", dict(bias_table=torch.randn(3, 945, device=dev), wh=8, ww=32, heads=3)),
    ("DAT 32x8 shift ", dict(bias_table=torch.randn(3, 945, device=dev), wh=32, ww=8, heads=3, head_off=3, shift=(16, 4))),
]
for name, kw in cases:
    for _ in range(3):
        ops.window_attention(qkv, B, S, S, out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.window_attention(qkv, B, S, S, out, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    heads = kw.get("heads", 6)
    nk = kw.get("kh", kw["wh"]) * kw.get("kw", kw["ww"])
    flops = 2.0 * 2 * M * heads * nk * 30
    byts = M * heads * 32 * 2 * 4
    print(f"{name} B={B} S={S}: {us:8.1f} us  {flops/us/1e6:6.1f} TFLOP/s (algorithmic)  {byts/us/1e3:7.1f} GB/s")
