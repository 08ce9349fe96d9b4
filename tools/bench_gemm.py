"""Micro-benchmark of ff_conv_gemm on the shapes that dominate the model (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import ops

dev = torch.device("cuda:0")
BF16, F32 = torch.bfloat16, torch.float32
Bn, S = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 64
only = sys.argv[3] if len(sys.argv) > 3 else None
M = Bn * S * S
cases = [
    # name, kind, cin, n_pad, act, res, aux, out32, out16
    ("qkv   1x1 192->576            ", 0, 192, 576, 0, False, False, False, True),
    ("proj  1x1 192->192 +res+aux   ", 0, 192, 192, 0, True, True, True, False),
    ("fc1   1x1 192->384 gelu       ", 0, 192, 384, 1, False, False, False, True),
    ("fc2   1x1 384->192 +res       ", 0, 384, 192, 0, True, False, True, False),
    ("cab1  3x3 192->64 gelu        ", 1, 192, 64, 1, False, False, False, True),
    ("cab2  3x3 64->192             ", 1, 64, 192, 0, False, False, False, True),
    ("rhag  3x3 192->192 +res       ", 1, 192, 192, 0, True, False, True, False),
    ("sgfn1 1x1 192->768 gelu       ", 0, 192, 768, 1, False, False, False, True),
    ("head  3x3 64->64 gelu (HRx16) ", 1, 64, 64, 1, False, False, False, True),
]
for name, kind, cin, n, act, res, aux, o32, o16 in cases:
    if only and only not in name:
        continue
    hr = "HRx16" in name
    B_, H_, W_ = Bn, (4 * S if hr else S), (4 * S if hr else S)
    Mm = B_ * H_ * W_
    taps = 9 if kind == 1 else 1
    x = torch.randn(Mm, cin, device=dev).to(BF16)
    w = (torch.randn(n, taps * cin, device=dev) / (taps * cin) ** 0.5).to(BF16)
    bias = torch.randn(n, device=dev)
    kw = dict(kind=kind, n_store=n, bias=bias, act=act)
    if res:
        kw["res"] = torch.randn(Mm, n, device=dev)
    if aux:
        kw["aux"] = torch.randn(Mm, n, device=dev).to(BF16)
        kw["aux_chan"] = torch.rand(B_, n, device=dev)
        kw["aux_alpha"] = 0.01
    if o32:
        kw["out_f32"] = torch.empty(Mm, n, device=dev)
    if o16:
        kw["out_bf16"] = torch.empty(Mm, n, device=dev, dtype=BF16)
    for _ in range(3):
        ops.conv_gemm(x, B_, H_, W_, cin, w, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 20
    e0.record()
    for _ in range(it):
        ops.conv_gemm(x, B_, H_, W_, cin, w, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / it
    flops = 2.0 * Mm * n * taps * cin
    byts = Mm * cin * 2 + (Mm * n * 4 if res else 0) + (Mm * n * 2 if aux else 0) + (Mm * n * 4 if o32 else 0) + (Mm * n * 2 if o16 else 0)
    print(f"{name} M={Mm:8d}: {us:8.1f} us  {flops/us/1e6:7.1f} TFLOP/s  {byts/us/1e3:7.1f} GB/s (compulsory)")
