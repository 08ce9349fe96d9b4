"""Fused MLP kernel vs the two conv_gemm launches at the bench shape (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import ops, packing
dev = torch.device("cuda:0")
B, S = int(sys.argv[1]), int(sys.argv[2])
M = B * S * S
g = torch.Generator().manual_seed(0)
t = torch.randn(M, 192, generator=g).to(dev, torch.bfloat16)
x = torch.randn(M, 192, generator=g).to(dev)
h = torch.empty(M, 384, device=dev, dtype=torch.bfloat16)
w1 = packing.pack_matrix(torch.randn(360, 180, generator=g) / 13, 384, 192, device=dev)
w2 = packing.pack_matrix(torch.randn(180, 360, generator=g) / 19, 192, 384, device=dev)
b1, b2 = torch.zeros(384, device=dev), torch.zeros(192, device=dev)
gam, bet = torch.ones(192, device=dev), torch.zeros(192, device=dev)
lno = torch.empty(M, 192, device=dev, dtype=torch.bfloat16)
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
def two(ln):
    ops.conv_gemm(t, B, S, S, 192, w1, n_store=384, bias=b1, act=ops.ACT_GELU, out_bf16=h)
    ops.conv_gemm(h, B, S, S, 384, w2, n_store=192, bias=b2, res=x, out_f32=x, ln=(gam, bet, 1e-5, 180, lno) if ln else None)
for ln in (False, True):
    a = timed(lambda: two(ln))
    b = timed(lambda: ops.mlp_fused(t, B, S, S, w1, b1, w2, b2, x, ln=(gam, bet, 1e-5, 180, lno) if ln else None))
    flops = 2.0 * M * 2 * 180 * 360
    print(f"MLP B={B} S={S} ln={ln}: two kernels {a:7.1f} us   fused {b:7.1f} us  ({flops / b / 1e6:.0f} TFLOP/s algorithmic, {(M * 192 * (2 + 8) + (M * 192 * 2 if ln else 0)) / b / 1e3:.0f} GB/s compulsory)")
