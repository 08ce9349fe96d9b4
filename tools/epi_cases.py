"""Micro-benchmark of the ff_conv_gemm launches that still take the generic epilogue (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import ops

dev = torch.device("cuda:0")
BF16, F32 = torch.bfloat16, torch.float32
only = sys.argv[1] if len(sys.argv) > 1 else None
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
B = 16


def run(name, B_, H_, W_, cin, n_pad, kind, **kw):
    if only and only not in name:
        return
    Mm = B_ * H_ * W_
    taps = 9 if kind == 1 else 1
    x = torch.randn(Mm, cin, device=dev).to(BF16)
    w = (torch.randn(n_pad, taps * cin, device=dev) / (taps * cin) ** 0.5).to(BF16)
    bias = torch.randn(n_pad, device=dev)
    f = lambda: ops.conv_gemm(x, B_, H_, W_, cin, w, kind=kind, bias=bias, **kw)
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1) * 1e3 / iters:9.1f} us")


P = B * 512 * 512
stack = torch.zeros(P, 12, device=dev)
run("narrow 3x3 64->3 f32 ld12 @512", B, 512, 512, 64, 16, 1, n_store=3, out_f32=stack[:, 3:])
run("narrow 3x3 64->3 +res f32 @512", B, 512, 512, 64, 16, 1, n_store=3, res=torch.randn(P, 3, device=dev), out_f32=torch.empty(P, 3, device=dev))
run("narrow 3x3 64->3 sigmoid f32 @512", B, 512, 512, 64, 16, 1, n_store=3, act=ops.ACT_SIGMOID, out_f32=torch.empty(P, 4, device=dev))
r16 = torch.randn(P, 64, device=dev).to(BF16)
run("3x3 64->64 +res16 -> bf16 @512", B, 512, 512, 64, 64, 1, n_store=64, res=r16, out_bf16=torch.empty(P, 64, device=dev, dtype=BF16))
run("3x3 64->64 +res16 +aux -> bf16 @512", B, 512, 512, 64, 64, 1, n_store=64, res=r16, aux=r16, aux_alpha=0.5, out_bf16=torch.empty(P, 64, device=dev, dtype=BF16))
run("3x3 64->64 gelu -> bf16 @512 (STORE ref)", B, 512, 512, 64, 64, 1, n_store=64, act=ops.ACT_GELU, out_bf16=torch.empty(P, 64, device=dev, dtype=BF16))
P2 = B * 128 * 1152
m16 = torch.randn(P2, 64, device=dev).to(BF16)
run("1x1 64->64 sigmoid*mul+res16 @128x1152", B, 128, 1152, 64, 64, 0, n_store=64, act=ops.ACT_SIGMOID, mul=m16, res=m16, out_bf16=torch.empty(P2, 64, device=dev, dtype=BF16))
run("1x1 64->64 +res16 @128x1152", B, 128, 1152, 64, 64, 0, n_store=64, res=m16, out_bf16=torch.empty(P2, 64, device=dev, dtype=BF16))
run("1x1 128->64 +res16 @128x1152", B, 128, 1152, 128, 64, 0, n_store=64, res=m16, out_bf16=torch.empty(P2, 64, device=dev, dtype=BF16))
P3 = B * 256 * 256
run("3x3 64->256 ps2 -> bf16 @256", B, 256, 256, 64, 256, 1, n_store=256, pixel_shuffle=2, out_bf16=torch.empty(P3 * 4, 64, device=dev, dtype=BF16))
run("1x1 128->256 ps2 +res f32 @256", B, 256, 256, 128, 256, 0, n_store=256, pixel_shuffle=2, res=torch.randn(P3 * 4, 64, device=dev), out_f32=torch.empty(P3 * 4, 64, device=dev))
run("nar16 3x3 64->16 f32 ld16 @512", B, 512, 512, 64, 16, 1, n_store=16, out_f32=torch.empty(P, 16, device=dev))
run("nar16 3x3 64->16 bf16 ld16 @512", B, 512, 512, 64, 16, 1, n_store=16, out_bf16=torch.empty(P, 16, device=dev, dtype=BF16))
run("nar8 3x3 64->8 f32 ld8 @512", B, 512, 512, 64, 16, 1, n_store=8, out_f32=torch.empty(P, 8, device=dev))
run("nar32 3x3 64->32 bf16 STORE @512", B, 512, 512, 64, 32, 1, n_store=32, out_bf16=torch.empty(P, 32, device=dev, dtype=BF16))
run("x1 1x1 64->64 +res16+mul16 @128x1152", B, 128, 1152, 64, 64, 0, n_store=64, mul=m16, res=m16, out_bf16=torch.empty(P2, 64, device=dev, dtype=BF16))
run("x1 1x1 64->64 sigmoid+res16 @128x1152", B, 128, 1152, 64, 64, 0, n_store=64, act=ops.ACT_SIGMOID, res=m16, out_bf16=torch.empty(P2, 64, device=dev, dtype=BF16))
run("x1 1x1 64->64 alpha+res16 @128x1152", B, 128, 1152, 64, 64, 0, n_store=64, alpha=0.7, res=m16, out_bf16=torch.empty(P2, 64, device=dev, dtype=BF16))
run("x1 3x3 64->64 +res16 -> bf16 @512", B, 512, 512, 64, 64, 1, n_store=64, res=r16, out_bf16=torch.empty(P, 64, device=dev, dtype=BF16))
