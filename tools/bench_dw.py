import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import ops
dev = torch.device("cuda:0"); BF16 = torch.bfloat16
B, S = int(sys.argv[1]), int(sys.argv[2])
M = B * S * S
def run(name, fn, byts):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"{name}: {us:8.1f} us  {byts/us/1e3:7.1f} GB/s (compulsory)")
t2 = torch.randn(M, 384, device=dev).to(BF16); h = torch.randn(M, 768, device=dev).to(BF16); gt = torch.empty(M, 384, device=dev, dtype=BF16)
w = torch.randn(9, 384, device=dev); b = torch.randn(384, device=dev)
run("DAT sgfn dw 384 + mul", lambda: ops.dwconv(t2, B, S, S, 384, 3, 3, w, b, gt, mul=h), M * 384 * 2 * 3)
qkv = torch.randn(M, 576, device=dev).to(BF16); cx = torch.empty(M, 192, device=dev, dtype=BF16)
w2 = torch.randn(9, 192, device=dev); b2 = torch.randn(192, device=dev)
run("DAT v dw 192 gelu     ", lambda: ops.dwconv(qkv, B, S, S, 192, 3, 3, w2, b2, cx, act=1, x_off=384), M * 192 * 2 * 2)
a = torch.randn(B * 16 * S * S, 128, device=dev).to(BF16); g = torch.empty(B * 16 * S * S, 64, device=dev, dtype=BF16)
w3 = torch.randn(9, 128, device=dev); b3 = torch.randn(128, device=dev)
run("NAF gate dw 128 @HR    ", lambda: ops.dwconv(a, B, 4 * S, 4 * S, 128, 3, 3, w3, b3, g, mode=1), B * 16 * S * S * (128 + 64) * 2)
x = torch.randn(M, 192, device=dev); gam = torch.randn(180, device=dev); bet = torch.randn(180, device=dev); t = torch.empty(M, 192, device=dev, dtype=BF16)
run("LN 180 fp32->bf16      ", lambda: ops.layernorm(x, M, 180, gam, bet, 1e-5, out_bf16=t, out_cols=192), M * (192 * 4 + 192 * 2))
