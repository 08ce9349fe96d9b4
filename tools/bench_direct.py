"""Micro-benchmark of ff_conv_direct on the head's HR layers (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import ops, packing
dev = torch.device("cuda:0")
B, S = 16, 512
P = B * S * S
def run(name, cin, cout, k, out_bf16, x_bf16=False, act=1):
    x = torch.rand(P, cin + (1 if cin % 2 else 0), device=dev)
    if x_bf16: x = x.to(torch.bfloat16)
    w = packing.pack_conv_direct(torch.randn(cout, cin, k, k), (cout + 7) // 8 * 8, dev)
    b = torch.randn((cout + 7) // 8 * 8, device=dev)
    o = torch.empty(P, (cout + 7) // 8 * 8, device=dev, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    kw = dict(out_bf16=o) if out_bf16 else dict(out_f32=o)
    f = lambda: ops.conv_direct(x, B, S, S, cin, k, w, b, n_store=cout, act=act, **kw)
    f(); f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) * 200:9.1f} us")
run("3->64 3x3 gelu -> bf16", 3, 64, 3, True)
run("3->64 3x3 -> f32", 3, 64, 3, False)
run("3->64 1x1 -> bf16", 3, 64, 1, True, act=0)
run("6->16 3x3 gelu -> f32", 6, 16, 3, False)
run("16->1 3x3 sigmoid -> f32", 16, 1, 3, False, act=4)
run("32(bf16)->8 1x1 gelu -> f32", 32, 8, 1, False, x_bf16=True)
run("8->1 3x3 sigmoid -> f32", 8, 1, 3, False, act=4)
