"""The edge refiner's small layers at output resolution through ff_conv_direct (development helper): python tools/bench_small_convs.py 16 512"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import ops, packing
dev = torch.device("cuda:0")
B, S = int(sys.argv[1]), int(sys.argv[2])
P = B * S * S
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
cases = [("a2 8->1 3x3", 8, 1, 3, 8, 1, False, False, ops.ACT_SIGMOID), ("egate2 16->1 3x3", 16, 1, 3, 16, 1, False, False, ops.ACT_SIGMOID),
         ("egate0 6->16 3x3", 6, 16, 3, 8, 16, False, False, ops.ACT_GELU), ("a0 32->8 1x1", 32, 8, 1, 64, 8, True, False, ops.ACT_GELU),
         ("pj 3->64 1x1", 3, 64, 1, 4, 64, False, True, ops.ACT_NONE), ("ms_mix 64->64 1x1 fp32", 64, 64, 1, 64, 64, False, False, ops.ACT_NONE)]
for name, cin, cout, k, ild, old, ibf, obf, act in cases:
    x = torch.randn(P, ild, device=dev).to(torch.bfloat16 if ibf else torch.float32)
    cp = (cout + 7) // 8 * 8
    w = packing.pack_conv_direct(torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5, cp, dev)
    b = torch.zeros(cp, device=dev)
    out = torch.empty(P, old, device=dev, dtype=torch.bfloat16 if obf else torch.float32)
    kw = dict(out_bf16=out) if obf else dict(out_f32=out)
    t = timed(lambda: ops.conv_direct(x, B, S, S, cin, k, w, b, n_store=cout, act=act, **kw))
    byts = P * (cin * (2 if ibf else 4) + cout * (2 if obf else 4))
    print(f"{name:20s} {t:8.1f} us   {byts / t / 1e3:7.0f} GB/s compulsory")
