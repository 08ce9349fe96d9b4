"""ff_pack_taps (3-channel 3x3 im2col, two terms) at output resolution (development helper): python tools/bench_pack.py 16 512"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import ops
dev = torch.device("cuda:0")
B, S = int(sys.argv[1]), int(sys.argv[2])
P = B * S * S
x = torch.rand(P, 4, device=dev)
im = torch.empty(P, 64, device=dev, dtype=torch.bfloat16)
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
t = timed(lambda: ops.pack_taps(x, B, S, S, 3, 3, 2, im))
print(f"pack_taps 3x3x3 -> 64 bf16, {B} x {S}x{S}: {t:7.1f} us  ({P * (16 + 128) / t / 1e3:.0f} GB/s compulsory)")
