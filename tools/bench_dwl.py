"""Times the large-kernel depthwise chain of the fusion head at the bench shape (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import ops, packing
B, S, C_ = int(sys.argv[1]), int(sys.argv[2]), 576
d = torch.device("cuda:0")
x = torch.randn(B * S * S, C_, device=d).to(torch.bfloat16)
out = torch.empty_like(x)
for kh, kw in ((5, 5), (1, 21), (21, 1)):
    w = packing.pack_dw(torch.randn(C_, 1, kh, kw), C_, device=d)
    for _ in range(3):
        ops.dwconv(x, B, S, S, C_, kh, kw, w, None, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.dwconv(x, B, S, S, C_, kh, kw, w, None, out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"dw {kh}x{kw} C={C_} P={B*S*S}: {us:.1f} us  {2*x.numel()*2/us/1e3:.0f} GB/s")
