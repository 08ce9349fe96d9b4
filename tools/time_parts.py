"""Device time of each expert and the head at one shape (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200.model import FreqFusionB200
B, S = int(sys.argv[1]), int(sys.argv[2])
m = FreqFusionB200("cuda:0", verbose=False)
r = m.runners()
x = torch.rand(B, 3, S, S, device="cuda:0")
stack = m._stack(B, S, S)
def t(fn, n=3):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
parts = {"hat": lambda: r["hat"].forward(x, stack, 0), "dat": lambda: r["dat"].forward(x, stack, 3), "nafnet": lambda: r["nafnet"].forward(x, stack, 6),
         "head": lambda: r["head"].forward(x, stack)}
tot = 0
for k, f in parts.items():
    ms = t(f); tot += ms
    print(f"{k:8s} {ms:8.2f} ms")
print(f"sum      {tot:8.2f} ms  -> {B*16*S*S/1e3/tot:.2f} Mpix/s")
