import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200.model import FreqFusionB200
B, S = int(sys.argv[1]), int(sys.argv[2])
m = FreqFusionB200("cuda:0", verbose=False)
x = torch.rand(B, 3, S, S, device="cuda:0")
out = torch.empty(B, 3, 4 * S, 4 * S, device="cuda:0")
for _ in range(3): m.forward(x, out=out)
torch.cuda.synchronize()
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager  ", t(lambda: m.forward(x, out=out)))
ref = out.clone()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    m.forward(x, out=out); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        m.forward(x, out=out)
torch.cuda.synchronize()
out.zero_(); g.replay(); torch.cuda.synchronize()
print("graph max diff vs eager", (out - ref).abs().max().item())
print("graph  ", t(lambda: g.replay()))
