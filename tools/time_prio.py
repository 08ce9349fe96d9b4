"""Device time of the full forward at one shape for the side-stream priorities in FFB200_STREAM_PRIO (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import model as M
B, S = int(sys.argv[1]), int(sys.argv[2])
M.GRAPH_MAX_LR_PIXELS = 0
m = M.FreqFusionB200("cuda:0", verbose=False)
x = torch.rand(B, 3, S, S, device="cuda:0")
out = torch.empty(B, 3, 4 * S, 4 * S, device="cuda:0")
def t(n=6):
    for _ in range(3): m.forward(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): m.forward(x, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
res = {}
for rep in range(2):
    for prio in sys.argv[3:]:
        os.environ["FFB200_STREAM_PRIO"] = prio
        if hasattr(m, "_streams"): del m._streams
        res.setdefault(prio, []).append(t())
for k, v in res.items():
    print(f"B={B} S={S} prio(dat,nafnet)={k}: " + " ".join(f"{a:.2f}" for a in v) + " ms")
