"""Device time of the full forward at one shape, eager vs CUDA-graph replay (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import model as M
B, S = int(sys.argv[1]), int(sys.argv[2])
m = M.FreqFusionB200("cuda:0", verbose=False)
x = torch.rand(B, 3, S, S, device="cuda:0")
out = torch.empty(B, 3, 4 * S, 4 * S, device="cuda:0")
def t(n=5):
    for _ in range(3): m.forward(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): m.forward(x, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
M.GRAPH_MAX_LR_PIXELS = 0
a = t()
M.GRAPH_MAX_LR_PIXELS = B * S * S
b = t()
print(f"B={B} S={S}: eager {a:.2f} ms ({B*16*S*S/1e3/a:.2f} Mpix/s)   graph replay {b:.2f} ms ({B*16*S*S/1e3/b:.2f} Mpix/s)")
