"""Quick device timing of one expert (development helper, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import weights, lib

name = sys.argv[1]
B, S = int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
sd = weights.make_state_dict(name, 0)
if name == "hat":
    from isr2_b200 import hat
    r = hat.HATRunner(sd, dev)
elif name == "dat":
    from isr2_b200 import dat
    r = dat.DATRunner(sd, dev)
else:
    from isr2_b200 import nafnet
    r = nafnet.NAFNetRunner(sd, dev)
x = torch.rand(B, 3, S, S, device=dev)
stack = torch.zeros(B * 16 * S * S, 12, device=dev)
for _ in range(2):
    r.forward(x, stack)
torch.cuda.synchronize()
n0 = lib.launch_count()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
iters = 3
for _ in range(iters):
    r.forward(x, stack)
e1.record()
torch.cuda.synchronize()
host = (time.perf_counter() - t0) / iters
ms = e0.elapsed_time(e1) / iters
print(f"{name} B={B} S={S}: {ms:.2f} ms/forward (device), host wall {host*1e3:.2f} ms, launches/forward {(lib.launch_count()-n0)//iters}, {B*16*S*S/ms/1e3:.2f} Mpix/s")
# CUDA graph
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    r.forward(x, stack)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        r.forward(x, stack)
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
e0.record()
for _ in range(iters):
    g.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"{name} B={B} S={S}: {ms:.2f} ms/forward (CUDA graph), {B*16*S*S/ms/1e3:.2f} Mpix/s")
