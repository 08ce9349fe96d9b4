"""Runs N forwards of ONE part of the model (hat | dat | nafnet | head) at a shape; the LAST forward sits between cudaProfilerStart /
cudaProfilerStop, so `ncu --profile-from-start off` lists exactly that part's kernels in steady state (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200.model import FreqFusionB200
part, B, S, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
m = FreqFusionB200("cuda:0", verbose=False)
r = m.runners()
x = torch.rand(B, 3, S, S, device="cuda:0")
stack = m._stack(B, S, S)
stack.uniform_(0, 1)
off = {"hat": 0, "dat": 3, "nafnet": 6}
for i in range(n):
    if i == n - 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    if part == "head":
        r["head"].forward(x, stack)
    else:
        r[part].forward(x, stack, off[part])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
