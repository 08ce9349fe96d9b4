"""Runs N forwards of the full model (for ncu launch lists).  The LAST forward sits between cudaProfilerStart / cudaProfilerStop, so
`ncu --profile-from-start off ...` captures exactly one steady-state step (every kernel of it, torch's included, and nothing of the
warm-up forwards whose workspace allocations launch fill kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200.model import FreqFusionB200
B, S, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
m = FreqFusionB200("cuda:0", verbose=False)
x = torch.rand(B, 3, S, S, device="cuda:0")
for i in range(n):
    if i == n - 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    m.forward(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
