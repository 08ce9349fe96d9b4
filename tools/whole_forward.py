"""Runs N whole-image forwards of an h x w LR image (for ncu launch lists of the un-tiled path)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200.model import FreqFusionB200
h, w, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
m = FreqFusionB200("cuda:0", verbose=False)
x = torch.rand(1, 3, h, w, device="cuda:0")
for _ in range(n):
    m.forward(x)
torch.cuda.synchronize()
print("ok")
