import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200.model import FreqFusionB200
B, S, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
m = FreqFusionB200("cuda:0", verbose=False)
r = m.runners()
x = torch.rand(B, 3, S, S, device="cuda:0")
stack = m._stack(B, S, S)
stack.uniform_(0, 1)
for _ in range(n):
    r["head"].forward(x, stack)
torch.cuda.synchronize()
print("ok")
