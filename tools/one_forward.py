"""Runs N forwards of one expert/head (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import weights

name, B, S, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
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
for _ in range(n):
    r.forward(x, stack)
torch.cuda.synchronize()
print("ok")
