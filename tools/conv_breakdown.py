"""Group the ff_conv_gemm launches of one full forward by shape: time, TFLOP/s, GB/s (development helper)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FFB200_EXPERT_STREAMS"] = "0"
import torch
from isr2_b200 import ops
from isr2_b200.model import FreqFusionB200
B, S = int(sys.argv[1]), int(sys.argv[2])
m = FreqFusionB200("cuda:0", verbose=False)
x = torch.rand(B, 3, S, S, device="cuda:0")
for _ in range(2): m.forward(x)
torch.cuda.synchronize()
ops.PROFILE = ops.KernelProfile()
m.forward(x)
torch.cuda.synchronize()
g = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for r in ops.PROFILE.records:
    e = g[r[5]]
    e[0] += 1; e[1] += r[0].elapsed_time(r[1]); e[2] += r[2]; e[3] += r[3]; e[4] += r[4]
tot = sum(e[1] for e in g.values())
print(f"total conv_gemm {tot:.2f} ms over {len(ops.PROFILE.records)} launches")
print("kind cin npad B H W act res aux mul gate ps f32 | n ms us/launch algoTF execTF GB/s")
for k, e in sorted(g.items(), key=lambda kv: -kv[1][1]):
    print(*k, "|", e[0], f"{e[1]:.2f} {1e3*e[1]/e[0]:.1f} {e[2]/e[1]/1e9:.0f} {e[3]/e[1]/1e9:.0f} {e[4]/e[1]/1e6:.0f}")
