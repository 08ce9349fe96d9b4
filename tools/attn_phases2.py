"""Phase timers of the persistent two-group attention variant (development helper, -DFF_ATTN_PROF build)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import ops, lib
L = lib.load()
dev = torch.device("cuda:0")
B, S = 16, 128
M = B * S * S
qkv = (torch.randn(M, 576, device=dev) * 0.3).to(torch.bfloat16)
out = torch.empty(M, 192, device=dev, dtype=torch.bfloat16)
buf = (C.c_ulonglong * 8)()
names = ["item top (wait/gather issue)", "S issue+wait", "pass 1 + sync", "pass 2", "sync", "PV issue+wait", "read-out + sync (+prefetch)"]
kw = dict(bias_table=torch.randn(6, 961, device=dev), wh=16, ww=16)
ops.window_attention(qkv, B, S, S, out, **kw)
L.ff_debug_attn_prof(buf, 1)
ops.window_attention(qkv, B, S, S, out, **kw)
L.ff_debug_attn_prof(buf, 1)
n = buf[7]
tot = sum(buf[i] for i in range(7))
tiles = B * (S // 16) ** 2 * 3 * 4
print(f"{n} groups, {tot / n:.0f} cycles per group, {tiles / n:.1f} tiles per group")
for i, nm in enumerate(names):
    print(f"  {nm:30s} {buf[i] / tiles:8.0f} cycles per tile  ({100 * buf[i] / tot:.1f} %)")
