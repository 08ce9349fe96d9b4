"""Phase timers of window_attention_tc4_kernel (development helper; needs a library built with FFB200_NVCC_EXTRA=-DFF_ATTN_PROF).
Prints average cycles per CTA (thread 0) for each phase, and the occupancy the runtime computes for the kernel."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import ops, lib
L = lib.load()
dev = torch.device("cuda:0")
B, S = int(sys.argv[1]), int(sys.argv[2])
M = B * S * S
qkv = (torch.randn(M, 576, device=dev) * 0.3).to(torch.bfloat16)
out = torch.empty(M, 192, device=dev, dtype=torch.bfloat16)
buf = (C.c_ulonglong * 12)()
names = ["setup: alloc, TMA / gather, table, sync", "wait S (mbarrier)", "pass 1 (max)", "O rescale", "pass 2 (exp)", "wait st", "barrier", "PV + next S issue", "read-out"]
for shift in ((0, 0), (8, 8)):
    kw = dict(bias_table=torch.randn(6, 961, device=dev), wh=16, ww=16, shift=shift)
    ops.window_attention(qkv, B, S, S, out, **kw)
    L.ff_debug_attn4_prof(buf, 1)
    ops.window_attention(qkv, B, S, S, out, **kw)
    occ = L.ff_debug_attn4_prof(buf, 1)
    n = buf[11]
    tot = sum(buf[i] for i in range(9))
    print(f"shift {shift}: {n} CTAs, {tot / n:.0f} cycles per CTA, occupancy {occ} CTAs / SM")
    for i, nm in enumerate(names):
        print(f"  {nm:42s} {buf[i] / n:8.0f} cycles per CTA  ({100 * buf[i] / tot:.1f} %)")
