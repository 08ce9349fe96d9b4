"""ff_hab_tail vs the residual GEMM + ff_mlp_fused pair at the bench shape (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import ops, packing
dev = torch.device("cuda:0")
B, S = int(sys.argv[1]), int(sys.argv[2])
M = B * S * S
g = torch.Generator().manual_seed(0)
BF = torch.bfloat16
att = torch.randn(M, 192, generator=g).to(dev, BF)
cab = torch.randn(M, 192, generator=g).to(dev, BF)
x = torch.randn(M, 192, generator=g).to(dev)
t = torch.empty(M, 192, device=dev, dtype=BF)
wp = packing.pack_matrix(torch.randn(180, 180, generator=g) / 13, 192, 192, device=dev)
w1 = packing.pack_matrix(torch.randn(360, 180, generator=g) / 13, 384, 192, device=dev)
w2 = packing.pack_matrix(torch.randn(180, 360, generator=g) / 19, 192, 384, device=dev)
bp, b1, b2 = torch.zeros(192, device=dev), torch.zeros(384, device=dev), torch.zeros(192, device=dev)
gam, bet = torch.ones(192, device=dev), torch.zeros(192, device=dev)
gam[180:] = 0
se = torch.rand(B, 192, device=dev)
wcat = torch.empty(B, 192, 384, dtype=BF, device=dev)
ops.build_concat_diag_weights(wp, se, 0.01, wcat)
lno = torch.empty(M, 192, device=dev, dtype=BF)
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
def two(c):
    if c:
        ops.conv_gemm(att, B, S, S, 192, wcat.view(B * 192, 384), n_store=192, w_batch_rows=192, bias=bp, x2=cab, res=x, out_f32=x, ln=(gam, bet, 1e-5, 180, t))
    else:
        ops.conv_gemm(att, B, S, S, 192, wp, n_store=192, bias=bp, res=x, out_f32=x, ln=(gam, bet, 1e-5, 180, t))
    ops.mlp_fused(t, B, S, S, w1, b1, w2, b2, x, ln=(gam, bet, 1e-5, 180, lno))
def one(c):
    if c == 2:      # the diagonal K block generated inside the kernel (what hat.py runs)
        ops.hab_tail(att, B, S, S, wp, bp, x, (gam, bet), w1, b1, w2, b2, x, a1=cab, a1_diag=se, a1_alpha=0.01, ln=(gam, bet, lno))
        return
    ops.hab_tail(att, B, S, S, wcat.view(B * 192, 384) if c else wp, bp, x, (gam, bet), w1, b1, w2, b2, x, a1=cab if c else None, wp_batch_rows=192 if c else 0,
                 ln=(gam, bet, lno))
for c in (2, True, False):
    a = timed(lambda: two(c))
    b = timed(lambda: one(c))
    byts = M * 192 * (2 * (2 if c else 1) + 8 + 2)
    print(f"HAB tail B={B} S={S} cab={c}: two kernels {a:7.1f} us   fused {b:7.1f} us  ({byts / b / 1e3:.0f} GB/s compulsory)")
