"""ff_naf_tail vs the three ff_conv_gemm passes (conv3 + residual + norm2, conv4 + SimpleGate, conv5 + residual + norm1') of a
64-channel NAFBlock at NAFNet-SR's full-resolution level of the bench shape (development helper): python tools/bench_naf.py 16 512"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import ops, packing
from isr2_b200.nafnet import _gate_perm
dev = torch.device("cuda:0")
B, S = int(sys.argv[1]), int(sys.argv[2])
M, C = B * S * S, 64
g = torch.Generator().manual_seed(0)
BF = torch.bfloat16
gt = torch.randn(M, C, generator=g).to(dev, BF)
x = torch.randn(M, C, generator=g).to(dev)
t = torch.empty(M, C, device=dev, dtype=BF)
gt2 = torch.empty(M, C, device=dev, dtype=BF)
w3 = (torch.randn(B * C, C, generator=g) / 8).to(dev, BF)
w4f = torch.randn(2 * C, C, generator=g) / 8
w4n = packing.pack_matrix(w4f, 2 * C, C, device=dev)
w4p = packing.pack_matrix(w4f[_gate_perm(C)], 2 * C, C, device=dev)
w5 = packing.pack_matrix(torch.randn(C, C, generator=g) / 8, C, C, device=dev)
b3, b4, b5 = torch.zeros(C, device=dev), torch.zeros(2 * C, device=dev), torch.zeros(C, device=dev)
gam, bet, one = torch.ones(C, device=dev), torch.zeros(C, device=dev), torch.ones(C, device=dev)
def timed(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
def three():
    ops.conv_gemm(gt, B, S, S, C, w3, n_store=C, w_batch_rows=C, bias=b3, col_scale=one, res=x, out_f32=x, ln=(gam, bet, 1e-6, C, t))
    ops.conv_gemm(t, B, S, S, C, w4p, n_store=2 * C, bias=b4, gate_pairs=1, out_bf16=gt2)
    ops.conv_gemm(gt2, B, S, S, C, w5, n_store=C, bias=b5, col_scale=one, res=x, out_f32=x, ln=(gam, bet, 1e-6, C, t))
def fused():
    ops.naf_tail(gt, B, S, S, w3, b3, x, (gam, bet), w4n, b4, w5, b5, x, w3_batch_rows=C, out_bf16=t, ln=(gam, bet))
a = timed(three)
b = timed(fused)
byts = M * C * (2 + 8 + 2)
print(f"NAF tail B={B} {S}x{S}: three kernels {a:7.1f} us   fused {b:7.1f} us  ({byts / b / 1e3:.0f} GB/s compulsory)")
