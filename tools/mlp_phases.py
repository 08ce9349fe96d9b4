"""Cycle counters of the fused-MLP kernel (block 0) from a -DFF_MLP_PROF build (development helper):
    FFB200_NVCC_EXTRA=-DFF_MLP_PROF python image-super-resolution-2_b200/build.py --force && python tools/mlp_phases.py 16 128"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import lib, ops, packing
dev = torch.device("cuda:0")
B, S = int(sys.argv[1]), int(sys.argv[2])
M = B * S * S
t = torch.randn(M, 192).to(dev, torch.bfloat16); x = torch.randn(M, 192).to(dev)
w1 = packing.pack_matrix(torch.randn(360, 180) / 13, 384, 192, device=dev); w2 = packing.pack_matrix(torch.randn(180, 360) / 19, 192, 384, device=dev)
b1, b2 = torch.zeros(384, device=dev), torch.zeros(192, device=dev)
L = lib.load()
buf = (ctypes.c_ulonglong * 32)()
ops.mlp_fused(t, B, S, S, w1, b1, w2, b2, x)
L.ff_debug_mlp_prof(buf, 1)
ops.mlp_fused(t, B, S, S, w1, b1, w2, b2, x)
L.ff_debug_mlp_prof(buf, 1)
names = {0: "producer: wait w_empty", 1: "mma g1: wait w_full", 2: "mma g1: wait acc1_empty", 3: "mma g1: issue", 4: "mma g2: wait w_full", 5: "mma g2: wait h_full",
         6: "mma g2: wait acc2_empty", 7: "mma g2: issue", 8: "mma: wait a_full", 10: "gelu: wait acc1_full", 11: "gelu: ld + arrive", 12: "gelu: wait h_empty",
         13: "gelu: math + store", 16: "final: wait acc2_full", 17: "final: epilogue"}
tiles = (M // 128 + 147) // 148
for k, n in names.items():
    print(f"{n:28s} {buf[k]:12d} cycles  ({buf[k] / tiles:9.0f} per tile)")
