"""Cycle counters of ff_hab_tail (block 0) from a -DFF_TAIL_PROF build (development helper):
    FFB200_NVCC_EXTRA=-DFF_TAIL_PROF python image-super-resolution-2_b200/build.py --force && python tools/tail_phases.py 16 128"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200 import lib, ops, packing
dev = torch.device("cuda:0")
B, S = int(sys.argv[1]), int(sys.argv[2])
cabon = len(sys.argv) < 4 or sys.argv[3] != "0"
M = B * S * S
BF = torch.bfloat16
att = torch.randn(M, 192).to(dev, BF); cab = torch.randn(M, 192).to(dev, BF); x = torch.randn(M, 192).to(dev)
wp = packing.pack_matrix(torch.randn(180, 180) / 13, 192, 192, device=dev)
w1 = packing.pack_matrix(torch.randn(360, 180) / 13, 384, 192, device=dev); w2 = packing.pack_matrix(torch.randn(180, 360) / 19, 192, 384, device=dev)
bp, b1, b2 = torch.zeros(192, device=dev), torch.zeros(384, device=dev), torch.zeros(192, device=dev)
gam, bet = torch.ones(192, device=dev), torch.zeros(192, device=dev)
wcat = torch.empty(B, 192, 384, dtype=BF, device=dev)
ops.build_concat_diag_weights(wp, torch.rand(B, 192, device=dev), 0.01, wcat)
lno = torch.empty(M, 192, device=dev, dtype=BF)
L = lib.load()
buf = (ctypes.c_ulonglong * 32)()
def run():
    ops.hab_tail(att, B, S, S, wcat.view(B * 192, 384) if cabon else wp, bp, x, (gam, bet), w1, b1, w2, b2, x, a1=cab if cabon else None,
                 wp_batch_rows=192 if cabon else 0, ln=(gam, bet, lno))
run(); L.ff_debug_tail_prof(buf, 1)
run(); L.ff_debug_tail_prof(buf, 1)
names = {0: "mma: wait w_full", 1: "mma g0: wait acc2_empty", 2: "mma g1: wait acc1_empty", 3: "mma g2: wait h_full", 4: "mma: wait a_full", 7: "mma: issue + rest",
         8: "mid: wait g0_full", 9: "mid: pass 1 work", 10: "mid: wait residual", 11: "mid: wait a_empty", 12: "mid: pass 2",
         16: "fin: wait acc2_full", 17: "fin: pass 1", 18: "fin: pass 2 (LN)"}
tiles = max(1, buf[31])
print(f"tiles of block 0: {tiles}")
for k, n in names.items():
    print(f"{n:28s} {buf[k]:12d} cycles  ({buf[k] / tiles:9.0f} per tile)")
