"""BASELINE.json configs[4]-shaped sweep on one GPU: LR tile side x tiles per batch -> Mpix/s of the full model (device-resident
inputs, CUDA events, 2 warm-up + 3 timed forwards per point).  Prints one JSON line per point and a table."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from isr2_b200.model import FreqFusionB200

dev = torch.device("cuda:0")
m = FreqFusionB200(dev, init_seed=0, verbose=False)
points = [(64, 1), (64, 4), (64, 16), (64, 64), (128, 1), (128, 4), (128, 16), (256, 1), (256, 4)]
rows = []
for S, B in points:
    x = torch.rand(B, 3, S, S, device=dev)
    out = torch.empty(B, 3, 4 * S, 4 * S, device=dev)
    for _ in range(2):
        m.forward(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m.forward(x, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    mpix = B * 16 * S * S / 1e6 / (ms / 1e3)
    rows.append((S, B, ms, mpix))
    print(json.dumps({"lr_tile": S, "tiles": B, "ms_per_forward": ms, "mpix_per_s": mpix, "tflops_algorithmic": 10.793 * mpix / 1e0 / 1e6 * 1e6 / 1e6}))
    del x, out
    torch.cuda.empty_cache()
print(f"{'LR tile':>8s} {'tiles':>6s} {'ms':>9s} {'Mpix/s':>8s}")
for S, B, ms, mpix in rows:
    print(f"{S:8d} {B:6d} {ms:9.2f} {mpix:8.2f}")
