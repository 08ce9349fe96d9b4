"""BASELINE.json configs[3]: DIV2K-shaped 339x510 LR images, tiles 128/32 (20 tiles per image) -> 1356x2040.
Reports unique-pixel throughput of the tiled path on device-resident images, and of io.main with PNG I/O."""
import sys, os, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image
from isr2_b200 import io as ffio, weights
from isr2_b200.model import FreqFusionB200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
imgs = [((torch.rand(1, 3, 339, 510, generator=g) * 255).round() / 255) for _ in range(n)]
m = FreqFusionB200(dev, verbose=False)
for _ in range(2):
    ffio.tiled_forward(m, imgs[0].to(dev), 128, 32, return_u8=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for im in imgs:
    u8 = ffio.tiled_forward(m, im.to(dev), 128, 32, return_u8=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
mp = n * 1356 * 2040 / 1e6
print(f"tiled forward (20 tiles/image, batched, stitched to uint8 on device): {n} images in {dt:.2f} s -> {mp/dt:.2f} unique Mpix/s ({n*20*512*512/1e6/dt:.2f} computed Mpix/s)")
# full plugin call with PNG I/O
root = tempfile.mkdtemp()
fusion = weights.save_checkpoints(root, seed=0)
os.environ["FFB200_PRETRAINED_ROOT"] = root
inp, outp = os.path.join(root, "in"), os.path.join(root, "out")
os.makedirs(inp)
for i, im in enumerate(imgs):
    Image.fromarray((im[0].permute(1, 2, 0).numpy() * 255).round().astype("uint8")).save(os.path.join(inp, f"{i:04d}.png"))
t0 = time.perf_counter()
ffio.main(fusion, inp, outp, dev)
dt = time.perf_counter() - t0
print(f"io.main incl. model build/weight packing + PNG decode/encode: {dt:.2f} s for {n} images")
