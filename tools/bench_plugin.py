"""Wall time of the plugin call models.team29_FreqFusion.main() on a tmpfs folder (development helper).
    python tools/bench_plugin.py N H W [steps]        # N synthetic PNGs of H x W LR pixels
Environment knobs of the pipeline are honoured (FFB200_TILE_BATCH, FFB200_GRAPH_MAX_LR_PIXELS, FFB200_FORCE_TILING, FFB200_IO_THREADS, ...)."""
import contextlib, io, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from PIL import Image
import bench
from isr2_b200 import io as ffio, weights

N, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
base = "/dev/shm" if os.path.isdir("/dev/shm") else None
root = tempfile.mkdtemp(prefix="ffb200_bp_", dir=base)
fusion = weights.save_checkpoints(root, seed=0)
os.environ["FFB200_PRETRAINED_ROOT"] = root
din, dout = os.path.join(root, "in"), os.path.join(root, "out")
os.makedirs(din)
for i in range(N):
    Image.fromarray(bench.synth_image(H, W, 100 + i)).save(os.path.join(din, f"{i:04d}.png"), compress_level=1)
dev = torch.device("cuda:0")
with contextlib.redirect_stdout(io.StringIO()):
    ffio.main(fusion, din, dout, dev)
    ffio.main(fusion, din, dout, dev)
torch.cuda.synchronize()
ts = []
for _ in range(steps):
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        ffio.main(fusion, din, dout, dev)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
ms = sum(ts) / len(ts)
print(f"main(): {N} x {H}x{W}: {ms:.1f} ms/call (min {min(ts):.1f})  -> {N * 16 * H * W / 1e3 / ms:.2f} unique Mpix/s   env: " +
      " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("FFB200_") and k != "FFB200_PRETRAINED_ROOT"))
shutil.rmtree(root, ignore_errors=True)
