import contextlib, io, os, shutil, sys, tempfile, time, cProfile, pstats
ROOT = "/root/repo"
sys.path.insert(0, ROOT)
import torch
from PIL import Image
import bench
from isr2_b200 import io as ffio, weights
root = tempfile.mkdtemp(prefix="ffb200_bp_", dir="/dev/shm")
fusion = weights.save_checkpoints(root, seed=0)
os.environ["FFB200_PRETRAINED_ROOT"] = root
din, dout = os.path.join(root, "in"), os.path.join(root, "out")
os.makedirs(din)
for i in range(16):
    Image.fromarray(bench.synth_image(128, 128, 100 + i)).save(os.path.join(din, f"{i:04d}.png"), compress_level=1)
dev = torch.device("cuda:0")
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(3): ffio.main(fusion, din, dout, dev)
torch.cuda.synchronize()
pr = cProfile.Profile()
with contextlib.redirect_stdout(io.StringIO()):
    pr.enable()
    for _ in range(3): ffio.main(fusion, din, dout, dev)
    pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(45)
shutil.rmtree(root, ignore_errors=True)
