"""PSNR / SSIM evaluation harness on the GPU: the host-side mirror of the reference's `eval.py` for the two full-reference scores
it takes from `utils/utils_image.py:287-312` (`cal_psnr_ssim`, eval.py:157).

Same command line, same file pairing and partitioning, same result files:

    python -m isr2_b200.evaluate --output_folder out/ --target_folder div2k-val/HR --metrics_save_path ./IQA_results --gpu_ids 0,1

* `output_folder` / `target_folder`: sorted `*.png`; the target of an output file is its name with `x4` removed (eval.py:142) and
  both folders must hold the same number of images (eval.py:188-191).
* the file list is cut into contiguous partitions, one per GPU id, the last one taking the remainder (eval.py:166-170); one
  process per GPU (`torch.multiprocessing`, eval.py:202-213), each decoding its files, copying the uint8 pixels to its GPU and
  running `ff_eval_psnr_ssim_u8` (OpenCV's 8-bit luma, scikit-image's 7x7 uniform-window SSIM, border 4).
* `<metrics_save_path>/<parent>--<folder>.csv` holds one row per file (`Filename, psnr, ssim`, keys sorted as eval.py:233-270
  sorts them) and `<parent>--<folder>.txt` the averages.

Not built: the six no-reference / perceptual scores of eval.py (LPIPS, DISTS, NIQE, MUSIQ,
MANIQA, CLIP-IQA) -- they are pretrained `pyiqa` networks whose weights are not available offline -- and with them the
"Total Score" line, which is a sum over exactly those six.
"""
import argparse
import csv
import os

import numpy as np
import torch
from PIL import Image

from . import ops

BORDER = 4      # cal_psnr_ssim's default (utils_image.py:287)


def list_pairs(output_folder, target_folder):
    """eval.py:184-191 + :142: sorted output PNGs, each paired with `name.replace('x4', '')` in the target folder."""
    outs = sorted(f for f in os.listdir(output_folder) if f.endswith(".png"))
    tgts = sorted(f for f in os.listdir(target_folder) if f.endswith(".png"))
    if len(outs) != len(tgts):
        raise AssertionError(f"The number of output images should be equal to the number of target images: {len(outs)} != {len(tgts)}")
    pairs = []
    for f in outs:
        t = os.path.join(target_folder, f.replace("x4", ""))
        if not os.path.exists(t):
            raise AssertionError(f"No such path: {t}")
        pairs.append((f, os.path.join(output_folder, f), t))
    return pairs


def partition(n_files, rank, num_gpus):
    """eval.py:166-170: contiguous partitions of floor(n / num_gpus) files, the last rank takes the remainder."""
    size = n_files // num_gpus
    start = rank * size
    end = (rank + 1) * size if rank != num_gpus - 1 else n_files
    return start, end


def _load_rgb_u8(path):
    """imread_uint(path, 3) (utils_image.py:105-117): uint8 RGB, grey images replicated to three channels, alpha dropped."""
    with Image.open(path) as im:
        return np.array(im.convert("RGB"), dtype=np.uint8)      # (a writable copy: torch.from_numpy needs one)


def evaluate_files(pairs, device):
    """{output file name: {'psnr': .., 'ssim': ..}} for (name, output path, target path) triples, computed on `device`."""
    device = torch.device(device)
    if device.type != "cuda":
        raise ValueError("isr2_b200.evaluate runs on CUDA devices only (there is no CPU path)")
    results = {}
    with torch.cuda.device(device):
        pending = []
        for name, po, pt in pairs:
            a, b = _load_rgb_u8(po), _load_rgb_u8(pt)
            if a.shape != b.shape:
                raise ValueError(f"{name}: output {a.shape} and target {b.shape} differ in size")
            da = torch.from_numpy(a).to(device, non_blocking=False)
            db = torch.from_numpy(b).to(device, non_blocking=False)
            pending.append((name, ops.eval_psnr_ssim_u8(da, db, BORDER)))      # the next pair is decoded while this one runs
        for name, out in pending:
            psnr, ssim = out.cpu().tolist()
            results[name] = {"psnr": psnr, "ssim": ssim}
    return results


def _worker(rank, gpu_id, pairs, return_dict, num_gpus):
    start, end = partition(len(pairs), rank, num_gpus)
    return_dict[rank] = evaluate_files(pairs[start:end], torch.device("cuda", gpu_id))


def write_results(results, output_folder, metrics_save_path):
    """eval.py:219-284: per-file CSV and the averages (only the keys this harness produces)."""
    folder_name = os.path.basename(output_folder.rstrip("/"))
    next_level = os.path.basename(os.path.dirname(output_folder.rstrip("/")))
    os.makedirs(metrics_save_path, exist_ok=True)
    csv_path = f"{metrics_save_path}/{next_level}--{folder_name}.csv"
    txt_path = f"{metrics_save_path}/{next_level}--{folder_name}.txt"
    keys = sorted({k for v in results.values() for k in v})
    averages = {k: float(np.mean([v.get(k, 0) for v in results.values()])) for k in keys}
    with open(csv_path, mode="w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Filename"] + keys)
        for name, v in results.items():
            w.writerow([name] + [v.get(k, "") for k in keys])
    with open(txt_path, "w") as f:
        for k, v in averages.items():
            f.write(f"{k}: {v}\n")
    return averages, csv_path, txt_path


def run(output_folder, target_folder, metrics_save_path, gpu_ids=(0,)):
    pairs = list_pairs(output_folder, target_folder)
    gpu_ids = list(gpu_ids)
    if len(gpu_ids) == 1:
        results = evaluate_files(pairs, torch.device("cuda", gpu_ids[0]))
    else:
        import torch.multiprocessing as mp
        ctx = mp.get_context("spawn")
        manager = ctx.Manager()
        return_dict = manager.dict()
        procs = [ctx.Process(target=_worker, args=(rank, gid, pairs, return_dict, len(gpu_ids))) for rank, gid in enumerate(gpu_ids)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        if any(p.exitcode != 0 for p in procs):
            raise RuntimeError(f"evaluation worker failed (exit codes {[p.exitcode for p in procs]})")
        results = {}
        for rank in sorted(return_dict.keys()):
            results.update(return_dict[rank])
    if len(results) != len(pairs):
        raise RuntimeError(f"{len(pairs) - len(results)} images were not evaluated")
    averages, csv_path, txt_path = write_results(results, output_folder, metrics_save_path)
    print("Average:")
    print(averages)
    print(f"results: {csv_path}, {txt_path}")
    return results, averages


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--output_folder", type=str, default="output_dir")
    parser.add_argument("--target_folder", type=str, default="div2k-val/HR")
    parser.add_argument("--metrics_save_path", type=str, default="./IQA_results")
    parser.add_argument("--gpu_ids", type=str, default="0")
    args = parser.parse_args(argv)
    run(args.output_folder, args.target_folder, args.metrics_save_path, [int(g) for g in args.gpu_ids.split(",")])


if __name__ == "__main__":
    main()
