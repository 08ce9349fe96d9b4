"""Cached-expert file format of the reference (src/data/cached_dataset.py:9-23, 45-67, 87-122, 135-200): the on-disk
pairs `{stem}[_p{0-4}]_hat_part.pt` (or `_drct_part.pt`) + `{stem}_rest_part.pt` written by the reference's feature
extraction, read here so that the B200 fusion head (`FreqFusionB200.forward_with_precomputed`) can serve that workflow.

Same discovery rules as `CachedSRDataset.__init__` (sorted `*_hat_part.pt` first, else `*_drct_part.pt`; stems without a
`_rest_part.pt` twin are dropped with a warning) and the same per-sample record as `__getitem__` without the training-time
augmentation: lr [3,h,w], hr [3,4h,4w], expert_imgs {hat, dat, nafnet} with the aliases drct -> hat, grl -> dat, batch
dimension squeezed, optional expert_feats.  Host-side plumbing only; the arithmetic stays in the CUDA head.
"""
import os
from pathlib import Path

import torch

EXPERT_KEY_MAP = {"drct": "hat", "grl": "dat"}      # cached_dataset.py:62-66


def _normalize_keys(d):
    """cached_dataset.py:202-212."""
    return {EXPERT_KEY_MAP.get(k, k): v for k, v in d.items()}


class CachedExpertStore:
    def __init__(self, feature_dir, load_features=False, verbose=False):
        self.feature_dir = Path(feature_dir)
        self.load_features = load_features
        if not self.feature_dir.exists():
            raise RuntimeError(f"Feature cache directory not found: {feature_dir}")
        hat_files = sorted(self.feature_dir.glob("*_hat_part.pt"))
        drct_files = sorted(self.feature_dir.glob("*_drct_part.pt"))
        if hat_files:
            self._sentinel_suffix, sentinel = "_hat_part.pt", hat_files
        elif drct_files:
            self._sentinel_suffix, sentinel = "_drct_part.pt", drct_files
        else:
            raise RuntimeError(f"No cached features found in {feature_dir}!\nExpected *_hat_part.pt or *_drct_part.pt files.")
        stems = [f.name.replace(self._sentinel_suffix, "") for f in sentinel]
        missing = [s for s in stems if not (self.feature_dir / f"{s}_rest_part.pt").exists()]
        if missing:
            print(f"Warning: {len(missing)} files missing rest_part counterparts")
        self.file_stems = [s for s in stems if s not in missing]
        if verbose:
            print(f"CachedExpertStore: {len(self.file_stems)} samples in {feature_dir} (*{self._sentinel_suffix} + *_rest_part.pt)")

    def __len__(self):
        return len(self.file_stems)

    def __getitem__(self, idx):
        stem = self.file_stems[idx]
        primary = torch.load(self.feature_dir / f"{stem}{self._sentinel_suffix}", weights_only=False)
        rest = torch.load(self.feature_dir / f"{stem}_rest_part.pt", weights_only=False)
        imgs = {}
        imgs.update(_normalize_keys(primary["outputs"]))
        imgs.update(_normalize_keys(rest["outputs"]))
        imgs = {k: (v.squeeze(0) if v.dim() == 4 else v) for k, v in imgs.items()}
        rec = {"lr": primary["lr"], "hr": primary["hr"], "expert_imgs": imgs, "filename": stem}
        if self.load_features:
            feats = {}
            feats.update(_normalize_keys(primary.get("features", {})))
            feats.update(_normalize_keys(rest.get("features", {})))
            rec["expert_feats"] = {k: (v.squeeze(0) if v.dim() == 4 else v) for k, v in feats.items()}
        return rec

    def batches(self, batch_size):
        """Consecutive samples of equal LR size grouped into batches of at most `batch_size` (the head is batch independent)."""
        cur, shape = [], None
        for i in range(len(self)):
            rec = self[i]
            s = tuple(rec["lr"].shape)
            if cur and (s != shape or len(cur) == batch_size):
                yield cur
                cur = []
            cur.append(rec)
            shape = s
        if cur:
            yield cur


def run_cached(model, store, batch_size=8, crop=4):
    """Fusion head over a cached-expert directory: returns a list of per-sample records
    {filename, sr (fp32 [3,4h,4w] on the device), psnr_y, ssim_y} -- PSNR / SSIM against the cached HR patch on the BT.601 Y
    channel with `crop` border pixels removed (src/utils/metrics.py), computed on the GPU."""
    from . import ops
    out = []
    dev = model.device
    for recs in store.batches(batch_size):
        lr = torch.stack([r["lr"] for r in recs]).float().to(dev)
        hr = torch.stack([r["hr"] for r in recs]).float().to(dev)
        ex = {k: torch.stack([r["expert_imgs"][k] for r in recs]).float().to(dev) for k in ("hat", "dat", "nafnet")}
        feats = None
        if all("expert_feats" in r and all(k in r["expert_feats"] for k in ("hat", "dat", "nafnet")) for r in recs):
            # cached features present (load_features=True): the collaborative branch runs, as in the reference's cached mode
            feats = {k: torch.stack([r["expert_feats"][k] for r in recs]).float().to(dev) for k in ("hat", "dat", "nafnet")}
        sr = model.forward_with_precomputed(lr, ex, feats)
        psnr, ssim = ops.psnr_y(sr, hr, crop).cpu(), ops.ssim_y(sr, hr, crop).cpu()
        for i, r in enumerate(recs):
            out.append({"filename": r["filename"], "sr": sr[i].clone(), "psnr_y": psnr[i].item(), "ssim_y": ssim[i].item()})
    return out
