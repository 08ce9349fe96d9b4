"""ORACLE (test infrastructure only -- never imported by the product path).

Plain-PyTorch fp32 restatement of the reference's collaborative branch, the part of
`CompleteEnhancedFusionSR.forward_with_precomputed(lr, expert_outputs, expert_features=...)` that runs whenever expert features
are passed (src/models/enhanced_fusion.py:466-496, :790-795): `EnhancedCollaborativeWithLKA.forward`
(src/models/large_kernel_attention.py:327-419) on a raw fusion state_dict (keys `collaborative.*`).
Pinned by tests/golden/head_collab_64.pt (output of the unmodified reference, oracle/make_golden_collab.py).
"""
import torch
import torch.nn.functional as F

from . import head as ohead

NAMES = ("hat", "dat", "nafnet")


def collaborative(sd, feats, outputs, return_mod=False):
    """feats: {'hat': [B,180,h,w], 'dat': [B,180,h,w], 'nafnet': [B,64,h,w]}; outputs: [hat, dat, nafnet] SR tensors.
    Returns the three modulated SR tensors (large_kernel_attention.py:327-419).  As in the reference (:337-378) a feature map with
    more channels than its align layer expects is truncated, one with fewer is zero padded, aligned maps are brought to the smallest
    spatial size among them (bilinear, align_corners=False), and an expert without features contributes zeros."""
    p = "collaborative."
    dim, heads = 128, 8
    aligned = {}
    for n in NAMES:
        if n not in feats:
            continue
        w_, b_ = sd[p + f"align_layers.{n}.weight"], sd[p + f"align_layers.{n}.bias"]
        f, want = feats[n], w_.shape[1]
        if f.shape[1] > want:
            f = f[:, :want]
        elif f.shape[1] < want:
            f = F.pad(f, (0, 0, 0, 0, 0, want - f.shape[1]))
        aligned[n] = F.conv2d(f, w_, b_)
    if not aligned:
        return (list(outputs), None) if return_mod else list(outputs)
    H, W = min(a.shape[2] for a in aligned.values()), min(a.shape[3] for a in aligned.values())
    for n in aligned:
        if aligned[n].shape[2:] != (H, W):
            aligned[n] = F.interpolate(aligned[n], size=(H, W), mode="bilinear", align_corners=False)
    B = next(iter(aligned.values())).shape[0]
    like = next(iter(aligned.values()))
    flat = torch.stack([aligned.get(n, torch.zeros_like(like)) for n in NAMES], 1).permute(0, 3, 4, 1, 2).reshape(B * H * W, 3, dim)
    n1 = F.layer_norm(flat, (dim,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    qkv = F.linear(n1, sd[p + "cross_attn.in_proj_weight"], sd[p + "cross_attn.in_proj_bias"])
    q, k, v = [t.view(-1, 3, heads, dim // heads).transpose(1, 2) for t in qkv.chunk(3, -1)]
    a = ((q * (dim // heads) ** -0.5) @ k.transpose(-2, -1)).softmax(-1) @ v
    a = a.transpose(1, 2).reshape(-1, 3, dim)
    flat = flat + F.linear(a, sd[p + "cross_attn.out_proj.weight"], sd[p + "cross_attn.out_proj.bias"])
    n2 = F.layer_norm(flat, (dim,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    flat = flat + F.linear(F.gelu(F.linear(n2, sd[p + "ffn.0.weight"], sd[p + "ffn.0.bias"])), sd[p + "ffn.2.weight"], sd[p + "ffn.2.bias"])
    enhanced = flat.reshape(B, H, W, 3, dim).permute(0, 3, 4, 1, 2)
    res, mods = [], []
    for i, out in enumerate(outputs):
        f = ohead.lka_block(sd, p + "lka_global.", enhanced[:, i])
        f = F.interpolate(f, size=out.shape[-2:], mode="bilinear", align_corners=False)
        m = F.gelu(F.conv2d(f, sd[p + f"modulation.{i}.0.weight"], sd[p + f"modulation.{i}.0.bias"])).mean((2, 3), keepdim=True)
        m = torch.sigmoid(F.conv2d(m, sd[p + f"modulation.{i}.3.weight"], sd[p + f"modulation.{i}.3.bias"]))
        mods.append(m.flatten(1))
        res.append((out * (1.0 + 0.2 * (m - 0.5))).clamp(0, 1))
    return (res, torch.stack(mods, 1)) if return_mod else res


def head_forward_with_features(sd, lr, experts, feats, return_intermediates=False):
    """forward_with_precomputed with expert_features: collaborative modulation of the expert outputs, then the usual head."""
    return ohead.head_forward(sd, lr, collaborative(sd, feats, experts), return_intermediates)


def synth_features(B, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return {"hat": torch.randn(B, 180, h, w, generator=g) * 0.5, "dat": torch.randn(B, 180, h, w, generator=g) * 0.5,
            "nafnet": torch.randn(B, 64, h, w, generator=g) * 0.5}


def synth_features_mixed(B, h, w, seed):
    """Features that take every branch of the reference's alignment step: 200 channels for hat (truncated to 180), 150 for dat
    (zero padded to 180), nafnet at twice the spatial size (its aligned map is resized down to h x w)."""
    g = torch.Generator().manual_seed(seed)
    return {"hat": torch.randn(B, 200, h, w, generator=g) * 0.5, "dat": torch.randn(B, 150, h, w, generator=g) * 0.5,
            "nafnet": torch.randn(B, 64, 2 * h, 2 * w, generator=g) * 0.5}
