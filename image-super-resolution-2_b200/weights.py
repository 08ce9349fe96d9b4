"""state_dict layout (state_layout.json) and a deterministic synthetic-weight factory.

The shipped fusion checkpoint and the pretrained expert weights are not available offline
(SURVEY.md section 0), so parity tests and bench.py run on seeded synthetic weights.  The factory is a
pure function of (tensor name, shape, seed): the same call regenerates the same tensors in the
build container (where they are loaded into the imported reference) and on the GPU box.
Identity-at-init tensors of the reference (NAFNet beta/gamma = 0, BatchNorm running stats 0/1,
LKA / ResBlock scales) are perturbed so those code paths are actually exercised.
"""
import hashlib
import json
import math
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LAYOUT = None

# Structural (input-independent) buffers: the factory leaves them out so the reference keeps the
# values it computes itself; the product recomputes them arithmetically.
STRUCTURAL = ("relative_position_index", "rpe_biases", "attn_mask_", "dct_basis", "low_mask", "mid_mask", "high_mask",
              "lo_row", "hi_row", "lo_col", "hi_col", "gaussian.kernel", "num_batches_tracked")


def layout():
    """{'hat'|'dat'|'nafnet'|'fusion': {key: (shape, dtype_name)}} -- the checkpoint contract."""
    global _LAYOUT
    if _LAYOUT is None:
        with open(os.path.join(_HERE, "state_layout.json")) as f:
            raw = json.load(f)
        _LAYOUT = {m: {k: (tuple(v[0]), v[1]) for k, v in d.items()} for m, d in raw.items()}
    return _LAYOUT


def _gen(name, seed):
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return torch.Generator().manual_seed(int.from_bytes(h[:7], "little"))


def _radial_logits(size):
    y = torch.linspace(-1, 1, size)
    yy, xx = torch.meshgrid(y, y, indexing="ij")
    return (3.0 * (0.5 - torch.sqrt(xx ** 2 + yy ** 2))).view(1, 1, size, size)


# name-suffix -> (base value, noise std) for scalars / small special parameters
_SPECIAL = [
    ("lka_block.scale1", 0.1, 0.02), ("lka_block.scale2", 0.1, 0.02), ("lka_global.scale1", 0.1, 0.02),
    ("lka_global.scale2", 0.1, 0.02), ("_res.scale", 0.1, 0.02), ("residual_scale", 0.1, 0.01),
    ("residual_weight_1_2", 0.2, 0.03), ("residual_weight_2_3", 0.2, 0.03), ("edge_strength", 0.15, 0.02),
    ("level_weights", 1.0 / 3, 0.1), ("fft.temperature", 5.0, 0.3), ("dct.band_scale", 1.0, 0.1),
    ("fft.band_scale", 1.0, 0.1), ("subband_scale", 1.0, 0.1), ("dct_importance", 1.0, 0.1),
    ("dwt_importance", 0.8, 0.1), ("fft_importance", 0.6, 0.1), ("expert_weights", 1.0, 0.1),
    ("band_importance", 1.0, 0.1), ("attn.temperature", 1.0, 0.2),
]


def make_tensor(name, shape, dtype, seed, model):
    g = _gen(name, seed)
    rn = lambda std=1.0: torch.randn(shape, generator=g) * std
    if dtype != "float32":
        return None
    for suf, base, std in _SPECIAL:
        if name.endswith(suf):
            return base + rn(std)
    if name.endswith("freq_mask_logits"):
        return _radial_logits(shape[-1]) + rn(0.3)
    if name.endswith("running_mean"):
        return rn(0.1)
    if name.endswith("running_var"):
        return 0.6 + 0.8 * torch.rand(shape, generator=g)
    if name.endswith("relative_position_bias_table"):
        return rn(0.4)
    if name.endswith(".beta") or name.endswith(".gamma"):     # NAFNet residual scales (zeros at init)
        return rn(0.25)
    if name.endswith(".bias"):
        return rn(0.03)
    if name.endswith(".weight") or name.endswith("in_proj_weight"):
        if len(shape) == 1:                                    # LayerNorm / BatchNorm / LayerNorm2d gain
            return 1.0 + rn(0.1)
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        gain = 0.8
        # keep image-space outputs inside [0,1] most of the time so clamps do not hide errors
        if name in ("conv_last.weight", "ending.weight") or name.endswith("refine_net.6.weight") or name.endswith("edge_refine.fusion.2.weight"):
            gain = 0.15
        return rn(gain / math.sqrt(fan_in))
    if name.endswith("in_proj_bias"):
        return rn(0.03)
    raise KeyError(f"weights.make_tensor: no rule for {name} {shape}")


def make_state_dict(model, seed=0):
    """Synthetic state_dict for model in {'hat','dat','nafnet','fusion'} (structural buffers omitted)."""
    sd = {}
    for name, (shape, dtype) in layout()[model].items():
        if any(s in name for s in STRUCTURAL):
            continue
        t = make_tensor(name, shape, dtype, seed, model)
        if t is not None:
            sd[name] = t.to(torch.float32).reshape(shape).contiguous()
    return sd


def heat(sd, model, logit_gain=2.5, table_gain=3.0):
    """'Hot' variant of a synthetic state dict: attention logits and relative-position biases in the range pretrained
    transformers reach (peaky softmax rows, |logit| of several units) instead of the near-uniform attention the benign factory
    gives -- stresses the exp2-softmax, the bias / mask path and the bf16 operand rounding.
    q and k rows of every qkv projection are scaled by `logit_gain` (logits by its square), bias tables / DAT's dynamic
    position-bias output layer / channel-attention temperatures by `table_gain`.  Returns a new dict."""
    out = {k: v.clone() for k, v in sd.items()}
    if model not in ("hat", "dat"):
        return out
    for k, v in out.items():
        if k.endswith("qkv.weight") or k.endswith("qkv.bias"):
            v[: 2 * 180] *= logit_gain
        elif k.endswith("relative_position_bias_table") or k.endswith("attn.temperature"):
            v *= table_gain
        elif ".pos.pos3.2." in k:
            v *= table_gain
    return out


def save_checkpoints(root, seed=0):
    """Write the four checkpoint files in the formats the reference ingests (io.py:131-137,164-165;
    expert_loader.py:99-169; nafnet/__init__.py:84-115).  Returns the fusion checkpoint path."""
    paths = {
        "hat": os.path.join(root, "pretrained", "hat", "HAT-L_SRx4_ImageNet-pretrain.pth"),
        "dat": os.path.join(root, "pretrained", "dat", "DAT_x4.pth"),
        "nafnet": os.path.join(root, "pretrained", "nafnet", "NAFNet-SIDD-width64.pth"),
    }
    for m, p in paths.items():
        os.makedirs(os.path.dirname(p), exist_ok=True)
        key = "params_ema" if m == "hat" else "params"
        torch.save({key: make_state_dict(m, seed)}, p)
    fusion = os.path.join(root, "fusion_synthetic.pth")
    torch.save({"epoch": 0, "model_state_dict": make_state_dict("fusion", seed), "metrics": {}}, fusion)
    return fusion
