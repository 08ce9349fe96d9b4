"""ORACLE (test infrastructure only -- never imported by the product path).

Plain-PyTorch fp32 restatement of the reference NAFNet-SR forward on a raw state_dict (inner `nafnet.` keys).
Follows /root/reference/src/models/nafnet/nafnet_arch.py (LayerNorm2d :26-41, NAFBlock.forward :110-131,
NAFNet.forward :195-225, check_image_size :219-225) and src/models/nafnet/__init__.py:117-139 (NAFNetSR.forward),
plus ExpertEnsemble.forward_nafnet (src/models/expert_loader.py:660-674).
"""
import torch
import torch.nn.functional as F


def _ln2d(x, w, b, eps=1e-6):
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return w.view(1, -1, 1, 1) * ((x - u) / torch.sqrt(s + eps)) + b.view(1, -1, 1, 1)


def _block(x, sd, p):
    c = x.shape[1]
    y = _ln2d(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    y = F.conv2d(y, sd[p + "conv1.weight"], sd[p + "conv1.bias"])
    y = F.conv2d(y, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1, groups=2 * c)
    y = y[:, :c] * y[:, c:]
    y = y * F.conv2d(y.mean((2, 3), keepdim=True), sd[p + "sca.1.weight"], sd[p + "sca.1.bias"])
    y = F.conv2d(y, sd[p + "conv3.weight"], sd[p + "conv3.bias"])
    x = x + y * sd[p + "beta"]
    y = F.conv2d(_ln2d(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"]), sd[p + "conv4.weight"], sd[p + "conv4.bias"])
    y = y[:, :c] * y[:, c:]
    y = F.conv2d(y, sd[p + "conv5.weight"], sd[p + "conv5.bias"])
    return x + y * sd[p + "gamma"]


def nafnet_forward(sd, inp, enc=(2, 2, 4, 8), dec=(2, 2, 2, 2), mid=12):
    B, C, H, W = inp.shape
    ps = 2 ** len(enc)
    inp = F.pad(inp, (0, (ps - W % ps) % ps, 0, (ps - H % ps) % ps))
    x = F.conv2d(inp, sd["intro.weight"], sd["intro.bias"], padding=1)
    skips = []
    for s, n in enumerate(enc):
        for k in range(n):
            x = _block(x, sd, f"encoders.{s}.{k}.")
        skips.append(x)
        x = F.conv2d(x, sd[f"downs.{s}.weight"], sd[f"downs.{s}.bias"], stride=2)
    for k in range(mid):
        x = _block(x, sd, f"middle_blks.{k}.")
    for s, n in enumerate(dec):
        x = F.pixel_shuffle(F.conv2d(x, sd[f"ups.{s}.0.weight"]), 2) + skips[-1 - s]
        for k in range(n):
            x = _block(x, sd, f"decoders.{s}.{k}.")
    x = F.conv2d(x, sd["ending.weight"], sd["ending.bias"], padding=1) + inp
    return x[:, :, :H, :W]


def forward_nafnet(sd, x, **kw):
    up = F.interpolate(x, scale_factor=4, mode="bicubic", align_corners=False)
    return nafnet_forward(sd, up, **kw).clamp(0, 1)
