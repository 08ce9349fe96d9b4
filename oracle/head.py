"""ORACLE (test infrastructure only -- never imported by the product path).

Plain-PyTorch fp32 restatement of the eval-mode fusion head `CompleteEnhancedFusionSR.forward_with_precomputed` /
phases 2-7 of `.forward` (reference src/models/enhanced_fusion.py:397-460, 502-688) on a raw state_dict, with
  multi_domain_frequency.py (DCT :66-196, DWT :203-299, FFT :306-385, AdaptiveBandFusionModule :415-526),
  large_kernel_attention.py (LargeKernelAttention :38-105, LKABlock :112-149, EnhancedCrossBandWithLKA :156-244),
  hierarchical_fusion.py :131-197, fusion_network.py (DynamicExpertSelector :167-236, MultiScaleFeatureExtractor :543-607),
  edge_enhancement.py :182-260.
Structural buffers (DCT basis, zigzag masks, db4 taps, Gaussian kernel) are recomputed here from their definitions.
"""
import math

import torch
import torch.nn.functional as F

DB4_LO = [-0.010597401784997278, 0.032883011666982945, 0.030841381835986965, -0.18703481171888114,
          -0.027983769416983849, 0.63088076792959036, 0.71484657055291582, 0.23037781330885523]
DB4_HI = [-0.23037781330885523, 0.71484657055291582, -0.63088076792959036, -0.027983769416983849,
          0.18703481171888114, 0.030841381835986965, -0.032883011666982945, -0.010597401784997278]


def dct_matrix(n=8):
    k = torch.arange(n, dtype=torch.float64).view(-1, 1)
    i = torch.arange(n, dtype=torch.float64).view(1, -1)
    d = math.sqrt(2.0 / n) * torch.cos(math.pi * k * (2 * i + 1) / (2 * n))
    d[0] = math.sqrt(1.0 / n)
    return d.float()


def zigzag_band(n=8):
    """band id (0 low, 1 mid, 2 high) per coefficient: zigzag rank < n*n//3 -> low, < 2*n*n//3 -> mid."""
    rank = torch.zeros(n, n, dtype=torch.long)
    idx = 0
    for s in range(2 * n - 1):
        rng = range(min(s, n - 1), max(0, s - n + 1) - 1, -1) if s % 2 == 0 else range(max(0, s - n + 1), min(s, n - 1) + 1)
        for i in rng:
            rank[i, s - i] = idx
            idx += 1
    band = torch.full((n, n), 2, dtype=torch.long)
    band[rank < 2 * n * n // 3] = 1
    band[rank < n * n // 3] = 0
    return band


def gaussian_1d(k=5, sigma=1.5):
    c = torch.arange(k, dtype=torch.float32) - k // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False, 0.0, 1e-5)


def _conv(x, sd, p, padding=0, **kw):
    return F.conv2d(x, sd[p + "weight"], sd.get(p + "bias"), padding=padding, **kw)


def _up(x, size):
    return F.interpolate(x, size=size, mode="bilinear", align_corners=False)


# ---------------------------------------------------------------- phase 2: 9-band decomposition
def dct_bands(sd, x):
    B, C, H, W = x.shape
    N = 8
    ph, pw = (N - H % N) % N, (N - W % N) % N
    xp = F.pad(x, (0, pw, 0, ph), mode="reflect") if (ph or pw) else x
    Hp, Wp = xp.shape[-2:]
    D = dct_matrix(N).to(x.device)
    blk = xp.reshape(B, C, Hp // N, N, Wp // N, N).permute(0, 1, 2, 4, 3, 5)
    Y = D @ blk @ D.t()
    band = zigzag_band(N).to(x.device)
    out = []
    for k in range(3):
        sp = D.t() @ (Y * (band == k).float()) @ D
        sp = sp.permute(0, 1, 2, 4, 3, 5).reshape(B, C, Hp, Wp)[:, :, :H, :W]
        out.append(sp * sd["multi_domain_freq.dct.band_scale"][k])
    return out


def dwt_bands(sd, x):
    B, C, H, W = x.shape
    lo, hi = torch.tensor(DB4_LO, device=x.device), torch.tensor(DB4_HI, device=x.device)
    row = lambda f: f.view(1, 1, 1, 8).repeat(C, 1, 1, 1)
    col = lambda f: f.view(1, 1, 8, 1).repeat(C, 1, 1, 1)
    xr = F.pad(x, (7, 7, 0, 0), mode="reflect")
    lo_r, hi_r = F.conv2d(xr, row(lo), stride=(1, 2), groups=C), F.conv2d(xr, row(hi), stride=(1, 2), groups=C)

    def cols(t):
        tp = F.pad(t, (0, 0, 7, 7), mode="reflect")
        return F.conv2d(tp, col(lo), stride=(2, 1), groups=C), F.conv2d(tp, col(hi), stride=(2, 1), groups=C)

    LL, LH = cols(lo_r)
    HL, HH = cols(hi_r)
    return [_up(sb, (H, W)) * sd["multi_domain_freq.dwt.subband_scale"][i] for i, sb in enumerate((LL, LH, HL, HH))]


def fft_mask(sd, H, W):
    m = _up(sd["multi_domain_freq.fft.freq_mask_logits"], (H, W // 2 + 1))
    return torch.sigmoid(m * sd["multi_domain_freq.fft.temperature"].clamp(min=1.0))


def fft_bands(sd, x):
    X = torch.fft.rfft2(x, norm="ortho")
    m = fft_mask(sd, x.shape[-2], x.shape[-1])
    low = torch.fft.irfft2(X * m, s=x.shape[-2:], norm="ortho")
    high = torch.fft.irfft2(X * (1 - m), s=x.shape[-2:], norm="ortho")
    s = sd["multi_domain_freq.fft.band_scale"]
    return [low * s[0], high * s[1]]


def decompose(sd, x):
    return dct_bands(sd, x) + dwt_bands(sd, x) + fft_bands(sd, x)


# ---------------------------------------------------------------- phase 3: cross-band attention + LKA
def lka_block(sd, p, x):
    n1 = _bn(x, sd, p + "norm1.")
    a = F.conv2d(n1, sd[p + "lka.local_conv.weight"], padding=2, groups=x.shape[1])
    a = F.conv2d(a, sd[p + "lka.h_conv.weight"], padding=(0, 10), groups=x.shape[1])
    a = F.conv2d(a, sd[p + "lka.v_conv.weight"], padding=(10, 0), groups=x.shape[1])
    a = torch.sigmoid(_bn(F.conv2d(a, sd[p + "lka.pw_conv.weight"]), sd, p + "lka.bn."))
    x = x + sd[p + "scale1"] * (n1 * a)
    f = _conv(F.gelu(_conv(_bn(x, sd, p + "norm2."), sd, p + "ffn.0.")), sd, p + "ffn.2.")
    return x + sd[p + "scale2"] * f


def cross_band(sd, bands):
    p = "cross_band_attn."
    B, _, H, W = bands[0].shape
    nb, dim, heads = len(bands), 64, 4
    proj = torch.stack([_conv(f, sd, p + "band_proj.") for f in bands], 1)             # B nb dim H W
    flat = proj.permute(0, 3, 4, 1, 2).reshape(B * H * W, nb, dim)
    n = F.layer_norm(flat, (dim,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    qkv = F.linear(n, sd[p + "band_attention.in_proj_weight"], sd[p + "band_attention.in_proj_bias"])
    q, k, v = [t.view(-1, nb, heads, dim // heads).transpose(1, 2) for t in qkv.chunk(3, -1)]
    a = ((q * (dim // heads) ** -0.5) @ k.transpose(-2, -1)).softmax(-1) @ v
    a = a.transpose(1, 2).reshape(-1, nb, dim)
    a = F.linear(a, sd[p + "band_attention.out_proj.weight"], sd[p + "band_attention.out_proj.bias"]) + flat
    a = a.reshape(B, H, W, nb, dim).permute(0, 3, 4, 1, 2)
    return [_conv(lka_block(sd, p + "lka_block.", a[:, i]), sd, p + "out_proj.") + bands[i] for i in range(nb)]


def band_fusion(sd, bands):
    p = "multi_domain_freq.band_fusion."
    imp = torch.cat([F.softplus(sd[p + "dct_importance"]), F.softplus(sd[p + "dwt_importance"]), F.softplus(sd[p + "fft_importance"])])
    imp = imp / (imp.sum() + 1e-8)
    w = [b * torch.sigmoid(_conv(b, sd, p + f"band_attention.{i}.conv.0.", 1)) * imp[i] for i, b in enumerate(bands)]
    cat = torch.cat(w, 1)
    t = _conv(F.gelu(_conv(cat, sd, p + "fusion_transform.0.")), sd, p + "fusion_transform.2.")
    g = torch.sigmoid(_conv(F.gelu(_conv(cat, sd, p + "fusion_gate.0.")), sd, p + "fusion_gate.2."))
    fused = t * g + 0.3 * _conv(torch.cat(bands[:3], 1), sd, p + "dct_residual.")
    return list(fused.chunk(3, 1))


# ---------------------------------------------------------------- phases 5-6
def hierarchical(sd, experts):
    p = "multi_res_fusion."
    st = torch.cat(experts, 1)
    H, W = st.shape[-2:]
    s1, s2 = (max(H // 4, 1), max(W // 4, 1)), (max(H // 2, 1), max(W // 2, 1))

    def stage(x, n):
        x = F.gelu(_conv(F.gelu(_conv(x, sd, p + f"stage{n}_conv.0.", 1)), sd, p + f"stage{n}_conv.2.", 1))
        g = torch.sigmoid(_conv(F.gelu(_conv(x, sd, p + f"stage{n}_gate.gate.0.")), sd, p + f"stage{n}_gate.gate.2."))
        x = x * g
        r = F.conv2d(F.gelu(F.conv2d(x, sd[p + f"stage{n}_res.block.0.weight"], padding=1)), sd[p + f"stage{n}_res.block.2.weight"], padding=1)
        return x + sd[p + f"stage{n}_res.scale"] * r

    f1 = stage(_up(st, s1), 1)
    f1u = _up(f1, s2)
    f2 = stage(torch.cat([f1u, _up(st, s2)], 1), 2) + sd[p + "residual_weight_1_2"] * f1u
    f2u = _up(f2, (H, W))
    f3 = stage(torch.cat([f2u, st], 1), 3) + sd[p + "residual_weight_2_3"] * f2u[:, :32]
    return torch.sigmoid(_conv(F.gelu(_conv(f3, sd, p + "to_rgb.0.", 1)), sd, p + "to_rgb.2.", 1))


def multiscale(sd, x):
    p = "multiscale."
    H, W = x.shape[-2:]
    br = lambda t, n: _bn(F.relu(F.conv2d(t, sd[p + n + ".0.weight"], padding=1)), sd, p + n + ".2.")
    f1 = br(x, "conv_1x")
    f2 = _up(br(F.interpolate(x, scale_factor=0.5, mode="bilinear", align_corners=False), "conv_2x"), (H, W))
    f4 = _up(br(F.interpolate(x, scale_factor=0.25, mode="bilinear", align_corners=False), "conv_4x"), (H, W))
    return F.conv2d(torch.cat([f1, f2, f4], 1), sd[p + "fusion.weight"])


def selector(sd, lr, feat):
    p = "dynamic_selector."
    d = F.relu(_conv(lr, sd, p + "difficulty_estimator.0.", 1))
    d = F.relu(_conv(d, sd, p + "difficulty_estimator.2.", 1))
    d = torch.sigmoid(_conv(d, sd, p + "difficulty_estimator.4.", 1))
    g = torch.sigmoid(_conv(F.relu(_conv(feat, sd, p + "expert_gate.0.", 1)), sd, p + "expert_gate.2."))
    g = torch.sigmoid(10.0 * (g - (0.7 - 0.4 * d)))
    mask = (g >= g.max(1, keepdim=True)[0] * 0.99).float()
    return torch.maximum(g, mask * 0.9), d


def fuse(sd, lr, experts, band_feats):
    H, W = experts[0].shape[-2:]
    mag = [b.abs().mean(1, keepdim=True) for b in band_feats]
    s = mag[0] + mag[1] + mag[2] + 1e-8
    guid = torch.cat([mag[2] / s, mag[1] / s, mag[0] / s], 1)
    fused = hierarchical(sd, experts)
    gh = _up(guid, (H, W))
    st = torch.stack(experts, 1)
    fused = fused * 0.7 + (st * gh.unsqueeze(2)).sum(1) * 0.3
    gates, diff = selector(sd, lr, multiscale(sd, lr))
    gh2, dh = _up(gates, (H, W)), _up(diff, (H, W))
    dyn = sum(e * gh2[:, i:i + 1] for i, e in enumerate(experts)) / (gh2.sum(1, keepdim=True) + 1e-8)
    return fused * (1 - 0.3 * dh) + dyn * (0.3 * dh)


# ---------------------------------------------------------------- phase 7
def edge_refine(sd, sr):
    p = "edge_refine."
    B, C, H, W = sr.shape
    k1 = gaussian_1d().to(sr.device)
    k2 = (k1[:, None] * k1[None, :]).expand(3, 1, 5, 5).contiguous()
    pyr, cur = [], sr
    for lvl in range(3):
        if lvl < 2:
            down = F.avg_pool2d(F.conv2d(cur, k2, padding=2, groups=3), 2, 2)
            pyr.append(cur - _up(down, cur.shape[-2:]))
            cur = down
        else:
            pyr.append(cur)
    lw = F.softmax(sd[p + "level_weights"], 0)
    feats = []
    for lvl, lap in enumerate(pyr):
        q = p + f"edge_refiners.{lvl}."
        o = F.gelu(_conv(lap, sd, q + "conv1.", 1))
        o = F.gelu(_conv(o, sd, q + "conv2.", 1))
        o = _conv(o, sd, q + "conv3.", 1) + _conv(lap, sd, q + "proj.")
        o = o * torch.sigmoid(_conv(F.gelu(_conv(o, sd, q + "attn.attn.0.")), sd, q + "attn.attn.2.", 1))
        if o.shape[-2:] != (H, W):
            o = _up(o, (H, W))
        feats.append(o * lw[lvl])
    edge = _conv(F.gelu(_conv(torch.cat(feats, 1), sd, p + "fusion.0.", 1)), sd, p + "fusion.2.", 1)
    gate = torch.sigmoid(_conv(F.gelu(_conv(torch.cat([sr, edge], 1), sd, p + "edge_gate.0.", 1)), sd, p + "edge_gate.2.", 1))
    return (sr + gate * sd[p + "edge_strength"] * edge).clamp(0, 1)


def refine(sd, fused, lr):
    r = fused
    for i in (0, 2, 4):
        r = F.gelu(_conv(r, sd, f"refine_net.{i}.", 1))
    r = _conv(r, sd, "refine_net.6.", 1)
    fused = fused + 0.1 * r + sd["residual_scale"] * _up(lr, fused.shape[-2:])
    return edge_refine(sd, fused.clamp(0, 1))


def head_forward(sd, lr, experts, return_intermediates=False):
    """experts: [hat, dat, nafnet] SR tensors.  Eval-mode path (collaborative learning skipped, enhanced_fusion.py:733-736)."""
    raw = cross_band(sd, decompose(sd, lr))
    bf = band_fusion(sd, raw)
    fused = fuse(sd, lr, experts, bf)
    out = refine(sd, fused, lr)
    if return_intermediates:
        return out, dict(band_features=bf, fused_before_refine=fused)
    return out
