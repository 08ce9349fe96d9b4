"""ORACLE (test infrastructure only -- never imported by the product path).

Plain-PyTorch fp32 restatement of the reference DAT forward (split_size [8,32], expansion 4) on a raw
state_dict.  Follows /root/reference/src/models/dat/dat_arch.py:
  SpatialAttention :219-342, DynamicPosBias :177-212, AdaptiveSpatialAttention :349-562 (mask :431-489, shift rule
  :426-429), AdaptiveChannelAttention :569-666, SGFN/SpatialGate :103-170, DATB :673-736, ResidualGroup :743-825,
  DAT.forward :996-1028, and ExpertEnsemble.forward_dat (src/models/expert_loader.py:623-652).
Pinned like oracle/hat.py (reference import in the build container + golden fixtures).
"""
import torch
import torch.nn.functional as F

from .hat import MEAN, pad_to_window

C, HEADS = 180, 6
SPLIT = (8, 32)


def rel_index(hs, ws):
    ys, xs = torch.meshgrid(torch.arange(hs), torch.arange(ws), indexing="ij")
    dy = ys.reshape(-1, 1) - ys.reshape(1, -1) + hs - 1
    dx = xs.reshape(-1, 1) - xs.reshape(1, -1) + ws - 1
    return dy * (2 * ws - 1) + dx


def rpe_offsets(hs, ws):
    by, bx = torch.meshgrid(torch.arange(1 - hs, hs), torch.arange(1 - ws, ws), indexing="ij")
    return torch.stack([by.reshape(-1), bx.reshape(-1)], 1).float()


def shift_mask(H, W, hs, ws, sy, sx):
    reg = torch.zeros(H, W)
    cnt = 0
    for a in (slice(0, -hs), slice(-hs, -sy), slice(-sy, None)):
        for b in (slice(0, -ws), slice(-ws, -sx), slice(-sx, None)):
            reg[a, b] = cnt
            cnt += 1
    win = reg.view(H // hs, hs, W // ws, ws).permute(0, 2, 1, 3).reshape(-1, hs * ws)
    d = win.unsqueeze(1) - win.unsqueeze(2)
    return torch.where(d != 0, torch.full_like(d, -100.0), torch.zeros_like(d))


def should_shift(rg, b):
    return (rg % 2 == 0 and b > 0 and (b - 2) % 4 == 0) or (rg % 2 != 0 and b % 4 == 0)


def _ln(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def dyn_pos_bias(sd, p, hs, ws):
    """DynamicPosBias MLP on the fixed offset grid -> [(2hs-1)(2ws-1), heads] (input independent)."""
    x = F.linear(rpe_offsets(hs, ws).to(sd[p + "pos_proj.weight"].device), sd[p + "pos_proj.weight"], sd[p + "pos_proj.bias"])
    for name in ("pos1", "pos2", "pos3"):
        x = F.linear(F.relu(_ln(x, sd[p + name + ".0.weight"], sd[p + name + ".0.bias"])), sd[p + name + ".2.weight"], sd[p + name + ".2.bias"])
    return x


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], False, 0.0, 1e-5)


def _win_attn(q, k, v, H, W, hs, ws, nh, bias, mask):
    B = q.shape[0]

    def part(t):
        t = t.view(B, H // hs, hs, W // ws, ws, nh, -1).permute(0, 1, 3, 5, 2, 4, 6)
        return t.reshape(-1, nh, hs * ws, t.shape[-1])

    qw, kw, vw = part(q), part(k), part(v)
    a = (qw * qw.shape[-1] ** -0.5) @ kw.transpose(-2, -1) + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        a = (a.view(B, nW, nh, hs * ws, hs * ws) + mask.unsqueeze(1).unsqueeze(0)).view(-1, nh, hs * ws, hs * ws)
    o = a.softmax(-1) @ vw
    o = o.view(B, H // hs, W // ws, nh, hs, ws, -1).permute(0, 1, 4, 2, 5, 3, 6)
    return o.reshape(B, H, W, -1)


def _aim_parts(sd, p):
    def chan(t):   # [B,C,H,W] -> [B,C,1,1]
        y = t.mean((2, 3), keepdim=True)
        y = F.conv2d(y, sd[p + "channel_interaction.1.weight"], sd[p + "channel_interaction.1.bias"])
        y = F.gelu(_bn(y, sd, p + "channel_interaction.2."))
        return F.conv2d(y, sd[p + "channel_interaction.4.weight"], sd[p + "channel_interaction.4.bias"])

    def spat(t):   # [B,C,H,W] -> [B,1,H,W]
        y = F.conv2d(t, sd[p + "spatial_interaction.0.weight"], sd[p + "spatial_interaction.0.bias"])
        y = F.gelu(_bn(y, sd, p + "spatial_interaction.1."))
        return F.conv2d(y, sd[p + "spatial_interaction.3.weight"], sd[p + "spatial_interaction.3.bias"])

    def dw(t):
        y = F.conv2d(t, sd[p + "dwconv.0.weight"], sd[p + "dwconv.0.bias"], padding=1, groups=C)
        return F.gelu(_bn(y, sd, p + "dwconv.1."))
    return chan, spat, dw


def spatial_attention(x, H, W, sd, p, rg, bi):
    B, L, _ = x.shape
    qkv = F.linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"]).view(B, H, W, 3, C)
    q, k, v = qkv[..., 0, :], qkv[..., 1, :], qkv[..., 2, :]
    pad = max(SPLIT)
    pb, pr = (pad - H % pad) % pad, (pad - W % pad) % pad
    padf = lambda t: F.pad(t, (0, 0, 0, pr, 0, pb))
    qp, kp, vp = padf(q), padf(k), padf(v)
    Hp, Wp = H + pb, W + pr
    shift = should_shift(rg, bi)
    outs = []
    for br in range(2):
        hs, ws = (SPLIT[0], SPLIT[1]) if br == 0 else (SPLIT[1], SPLIT[0])
        sl = slice(0, C // 2) if br == 0 else slice(C // 2, C)
        sy, sx = hs // 2, ws // 2
        table = dyn_pos_bias(sd, p + f"attns.{br}.pos.", hs, ws)
        bias = table[rel_index(hs, ws).reshape(-1).to(x.device)].view(hs * ws, hs * ws, 3).permute(2, 0, 1)
        qq, kk, vv = qp[..., sl], kp[..., sl], vp[..., sl]
        mask = None
        if shift:
            qq, kk, vv = [torch.roll(t, (-sy, -sx), (1, 2)) for t in (qq, kk, vv)]
            mask = shift_mask(Hp, Wp, hs, ws, sy, sx).to(x)
        o = _win_attn(qq, kk, vv, Hp, Wp, hs, ws, 3, bias, mask)
        if shift:
            o = torch.roll(o, (sy, sx), (1, 2))
        outs.append(o[:, :H, :W].reshape(B, L, C // 2))
    att = torch.cat(outs, 2)
    chan, spat, dw = _aim_parts(sd, p)
    conv_x = dw(v.permute(0, 3, 1, 2))
    cmap = chan(conv_x).permute(0, 2, 3, 1).reshape(B, 1, C)
    smap = spat(att.transpose(1, 2).reshape(B, C, H, W))
    att = att * torch.sigmoid(cmap)
    conv_x = (torch.sigmoid(smap) * conv_x).permute(0, 2, 3, 1).reshape(B, L, C)
    return F.linear(att + conv_x, sd[p + "proj.weight"], sd[p + "proj.bias"])


def channel_attention(x, H, W, sd, p):
    B, N, _ = x.shape
    qkv = F.linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"]).view(B, N, 3, HEADS, C // HEADS).permute(2, 0, 3, 4, 1)
    q, k, v = qkv[0], qkv[1], qkv[2]                    # [B, heads, d, N]
    v_img = v.reshape(B, C, H, W)
    q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
    a = ((q @ k.transpose(-2, -1)) * sd[p + "temperature"]).softmax(-1)
    att = (a @ v).permute(0, 3, 1, 2).reshape(B, N, C)
    chan, spat, dw = _aim_parts(sd, p)
    conv_x = dw(v_img)
    cmap = chan(att.transpose(1, 2).reshape(B, C, H, W))
    smap = spat(conv_x).permute(0, 2, 3, 1).reshape(B, N, 1)
    att = att * torch.sigmoid(smap)
    conv_x = (conv_x * torch.sigmoid(cmap)).permute(0, 2, 3, 1).reshape(B, N, C)
    return F.linear(att + conv_x, sd[p + "proj.weight"], sd[p + "proj.bias"])


def sgfn(x, H, W, sd, p):
    B, N, _ = x.shape
    y = F.gelu(F.linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
    x1, x2 = y.chunk(2, -1)
    x2 = _ln(x2, sd[p + "sg.norm.weight"], sd[p + "sg.norm.bias"])
    hc = x2.shape[-1]
    x2 = F.conv2d(x2.transpose(1, 2).reshape(B, hc, H, W), sd[p + "sg.conv.weight"], sd[p + "sg.conv.bias"], padding=1, groups=hc)
    return F.linear(x1 * x2.flatten(2).transpose(1, 2), sd[p + "fc2.weight"], sd[p + "fc2.bias"])


def dat_forward(sd, x, groups=6, blocks=6):
    B, _, H, W = x.shape
    x = x - MEAN.to(x)
    x0 = F.conv2d(x, sd["conv_first.weight"], sd["conv_first.bias"], padding=1)
    t = _ln(x0.flatten(2).transpose(1, 2), sd["before_RG.1.weight"], sd["before_RG.1.bias"])
    for rg in range(groups):
        res = t
        for bi in range(blocks):
            p = f"layers.{rg}.blocks.{bi}."
            n1 = _ln(t, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
            if bi % 2 == 0:
                t = t + spatial_attention(n1, H, W, sd, p + "attn.", rg, bi)
            else:
                t = t + channel_attention(n1, H, W, sd, p + "attn.")
            t = t + sgfn(_ln(t, sd[p + "norm2.weight"], sd[p + "norm2.bias"]), H, W, sd, p + "ffn.")
        img = t.transpose(1, 2).reshape(B, C, H, W)
        img = F.conv2d(img, sd[f"layers.{rg}.conv.weight"], sd[f"layers.{rg}.conv.bias"], padding=1)
        t = res + img.flatten(2).transpose(1, 2)
    t = _ln(t, sd["norm.weight"], sd["norm.bias"])
    y = t.transpose(1, 2).reshape(B, C, H, W)
    y = F.conv2d(y, sd["conv_after_body.weight"], sd["conv_after_body.bias"], padding=1) + x0
    y = F.leaky_relu(F.conv2d(y, sd["conv_before_upsample.0.weight"], sd["conv_before_upsample.0.bias"], padding=1), 0.01)
    y = F.pixel_shuffle(F.conv2d(y, sd["upsample.0.weight"], sd["upsample.0.bias"], padding=1), 2)
    y = F.pixel_shuffle(F.conv2d(y, sd["upsample.2.weight"], sd["upsample.2.bias"], padding=1), 2)
    return F.conv2d(y, sd["conv_last.weight"], sd["conv_last.bias"], padding=1) + MEAN.to(x)


def forward_dat(sd, x, groups=6, blocks=6):
    h, w = x.shape[-2:]
    return dat_forward(sd, pad_to_window(x), groups, blocks)[:, :, : 4 * h, : 4 * w].clamp(0, 1)
