"""GPU parity tests of the individual ffb200 kernels against plain fp32 PyTorch on the same
(bf16-rounded) operands.  Tolerances are stated per test: operands are bf16, accumulation fp32."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16, F32 = torch.bfloat16, torch.float32


def _dev():
    return torch.device("cuda:0")


def _nhwc(x):  # NCHW -> [pixels, C]
    b, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(b * h * w, c).contiguous()


def _nchw(x2d, b, h, w):
    return x2d.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("kind,cin,cout,H,W,B", [
    ("1x1", 192, 576, 16, 16, 1), ("1x1", 384, 192, 32, 48, 2), ("3x3", 192, 64, 32, 32, 2), ("3x3", 64, 192, 16, 32, 1),
    ("3x3", 64, 256, 32, 32, 1), ("3x3", 64, 16, 32, 32, 1), ("1x1", 1024, 2048, 16, 16, 1), ("2x2s2", 64, 128, 32, 64, 2),
    ("3x3", 128, 32, 64, 64, 1), ("1x1", 64, 128, 8, 16, 3), ("1x1", 64, 3, 16, 32, 2), ("2x2s2", 64, 2, 32, 32, 1),
])
@pytest.mark.parametrize("simt", [0, 1])
def test_conv_gemm_plain(kind, cin, cout, H, W, B, simt):
    from isr2_b200 import ops, packing
    if simt and cin * cout > 192 * 600:
        pytest.skip("SIMT debug loop only checked on the small shapes")
    g = torch.Generator().manual_seed(1)
    k = {"1x1": 1, "3x3": 3, "2x2s2": 2}[kind]
    x = torch.randn(B, cin, H, W, generator=g).to(BF16).float()
    w = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(BF16).float()
    bias = torch.randn(cout, generator=g)
    if kind == "2x2s2":
        ref = F.conv2d(x, w, bias, stride=2)
    else:
        ref = F.conv2d(x, w, bias, padding=k // 2)
    xd = _nhwc(x).to(_dev(), BF16)
    n_pad = (cout + 15) // 16 * 16
    wd = packing.pack_conv(w, n_pad, cin, device=_dev())
    bd = packing.pack_vector(bias, n_pad, device=_dev())
    Ho, Wo = ref.shape[-2:]
    out = torch.zeros(B * Ho * Wo, cout, dtype=F32, device=_dev())
    ops.conv_gemm(xd, B, H, W, cin, wd, kind={"1x1": 0, "3x3": 1, "2x2s2": 2}[kind], n_store=cout, bias=bd, out_f32=out, debug_simt=simt)
    torch.cuda.synchronize()
    got = _nchw(out.cpu(), B, Ho, Wo)
    err = (got - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


def test_conv_gemm_epilogue_chain():
    """bias -> GELU -> alpha -> col_scale -> mul -> aux*chan -> residual -> clamp, bf16 and fp32 stores."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(2)
    B, H, W, cin, cout = 2, 16, 32, 192, 192
    P = B * H * W
    x = torch.randn(P, cin, generator=g).to(BF16)
    w = (torch.randn(cout, cin, generator=g) / math.sqrt(cin)).to(BF16)
    bias, cs = torch.randn(cout, generator=g), torch.rand(cout, generator=g) + 0.5
    mul = torch.randn(P, cout, generator=g).to(BF16)
    aux = torch.randn(P, cout, generator=g).to(BF16)
    chan = torch.rand(B, cout, generator=g)
    res = torch.randn(P, cout, generator=g)
    v = F.gelu(x.float() @ w.float().t() + bias) * 0.7 * cs * mul.float()
    v = v + 0.3 * aux.float() * chan.repeat_interleave(H * W, 0) + res
    ref = v.clamp(0, 1)
    d = _dev()
    o32 = torch.zeros(P, cout, device=d)
    o16 = torch.zeros(P, cout, device=d, dtype=BF16)
    ops.conv_gemm(x.to(d), B, H, W, cin, w.to(d), n_store=cout, bias=bias.to(d), act=ops.ACT_GELU, alpha=0.7, col_scale=cs.to(d),
                  mul=mul.to(d), aux=aux.to(d), aux_chan=chan.to(d), aux_alpha=0.3, res=res.to(d), post_act=ops.ACT_CLAMP01,
                  out_f32=o32, out_bf16=o16)
    torch.cuda.synchronize()
    assert (o32.cpu() - ref).abs().max().item() < 3e-3
    assert (o16.cpu().float() - ref).abs().max().item() < 1e-2


def test_conv_gemm_pixel_shuffle_and_narrow_store():
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(3)
    B, H, W, cin = 1, 16, 16, 64
    x = torch.randn(B, cin, H, W, generator=g).to(BF16).float()
    w = (torch.randn(256, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(BF16).float()
    bias = torch.randn(256, generator=g)
    ref = F.pixel_shuffle(F.conv2d(x, w, bias, padding=1), 2)
    d = _dev()
    rows = packing.pixel_shuffle_rows(256)
    out = torch.zeros(B * 4 * H * W, 64, dtype=BF16, device=d)
    ops.conv_gemm(_nhwc(x).to(d, BF16), B, H, W, cin, packing.pack_conv(w, 256, cin, row_index=rows, device=d), kind=1, n_store=256,
                  bias=packing.pack_vector(bias, 256, index=rows, device=d), pixel_shuffle=2, out_bf16=out)
    torch.cuda.synchronize()
    assert (_nchw(out.cpu().float(), B, 2 * H, 2 * W) - ref).abs().max().item() < 3e-2
    # narrow (3-channel) store into a 12-wide fp32 stack at channel offset 3
    w3 = (torch.randn(3, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(BF16).float()
    b3 = torch.randn(3, generator=g)
    ref3 = F.conv2d(x, w3, b3, padding=1)
    stack = torch.full((B * H * W, 12), 7.0, device=d)
    ops.conv_gemm(_nhwc(x).to(d, BF16), B, H, W, cin, packing.pack_conv(w3, 16, cin, device=d), kind=1, n_store=3,
                  bias=packing.pack_vector(b3, 16, device=d), out_f32=stack[:, 3:])
    torch.cuda.synchronize()
    s = stack.cpu()
    assert (_nchw(s[:, 3:6], B, H, W) - ref3).abs().max().item() < 3e-3
    assert torch.all(s[:, :3] == 7.0) and torch.all(s[:, 6:] == 7.0)


def test_conv_gemm_gate_pairs_and_batch_weights():
    from isr2_b200 import ops
    g = torch.Generator().manual_seed(4)
    B, H, W, cin, c = 2, 16, 16, 64, 64
    P = B * H * W
    x = torch.randn(P, cin, generator=g).to(BF16)
    w = (torch.randn(2 * c, cin, generator=g) / math.sqrt(cin)).to(BF16)     # rows: x1 (c) then x2 (c)
    bias = torch.randn(2 * c, generator=g)
    y = x.float() @ w.float().t() + bias
    ref = y[:, :c] * y[:, c:]
    # interleave rows: chunk g holds x1[8g:8g+8] then x2[8g:8g+8]
    perm = torch.cat([torch.cat([torch.arange(8 * i, 8 * i + 8), c + torch.arange(8 * i, 8 * i + 8)]) for i in range(c // 8)])
    d = _dev()
    out = torch.zeros(P, c, dtype=BF16, device=d)
    ops.conv_gemm(x.to(d), B, H, W, cin, w[perm].contiguous().to(d), n_store=2 * c, bias=bias[perm].contiguous().to(d), gate_pairs=1, out_bf16=out)
    torch.cuda.synchronize()
    assert (out.cpu().float() - ref).abs().max().item() < 5e-2 * max(1.0, ref.abs().max().item() / 4)
    # per-sample weights
    wb = (torch.randn(B, 192, 64, generator=g) / 8).to(BF16)
    refb = torch.cat([x[b * H * W:(b + 1) * H * W].float() @ wb[b].float().t() for b in range(B)])
    o2 = torch.zeros(P, 192, device=d)
    ops.conv_gemm(x.to(d), B, H, W, cin, wb.reshape(B * 192, 64).to(d), n_store=192, w_batch_rows=192, out_f32=o2)
    torch.cuda.synchronize()
    assert (o2.cpu() - refb).abs().max().item() < 3e-3


@pytest.mark.parametrize("mode", ["sa", "sa_shift", "oca", "dat0_shift", "dat1"])
def test_window_attention(mode):
    from isr2_b200 import ops
    from oracle import hat as ohat
    g = torch.Generator().manual_seed(5)
    B, H, W = 2, 32, 64
    heads, hd = 6, 30
    qs, k, v = [torch.randn(B, H, W, heads, hd, generator=g).to(BF16).float() for _ in range(3)]
    qs = qs * 0.3
    q = qs / 1.4426950408889634          # the kernel receives q * log2(e) (bf16) and runs an exp2 softmax
    qkv = torch.zeros(B * H * W, 576, dtype=BF16)
    for i, t in enumerate((qs.to(BF16).float(), k, v)):
        pad = torch.zeros(B, H, W, heads, 32)
        pad[..., :hd] = t
        if i == 2:
            pad[..., 31] = 1.0      # all-ones v column: carries the softmax row sum through the P.V MMA
        qkv[:, i * 192:(i + 1) * 192] = pad.reshape(B * H * W, 192).to(BF16)
    q = qs.to(BF16).float() / 1.4426950408889634
    d = _dev()
    out = torch.zeros(B * H * W, 192, dtype=BF16, device=d)

    def attend(qw, kw_, vw, bias, mask):   # [nW*B, heads, n, d]
        a = qw @ kw_.transpose(-2, -1) + bias.unsqueeze(0)
        if mask is not None:
            nw = mask.shape[0]
            a = (a.view(-1, nw, a.shape[1], a.shape[2], a.shape[3]) + mask.unsqueeze(1).unsqueeze(0)).view(a.shape)
        return a.softmax(-1) @ vw

    def part(t, wh, ww):   # [B,H,W,heads,d] -> [B*nW, heads, wh*ww, d]
        x = t.view(B, H // wh, wh, W // ww, ww, t.shape[3], t.shape[4]).permute(0, 1, 3, 5, 2, 4, 6)
        return x.reshape(-1, t.shape[3], wh * ww, t.shape[4])

    def unpart(o, wh, ww):  # inverse -> [B,H,W,heads,d]
        nh = o.shape[1]
        x = o.view(B, H // wh, W // ww, nh, wh, ww, o.shape[-1]).permute(0, 1, 4, 2, 5, 3, 6)
        return x.reshape(B, H, W, nh, o.shape[-1])

    if mode in ("sa", "sa_shift"):
        table = torch.randn(961, heads, generator=g)
        sh = 8 if mode == "sa_shift" else 0
        roll = (lambda t: torch.roll(t, (-sh, -sh), (1, 2))) if sh else (lambda t: t)
        bias = table[ohat.rpi_sa().reshape(-1)].view(256, 256, heads).permute(2, 0, 1)
        mask = ohat.shift_mask(H, W) if sh else None
        o = unpart(attend(part(roll(q), 16, 16), part(roll(k), 16, 16), part(roll(v), 16, 16), bias, mask), 16, 16)
        ref = torch.roll(o, (sh, sh), (1, 2)) if sh else o
        ops.window_attention(qkv.to(d), B, H, W, out, bias_table=table.t().contiguous().to(d), wh=16, ww=16, shift=(sh, sh))
    elif mode == "oca":
        table = torch.randn(1521, heads, generator=g)
        bias = table[ohat.rpi_oca().reshape(-1)].view(256, 576, heads).permute(2, 0, 1)

        def ext(t):
            x = F.pad(t.permute(0, 3, 4, 1, 2).reshape(B, heads * hd, H, W), (4, 4, 4, 4))
            x = x.unfold(2, 24, 16).unfold(3, 24, 16)  # B C nh nw 24 24
            x = x.permute(0, 2, 3, 4, 5, 1).reshape(-1, 576, heads, hd).permute(0, 2, 1, 3)
            return x
        ref = unpart(attend(part(q, 16, 16), ext(k), ext(v), bias, None), 16, 16)
        ops.window_attention(qkv.to(d), B, H, W, out, bias_table=table.t().contiguous().to(d), wh=16, ww=16, kh=24, kw=24, kpad=(4, 4), rel_sign=-1,
                             rel_off=(-7, -7), rel_stride=39)
    else:
        from oracle import dat as odat
        br = 0 if mode.startswith("dat0") else 1
        wh, ww = (8, 32) if br == 0 else (32, 8)
        sh = (wh // 2, ww // 2) if mode.endswith("shift") else (0, 0)
        table = torch.randn((2 * wh - 1) * (2 * ww - 1), 3, generator=g)
        bias = table[odat.rel_index(wh, ww).reshape(-1)].view(256, 256, 3).permute(2, 0, 1)
        sl = slice(0, 3) if br == 0 else slice(3, 6)
        roll = (lambda t: torch.roll(t, (-sh[0], -sh[1]), (1, 2))) if sh[0] else (lambda t: t)
        mask = odat.shift_mask(H, W, wh, ww, sh[0], sh[1]) if sh[0] else None
        o = unpart(attend(part(roll(q[:, :, :, sl]), wh, ww), part(roll(k[:, :, :, sl]), wh, ww), part(roll(v[:, :, :, sl]), wh, ww), bias, mask), wh, ww)
        o = torch.roll(o, sh, (1, 2)) if sh[0] else o
        ref = torch.zeros(B, H, W, heads, hd)
        ref[:, :, :, sl] = o
        ops.window_attention(qkv.to(d), B, H, W, out, bias_table=table.t().contiguous().to(d), wh=wh, ww=ww, shift=sh, heads=3, head_off=3 * br)
    torch.cuda.synchronize()
    got = out.cpu().float().view(B, H, W, heads, 32)
    assert torch.all(got[..., 30] == 0)
    err = (got[..., :hd] - ref).abs().max().item()
    assert err < 3e-2, f"max abs err {err}"


def test_layernorm_gap_vec_linear():
    from isr2_b200 import ops
    g = torch.Generator().manual_seed(6)
    d = _dev()
    rows, C_ = 1001, 180
    x = torch.randn(rows, 192, generator=g) * 3 + 1
    gam, bet = torch.randn(C_, generator=g), torch.randn(C_, generator=g)
    ref = F.layer_norm(x[:, :C_], (C_,), gam, bet, 1e-5)
    o16 = torch.full((rows, 192), 5.0, dtype=BF16, device=d)
    o32 = torch.full((rows, 192), 5.0, device=d)
    ops.layernorm(x.to(d), rows, C_, gam.to(d), bet.to(d), 1e-5, out_bf16=o16, out_cols=192, out_f32=o32)
    torch.cuda.synchronize()
    assert (o32.cpu()[:, :C_] - ref).abs().max().item() < 1e-4
    assert torch.all(o32.cpu()[:, C_:] == 0) and torch.all(o16.cpu()[:, C_:] == 0)
    assert (o16.cpu().float()[:, :C_] - ref).abs().max().item() < 4e-3 * max(1.0, ref.abs().max().item())     # bf16 half ulp = 2^-9 relative
    # bf16 input with channel offset (DAT SpatialGate norm) and wide rows (NAFNet 1024)
    xb = torch.randn(64, 2048, generator=g).to(BF16)
    g2, b2 = torch.randn(1024, generator=g), torch.randn(1024, generator=g)
    ref2 = F.layer_norm(xb.float()[:, 1024:], (1024,), g2, b2, 1e-6)
    o2 = torch.zeros(64, 1024, device=d)
    ops.layernorm(xb.to(d), 64, 1024, g2.to(d), b2.to(d), 1e-6, out_f32=o2, out_cols=1024, x_off=1024)
    torch.cuda.synchronize()
    assert (o2.cpu() - ref2).abs().max().item() < 1e-4
    # gap + vec_linear
    B, P = 3, 4096
    t = torch.randn(B * P, 192, generator=g).to(BF16)
    outg = torch.zeros(B, 192, device=d)
    scratch = torch.zeros(1, B * 64 * 192, device=d)
    ops.gap(t.to(d), B, P, 180, outg, scratch)
    torch.cuda.synchronize()
    refg = t.float().view(B, P, 192).mean(1)
    assert (outg.cpu()[:, :180] - refg[:, :180]).abs().max().item() < 1e-5
    Wm, bm = torch.randn(6, 180, generator=g), torch.randn(6, generator=g)
    y = torch.zeros(B, 8, device=d)
    ops.vec_linear(outg, B, 180, Wm.to(d), bm.to(d), 6, ops.ACT_RELU, y, y_cols=8)
    torch.cuda.synchronize()
    refy = F.relu(outg.cpu()[:, :180] @ Wm.t() + bm)
    assert (y.cpu()[:, :6] - refy).abs().max().item() < 1e-4 and torch.all(y.cpu()[:, 6:] == 0)


def test_dwconv_and_conv_direct():
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(7)
    d = _dev()
    B, H, W, C_ = 2, 16, 32, 64
    x = torch.randn(B, C_, H, W, generator=g).to(BF16).float()
    for kh, kw in ((3, 3), (5, 5), (1, 21), (21, 1)):
        w = torch.randn(C_, 1, kh, kw, generator=g) / math.sqrt(kh * kw)
        b = torch.randn(C_, generator=g)
        ref = F.gelu(F.conv2d(x, w, b, padding=(kh // 2, kw // 2), groups=C_))
        out = torch.zeros(B * H * W, C_, dtype=BF16, device=d)
        ops.dwconv(_nhwc(x).to(d, BF16), B, H, W, C_, kh, kw, packing.pack_dw(w, C_, device=d), b.to(d), out, act=ops.ACT_GELU)
        torch.cuda.synchronize()
        assert (_nchw(out.cpu().float(), B, H, W) - ref).abs().max().item() < 3e-2
    # SimpleGate mode
    w = torch.randn(C_, 1, 3, 3, generator=g) / 3
    b = torch.randn(C_, generator=g)
    y = F.conv2d(x, w, b, padding=1, groups=C_)
    ref = y[:, :C_ // 2] * y[:, C_ // 2:]
    out = torch.zeros(B * H * W, C_ // 2, dtype=BF16, device=d)
    ops.dwconv(_nhwc(x).to(d, BF16), B, H, W, C_, 3, 3, packing.pack_dw(w, C_, device=d), b.to(d), out, mode=1)
    torch.cuda.synchronize()
    assert (_nchw(out.cpu().float(), B, H, W) - ref).abs().max().item() < 5e-2
    # direct conv, fp32, 3 -> 20 channels, 3x3 and 1x1, from an NCHW image through nchw_to_nhwc
    img = torch.rand(B, 3, H, W, generator=g)
    mean = torch.tensor([0.4, 0.5, 0.6])
    nh = torch.zeros(B * H * W, 4, device=d)
    ops.nchw_to_nhwc(img.to(d), nh, sub=mean.to(d))
    for k in (1, 3):
        w = torch.randn(20, 3, k, k, generator=g)
        b = torch.randn(20, generator=g)
        ref = F.relu(F.conv2d(img - mean.view(1, 3, 1, 1), w, b, padding=k // 2))
        out = torch.zeros(B * H * W, 24, device=d)
        ops.conv_direct(nh, B, H, W, 3, k, packing.pack_conv_direct(w, 24, d), packing.pack_vector(b, 24, device=d), n_store=24, act=ops.ACT_RELU, out_f32=out)
        torch.cuda.synchronize()
        assert (_nchw(out.cpu()[:, :20], B, H, W) - ref).abs().max().item() < 1e-4
    back = torch.zeros(B, 3, H, W, device=d)
    ops.nhwc_to_nchw(nh, 0, 3, back)
    torch.cuda.synchronize()
    assert (back.cpu() - (img - mean.view(1, 3, 1, 1))).abs().max().item() < 1e-6


@pytest.mark.parametrize("cin,n,H,W,B,kind", [(384, 192, 32, 32, 2, 0), (64, 64, 16, 32, 1, 0), (192, 192, 32, 16, 1, 1), (256, 512, 16, 16, 2, 0), (128, 128, 8, 16, 1, 0)])
def test_conv_gemm_residual_tma_epilogue(cin, n, H, W, B, kind):
    """out_f32 = (acc + bias) * alpha * col_scale + res_f32, in place, with the bf16 copy (TMA load/store epilogue)."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(8)
    k = 3 if kind == 1 else 1
    x = torch.randn(B, cin, H, W, generator=g).to(BF16).float()
    w = (torch.randn(n, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(BF16).float()
    bias, cs = torch.randn(n, generator=g), torch.rand(n, generator=g) + 0.5
    res = torch.randn(B * H * W, n, generator=g)
    conv = _nhwc(F.conv2d(x, w, bias, padding=k // 2))
    ref = conv * 0.7 * cs + res
    d = _dev()
    stream = res.clone().to(d)
    o16 = torch.zeros(B * H * W, n, dtype=BF16, device=d)
    ops.conv_gemm(_nhwc(x).to(d, BF16), B, H, W, cin, packing.pack_conv(w, n, cin, device=d), kind=kind, n_store=n, bias=bias.to(d), alpha=0.7,
                  col_scale=cs.to(d), res=stream, out_f32=stream, out_bf16=o16)
    torch.cuda.synchronize()
    assert (stream.cpu() - ref).abs().max().item() < 3e-3
    assert (o16.cpu().float() - ref).abs().max().item() < 3e-2
    # separate output buffer, no bf16 copy, no col_scale
    out = torch.zeros(B * H * W, n, device=d)
    ops.conv_gemm(_nhwc(x).to(d, BF16), B, H, W, cin, packing.pack_conv(w, n, cin, device=d), kind=kind, n_store=n, bias=bias.to(d),
                  res=res.to(d), out_f32=out)
    torch.cuda.synchronize()
    assert (out.cpu() - (conv + res)).abs().max().item() < 3e-3


@pytest.mark.parametrize("C_", [64, 128])
def test_layernorm_narrow_rows(C_):
    from isr2_b200 import ops
    g = torch.Generator().manual_seed(9)
    rows = 4099
    x = torch.randn(rows, C_, generator=g) * 2 + 0.5
    gam, bet = torch.randn(C_, generator=g), torch.randn(C_, generator=g)
    ref = F.layer_norm(x, (C_,), gam, bet, 1e-6)
    d = _dev()
    out = torch.zeros(rows, C_, dtype=BF16, device=d)
    ops.layernorm(x.to(d), rows, C_, gam.to(d), bet.to(d), 1e-6, out_bf16=out, out_cols=C_)
    torch.cuda.synchronize()
    err = (out.cpu().float() - ref).abs()
    assert (err <= 4e-3 * ref.abs() + 1e-3).all(), err.max().item()     # bf16 output rounding only


@pytest.mark.parametrize("cin,n,H,W,B", [(192, 64, 128, 128, 2), (64, 64, 48, 32, 1), (128, 32, 32, 48, 3), (64, 16, 64, 64, 2), (192, 48, 16, 16, 1), (64, 160, 32, 32, 1), (64, 3, 32, 32, 2), (128, 4, 16, 48, 1),
                                            (64, 64, 192, 128, 1), (64, 16, 256, 128, 2), (64, 32, 128, 160, 1)])
@pytest.mark.parametrize("epi", ["store", "gelu", "res", "generic"])
def test_conv_gemm_halo_3x3(cin, n, H, W, B, epi):
    """3x3 convs with <= 64-wide N tiles take the halo-slab variant (A loaded once per 64-channel chunk, nine taps address it):
    every epilogue flavour against F.conv2d on the same bf16 operands, multi-chunk K, multi-tile persistent loops, non-square images;
    the last three cases give every persistent CTA several tiles."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, cin, H, W, generator=g).to(BF16).float()
    w = (torch.randn(n, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(BF16).float()
    bias = torch.randn(n, generator=g)
    conv = _nhwc(F.conv2d(x, w, bias, padding=1))
    d = _dev()
    n_pad = (n + 15) // 16 * 16
    xd, wd, bd = _nhwc(x).to(d, BF16), packing.pack_conv(w, n_pad, cin, device=d), packing.pack_vector(bias, n_pad, device=d)
    P = B * H * W
    if epi == "store":
        out = torch.zeros(P, n_pad, dtype=BF16, device=d)
        ops.conv_gemm(xd, B, H, W, cin, wd, kind=1, n_store=n, bias=bd, out_bf16=out)
        got, ref, tol = out.float().cpu()[:, :n], conv, 3e-2
    elif epi == "gelu":
        out = torch.zeros(P, n_pad, dtype=BF16, device=d)
        ops.conv_gemm(xd, B, H, W, cin, wd, kind=1, n_store=n, bias=bd, act=ops.ACT_GELU, out_bf16=out)
        got, ref, tol = out.float().cpu()[:, :n], F.gelu(conv), 3e-2
    elif epi == "res":
        res = torch.randn(P, n_pad, generator=g)
        stream = res.clone().to(d)
        ops.conv_gemm(xd, B, H, W, cin, wd, kind=1, n_store=n, bias=bd, res=stream, out_f32=stream)
        got, ref, tol = stream.cpu()[:, :n], conv + res[:, :n], 3e-3
    else:
        res = torch.randn(P, n_pad, generator=g).to(BF16)
        out = torch.zeros(P, n_pad, device=d)
        ops.conv_gemm(xd, B, H, W, cin, wd, kind=1, n_store=n, bias=bd, act=ops.ACT_LRELU, res=res.to(d), post_act=ops.ACT_CLAMP01, out_f32=out)
        got, ref, tol = out.cpu()[:, :n], (F.leaky_relu(conv, 0.01) + res.float()[:, :n]).clamp(0, 1), 3e-3
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), f"max abs err {err}"


@pytest.mark.parametrize("B,H,W,C_", [(2, 16, 32, 64), (1, 24, 64, 192), (3, 8, 32, 128), (2, 40, 96, 64), (1, 13, 20, 192), (3, 8, 4, 32), (1, 41, 36, 48)])
def test_dwconv3x3_variants(B, H, W, C_):
    """3x3 depthwise conv, both code paths (TMA-staged tiles when W % 32 == 0, H % 8 == 0 and the channels tile by 64 / 32+32;
    the register kernel otherwise): image borders, multi-tile persistent loops, channel offset in a wider buffer, every
    activation flavour, the fused multiplier and the SimpleGate pair mode, against F.conv2d."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(21)
    d = _dev()
    xw = torch.randn(B, C_ + 16, H, W, generator=g).to(BF16).float()      # the conv reads channels [8, 8 + C_)
    x = xw[:, 8:8 + C_]
    w = torch.randn(C_, 1, 3, 3, generator=g) / 3
    b = torch.randn(C_, generator=g)
    y = F.conv2d(x, w, b, padding=1, groups=C_)
    xd, wd, bd = _nhwc(xw).to(d, BF16), packing.pack_dw(w, C_, device=d), b.to(d)
    mul = torch.randn(B * H * W, C_, generator=g).to(BF16)
    for act, fn in ((ops.ACT_NONE, lambda t: t), (ops.ACT_GELU, F.gelu), (ops.ACT_RELU, F.relu)):
        for m in (None, mul):
            out = torch.zeros(B * H * W, C_, dtype=BF16, device=d)
            ops.dwconv(xd, B, H, W, C_, 3, 3, wd, bd, out, act=act, mul=m.to(d) if m is not None else None, x_off=8)
            torch.cuda.synchronize()
            ref = _nhwc(fn(y)) * (m.float() if m is not None else 1.0)
            err = (out.cpu().float() - ref).abs().max().item()
            assert err < 2e-2 * max(1.0, ref.abs().max().item()), f"act={act} mul={m is not None}: {err}"
    ref = _nhwc(y[:, :C_ // 2] * y[:, C_ // 2:])
    out = torch.zeros(B * H * W, C_ // 2, dtype=BF16, device=d)
    ops.dwconv(xd, B, H, W, C_, 3, 3, wd, bd, out, mode=1, x_off=8)
    torch.cuda.synchronize()
    err = (out.cpu().float() - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), f"gate: {err}"


@pytest.mark.parametrize("B,H,W,C_", [(2, 64, 64, 128), (1, 128, 64, 576), (3, 64, 128, 64), (1, 37, 50, 576), (2, 21, 70, 64), (1, 339, 510, 64)])
@pytest.mark.parametrize("kh,kw", [(5, 5), (1, 21), (21, 1)])
def test_dwconv_large_kernel(B, H, W, C_, kh, kw):
    """The fusion head's large-kernel-attention depthwise chain (5x5, 1x21, 21x1; reference
    src/models/enhanced_fusion.py LKA block) on the TMA-staged sliding-run kernel: image borders on every side, several
    spatial and channel tiles per CTA, with and without bias, against F.conv2d on the same bf16-rounded input."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(100 * kh + kw)
    d = _dev()
    x = torch.randn(B, C_, H, W, generator=g).to(BF16).float()
    w = torch.randn(C_, 1, kh, kw, generator=g) / math.sqrt(kh * kw)
    b = torch.randn(C_, generator=g)
    xd, wd = _nhwc(x).to(d, BF16), packing.pack_dw(w, C_, device=d)
    for bias in (None, b):
        ref = _nhwc(F.conv2d(x, w, bias, padding=(kh // 2, kw // 2), groups=C_))
        out = torch.zeros(B * H * W, C_, dtype=BF16, device=d)
        ops.dwconv(xd, B, H, W, C_, kh, kw, wd, bias.to(d) if bias is not None else None, out)
        torch.cuda.synchronize()
        err = (out.cpu().float() - ref).abs().max().item()
        assert err < 1e-2 * max(1.0, ref.abs().max().item()), f"{kh}x{kw} bias={bias is not None}: {err}"


@pytest.mark.parametrize("kind,cin,H,W,B", [(0, 64, 16, 32, 2), (1, 64, 32, 32, 2), (0, 128, 128, 144, 1), (1, 128, 48, 16, 3)])
@pytest.mark.parametrize("ops_", ["res", "mul", "aux", "res+mul", "res+aux", "res+mul+aux"])
def test_conv_gemm_bf16_operand_epilogue(kind, cin, H, W, B, ops_):
    """out_bf16 = post(act(acc + bias) * alpha * col_scale * mul + aux_alpha * aux * aux_chan + res) with bf16 operand tensors:
    the TMA-operand epilogue (N tile 64), 1x1 and 3x3 (halo and plain main loops), multi-tile persistent loops."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(31)
    n, k = 64, (3 if kind == 1 else 1)
    x = torch.randn(B, cin, H, W, generator=g).to(BF16).float()
    w = (torch.randn(n, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(BF16).float()
    bias, cs = torch.randn(n, generator=g), torch.rand(n, generator=g) + 0.5
    P = B * H * W
    res = torch.randn(P, n, generator=g).to(BF16)
    mul = torch.randn(P, n, generator=g).to(BF16)
    aux = torch.randn(P, 80, generator=g).to(BF16)       # wider buffer: pitch 80, the first 64 columns are the operand
    chan = torch.rand(B, n, generator=g)
    v = torch.sigmoid(_nhwc(F.conv2d(x, w, bias, padding=k // 2))) * 0.7 * cs
    kw = {}
    if "mul" in ops_:
        v = v * mul.float(); kw["mul"] = mul.to(_dev())
    if "aux" in ops_:
        v = v + 0.3 * aux[:, :n].float() * chan.repeat_interleave(H * W, 0); kw.update(aux=aux.to(_dev()), aux_chan=chan.to(_dev()), aux_alpha=0.3)
    if "res" in ops_:
        v = v + res.float(); kw["res"] = res.to(_dev())
    ref = v.clamp(-1.5, 1.5) if False else F.leaky_relu(v, 0.01)
    d = _dev()
    out = torch.zeros(P, n, dtype=BF16, device=d)
    ops.conv_gemm(_nhwc(x).to(d, BF16), B, H, W, cin, packing.pack_conv(w, n, cin, device=d), kind=kind, n_store=n, bias=bias.to(d), act=ops.ACT_SIGMOID,
                  alpha=0.7, col_scale=cs.to(d), post_act=ops.ACT_LRELU, out_bf16=out, **kw)
    torch.cuda.synchronize()
    err = (out.cpu().float() - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


@pytest.mark.parametrize("kind,cin,n,H,W,B", [(0, 128, 256, 32, 32, 2), (1, 64, 256, 16, 48, 1), (0, 256, 512, 16, 16, 1)])
def test_conv_gemm_pixel_shuffle_tma_epilogues(kind, cin, n, H, W, B):
    """PixelShuffle(2) folded into the layer through 5-D tensor maps: bf16 store (+LeakyReLU) and the fp32 residual epilogue
    (out = shuffle(conv) + res, in place, with the bf16 copy) against F.pixel_shuffle."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(41)
    k = 3 if kind == 1 else 1
    x = torch.randn(B, cin, H, W, generator=g).to(BF16).float()
    w = (torch.randn(n, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(BF16).float()
    bias = torch.randn(n, generator=g)
    ref = F.pixel_shuffle(F.conv2d(x, w, bias, padding=k // 2), 2)
    d = _dev()
    rows = packing.pixel_shuffle_rows(n)
    wd, bd = packing.pack_conv(w, n, cin, row_index=rows, device=d), packing.pack_vector(bias, n, index=rows, device=d)
    xd = _nhwc(x).to(d, BF16)
    P4, cq = B * 4 * H * W, n // 4
    out = torch.zeros(P4, cq, dtype=BF16, device=d)
    ops.conv_gemm(xd, B, H, W, cin, wd, kind=kind, n_store=n, bias=bd, act=ops.ACT_LRELU, pixel_shuffle=2, out_bf16=out)
    torch.cuda.synchronize()
    r1 = F.leaky_relu(ref, 0.01)
    assert (_nchw(out.cpu().float(), B, 2 * H, 2 * W) - r1).abs().max().item() < 2e-2 * max(1.0, r1.abs().max().item())
    res = torch.randn(P4, cq, generator=g)
    stream = res.clone().to(d)
    o16 = torch.zeros(P4, cq, dtype=BF16, device=d)
    ops.conv_gemm(xd, B, H, W, cin, wd, kind=kind, n_store=n, bias=bd, pixel_shuffle=2, res=stream, out_f32=stream, out_bf16=o16)
    torch.cuda.synchronize()
    r2 = ref + _nchw(res, B, 2 * H, 2 * W)
    assert (_nchw(stream.cpu(), B, 2 * H, 2 * W) - r2).abs().max().item() < 3e-3 * max(1.0, r2.abs().max().item())
    assert (_nchw(o16.cpu().float(), B, 2 * H, 2 * W) - r2).abs().max().item() < 2e-2 * max(1.0, r2.abs().max().item())


@pytest.mark.parametrize("cin,cout,k,H,W", [(3, 64, 3, 512, 512), (6, 16, 3, 512, 544), (16, 1, 3, 544, 512), (3, 64, 1, 512, 512)])
def test_conv_direct_large_images(cin, cout, k, H, W):
    """fp32 direct conv on HR-sized images (the four-rows-per-thread variant): borders, partial channel groups, bf16 and fp32 stores."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(51)
    d = _dev()
    x = torch.rand(1, cin, H, W, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    b = torch.randn(cout, generator=g)
    ref = _nhwc(F.gelu(F.conv2d(x, w, b, padding=k // 2)))
    cpad = (cout + 7) // 8 * 8
    xin = torch.zeros(H * W, cin + 1, device=d)
    xin[:, :cin] = _nhwc(x).to(d)
    o32 = torch.zeros(H * W, cpad, device=d)
    o16 = torch.zeros(H * W, cpad, device=d, dtype=BF16)
    ops.conv_direct(xin, 1, H, W, cin, k, packing.pack_conv_direct(w, cpad, d), packing.pack_vector(b, cpad, device=d), n_store=cout, act=ops.ACT_GELU, out_f32=o32)
    ops.conv_direct(xin, 1, H, W, cin, k, packing.pack_conv_direct(w, cpad, d), packing.pack_vector(b, cpad, device=d), n_store=cout, act=ops.ACT_GELU, out_bf16=o16)
    torch.cuda.synchronize()
    assert (o32.cpu()[:, :cout] - ref).abs().max().item() < 1e-4
    assert (o16.cpu().float()[:, :cout] - ref).abs().max().item() < 2e-2


def test_conv_gemm_col_sums_and_gap_finalize():
    """Squeeze-excite pool fused into the conv's store epilogue: per-tile column sums + ff_gap_finalize == mean over pixels
    of the fp32 conv output (3x3 halo-free N=192 layer and a 1x1 layer with two N tiles), deterministic across runs."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(61)
    d = _dev()
    for kind, cin, n, B, H, W in ((1, 64, 192, 3, 32, 48), (0, 128, 384, 2, 16, 32)):
        k = 3 if kind == 1 else 1
        x = torch.randn(B, cin, H, W, generator=g).to(BF16).float()
        w = (torch.randn(n, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(BF16).float()
        bias = torch.randn(n, generator=g)
        conv = F.conv2d(x, w, bias, padding=k // 2)
        ref = conv.mean(dim=(2, 3))
        out = torch.zeros(B * H * W, n, dtype=BF16, device=d)
        part = torch.zeros(B * (H * W // 32), n, device=d)
        pooled = [torch.zeros(B, n, device=d) for _ in range(2)]
        for it in range(2):
            ops.conv_gemm(_nhwc(x).to(d, BF16), B, H, W, cin, packing.pack_conv(w, n, cin, device=d), kind=kind, n_store=n, bias=bias.to(d), out_bf16=out, col_sums=part)
            ops.gap_finalize(part, B, H * W // 32, n, 1.0 / (H * W), pooled[it])
        torch.cuda.synchronize()
        assert (pooled[0].cpu() - ref).abs().max().item() < 1e-3
        assert torch.equal(pooled[0], pooled[1])
        assert (_nchw(out.cpu().float(), B, H, W) - conv).abs().max().item() < 3e-2 * max(1.0, conv.abs().max().item())


@pytest.mark.parametrize("B,H,W,ld", [(2, 32, 48, 8), (1, 64, 64, 4), (1, 6, 10, 5)])
def test_gauss_down(B, H, W, ld):
    """Laplacian-pyramid down step (5x5 Gaussian blur with zero padding, then 2x2 mean; reference
    src/models/fusion_network.py LaplacianPyramidRefinement) as one 6x6 stencil, against F.conv2d + avg_pool2d."""
    import ctypes as C_
    from isr2_b200 import lib as L
    g = torch.Generator().manual_seed(5)
    d = _dev()
    x = torch.rand(B, 3, H, W, generator=g)
    k1 = torch.tensor([1.0, 4.0, 6.0, 4.0, 1.0]) / 16.0
    k2 = (k1[:, None] * k1[None, :]).expand(3, 1, 5, 5).contiguous()
    ref = F.avg_pool2d(F.conv2d(x, k2, padding=2, groups=3), 2)
    cur = torch.zeros(B * H * W, ld)
    cur[:, :3] = x.permute(0, 2, 3, 1).reshape(-1, 3)
    cur, k1d = cur.to(d), k1.to(d)
    down = torch.full((B * (H // 2) * (W // 2), 4), 7.0, device=d)
    L.check(L.load().ff_gauss_down(C_.c_void_p(cur.data_ptr()), ld, B, H, W, C_.c_void_p(k1d.data_ptr()), C_.c_void_p(down.data_ptr()), 4, None), "ff_gauss_down")
    torch.cuda.synchronize()
    got = down.cpu()
    assert (got[:, :3] - ref.permute(0, 2, 3, 1).reshape(-1, 3)).abs().max().item() < 1e-6
    assert torch.all(got[:, 3] == 0)


def test_scale_weight_cols_matches_scaled_input():
    """conv(x * s_b) == conv_b(x) with the per-sample weights of ff_scale_weight_cols (NAFNet SCA, nafnet_arch.py:118)."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(9)
    d = _dev()
    B, H, W, c = 3, 16, 32, 64
    x = torch.randn(B * H * W, c, generator=g).to(BF16)
    w = torch.randn(c, c, generator=g) / 8
    s = torch.rand(B, c, generator=g) + 0.5
    wb = torch.zeros(B, c, c, dtype=BF16, device=d)
    ops.scale_weight_cols(w.to(d), s.to(d), wb)
    torch.cuda.synchronize()
    assert torch.equal(wb.cpu(), (w[None] * s[:, None, :]).to(BF16))
    out = torch.zeros(B * H * W, c, dtype=BF16, device=d)
    ops.conv_gemm(x.to(d), B, H, W, c, wb.view(B * c, c), n_store=c, w_batch_rows=c, out_bf16=out)
    torch.cuda.synchronize()
    ref = torch.einsum("bpk,bnk->bpn", x.float().view(B, H * W, c), (w[None] * s[:, None, :]).to(BF16).float()).reshape(-1, c)
    assert (out.cpu().float() - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("mode,act,with_mul", [(0, 1, False), (1, 0, False), (0, 0, True)])
def test_dwconv_pool(mode, act, with_mul):
    """3x3 depthwise conv with the average-pool partials fused in (NAFNet SCA on the SimpleGate output, DAT channel
    interaction on the conv branch): outputs identical to ff_dwconv, pooled mean == mean of the fp32 result, deterministic."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(77 + mode)
    d = _dev()
    B, H, W, C_ = 3, 24, 64, 128
    cout = C_ // 2 if mode == 1 else C_
    x = torch.randn(B, C_, H, W, generator=g).to(BF16).float()
    w = torch.randn(C_, 1, 3, 3, generator=g) / 3
    b = torch.randn(C_, generator=g)
    y = F.conv2d(x, w, b, padding=1, groups=C_)
    mul = torch.randn(B * H * W, cout, generator=g).to(BF16) if with_mul else None
    if mode == 1:
        full = _nhwc(y[:, :cout] * y[:, cout:])
    else:
        full = _nhwc(F.gelu(y) if act == 1 else y) * (mul.float() if with_mul else 1.0)
    ref_mean = full.view(B, H * W, cout).mean(1)
    rows = ops.dwconv_pool_rows(H, W, cout, mode)
    assert rows == ((H + 15) // 16 if mode == 1 else H // 8) * (W // 32)      # SimpleGate tiles are 16 rows tall, the last one partial
    assert ops.dwconv_pool_rows(H, W + 8, cout, mode) == 0
    xd, wd, bd = _nhwc(x).to(d, BF16), packing.pack_dw(w, C_, device=d), b.to(d)
    md = mul.to(d) if with_mul else None
    plain = torch.zeros(B * H * W, cout, dtype=BF16, device=d)
    ops.dwconv(xd, B, H, W, C_, 3, 3, wd, bd, plain, act=act, mode=mode, mul=md)
    pooled = []
    for _ in range(2):
        out = torch.zeros(B * H * W, cout, dtype=BF16, device=d)
        part = torch.full((B * rows, cout), 3.0, device=d)
        mean = torch.zeros(B, cout, device=d)
        ops.dwconv_pool(xd, B, H, W, C_, wd, bd, out, part, act=act, mode=mode, mul=md)
        ops.gap_finalize(part, B, rows, cout, 1.0 / (H * W), mean)
        torch.cuda.synchronize()
        assert torch.equal(out, plain)
        pooled.append(mean)
    assert torch.equal(pooled[0], pooled[1])
    assert (pooled[0].cpu() - ref_mean).abs().max().item() < 2e-3 * max(1.0, ref_mean.abs().max().item())


def test_pack_taps_and_split_bf16_convs():
    """fp32 layers on the tensor cores through split-bf16 operands (ff_pack_taps): the packed rows hold exactly (hi, lo, hi)
    terms / the zero-padded 3x3 neighbourhood, and the resulting convs track fp32 F.conv2d far below bf16 precision."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(123)
    d = _dev()
    # (a) 64 -> 32 3x3 fp32 layer: terms = 3 split, conv_gemm does the 3x3
    B, H, W, cin, n = 2, 32, 48, 64, 32
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(n, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    bias = torch.randn(n, generator=g)
    xr = _nhwc(x).contiguous()
    sp = torch.zeros(B * H * W, 192, dtype=BF16, device=d)
    ops.pack_taps(xr.to(d), B, H, W, cin, 1, 3, sp)
    torch.cuda.synchronize()
    hi = xr.to(BF16)
    lo = (xr - hi.float()).to(BF16)
    assert torch.equal(sp.cpu(), torch.cat([hi, lo, hi], 1))
    out = torch.zeros(B * H * W, n, device=d)
    ops.conv_gemm(sp, B, H, W, 192, packing.pack_conv_split3(w, n, device=d), kind=1, n_store=n, bias=bias.to(d), act=ops.ACT_RELU, out_f32=out)
    torch.cuda.synchronize()
    ref = _nhwc(F.relu(F.conv2d(x.double(), w.double(), bias.double(), padding=1))).float()
    err = (out.cpu() - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), err
    # (b) 3 -> 64 3x3 image layer: im2col (k = 3) with terms = 2, conv_gemm is a 1x1 over one 64-wide k-block
    for (B, H, W) in ((3, 13, 37), (1, 8, 32)):      # partial and exactly-one 8 x 32 tiles of the packing kernel
        img = torch.randn(B, 3, H, W, generator=g)
        rows = torch.full((B * H * W, 4), 7.0)
        rows[:, :3] = _nhwc(img)
        im = torch.full((B * H * W, 64), 5.0, dtype=BF16, device=d)
        ops.pack_taps(rows.to(d), B, H, W, 3, 3, 2, im)
        torch.cuda.synchronize()
        cols = F.unfold(img, 3, padding=1).view(B, 3, 9, H * W).permute(0, 3, 2, 1).reshape(B * H * W, 27)
        chi = cols.to(BF16)
        assert torch.equal(im.cpu(), torch.cat([chi, (cols - chi.float()).to(BF16), torch.zeros(B * H * W, 10, dtype=BF16)], 1))
    B, H, W = 2, 24, 48
    img = torch.rand(B, 3, H, W, generator=g)
    w = torch.randn(64, 3, 3, 3, generator=g) / math.sqrt(27)
    bias = torch.randn(64, generator=g)
    rows = torch.zeros(B * H * W, 4)
    rows[:, :3] = _nhwc(img)
    im = torch.full((B * H * W, 64), 5.0, dtype=BF16, device=d)
    ops.pack_taps(rows.to(d), B, H, W, 3, 3, 2, im)
    torch.cuda.synchronize()
    cols = F.unfold(img, 3, padding=1).view(B, 3, 9, H * W).permute(0, 3, 2, 1).reshape(B * H * W, 27)     # [p][tap][c]
    chi = cols.to(BF16)
    clo = (cols - chi.float()).to(BF16)
    exp = torch.cat([chi, clo, torch.zeros(B * H * W, 10, dtype=BF16)], 1)
    assert torch.equal(im.cpu(), exp)
    out = torch.zeros(B * H * W, 64, device=d)
    ops.conv_gemm(im, B, H, W, 64, packing.pack_conv_im2col2(w, 64, device=d), n_store=64, bias=bias.to(d), out_f32=out)
    torch.cuda.synchronize()
    ref = _nhwc(F.conv2d(img, w.to(BF16).float(), bias, padding=1))
    assert (out.cpu() - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())
    ref32 = _nhwc(F.conv2d(img, w, bias, padding=1))
    assert (out.cpu() - ref32).abs().max().item() < 1e-2


@pytest.mark.parametrize("rows,C_,cols,off", [(1000, 360, 384, 384), (37, 256, 256, 0), (515, 480, 512, 8)])
def test_layernorm_bf16_wide_rows(rows, C_, cols, off):
    """bf16-input LayerNorm over 129..512 columns (DAT SGFN norm on the gated half of the hidden tensor, dat_arch.py:118):
    the 16-lanes-per-row octet kernel against F.layer_norm, padding columns zero, ragged row count."""
    from isr2_b200 import ops
    g = torch.Generator().manual_seed(rows)
    d = _dev()
    x = (torch.randn(rows, off + cols + 8, generator=g) * 2 + 0.5).to(BF16)
    gam, bet = torch.randn(C_, generator=g), torch.randn(C_, generator=g)
    ref = F.layer_norm(x[:, off:off + C_].float(), (C_,), gam, bet, 1e-5)
    out = torch.full((rows, cols), 9.0, dtype=BF16, device=d)
    ops.layernorm(x.to(d), rows, C_, gam.to(d), bet.to(d), 1e-5, out_bf16=out, out_cols=cols, x_off=off)
    torch.cuda.synchronize()
    got = out.cpu().float()
    assert ((got[:, :C_] - ref).abs() <= 4e-3 * ref.abs() + 1e-3).all()
    assert torch.all(got[:, C_:] == 0)


@pytest.mark.parametrize("cin,n,c_real,H,W,B,kind,aux,with_bf16", [
    (384, 192, 180, 32, 32, 2, 0, False, False), (192, 192, 180, 32, 16, 1, 1, False, True), (192, 192, 180, 16, 32, 2, 0, True, False),
    (64, 64, 64, 16, 32, 1, 0, False, False), (128, 128, 128, 24, 16, 3, 0, False, True), (256, 256, 256, 8, 16, 2, 0, False, False)])
def test_conv_gemm_fused_layernorm(cin, n, c_real, H, W, B, kind, aux, with_bf16):
    """FFConvGemm.ln_*: the residual epilogue (EPI_RES / EPI_RES_AUX) also emits LayerNorm(new residual row) as bf16 -- the row is
    re-read from TMEM after both warps of a lane quadrant exchanged their partial sums.  Checked against F.layer_norm of the fp32
    result; padding columns (c_real..n) carry zero weights / bias / residual / gamma / beta and must come out exactly 0."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(19)
    k = 3 if kind == 1 else 1
    P = B * H * W
    x = torch.randn(B, cin, H, W, generator=g).to(BF16).float()
    w = torch.zeros(n, cin, k, k)
    w[:c_real] = (torch.randn(c_real, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(BF16).float()
    bias, res = torch.zeros(n), torch.zeros(P, n)
    bias[:c_real] = torch.randn(c_real, generator=g)
    res[:, :c_real] = torch.randn(P, c_real, generator=g) * 2 + 0.7        # non-zero row mean
    gamma, beta = torch.zeros(n), torch.zeros(n)
    gamma[:c_real], beta[:c_real] = 1 + 0.2 * torch.randn(c_real, generator=g), 0.1 * torch.randn(c_real, generator=g)
    ref = _nhwc(F.conv2d(x, w, bias, padding=k // 2)) + res
    d = _dev()
    kw = {}
    if aux:
        av = torch.zeros(P, n)
        av[:, :c_real] = torch.randn(P, c_real, generator=g)
        av = av.to(BF16)
        chan = torch.rand(B, n, generator=g)
        ref = ref + 0.3 * av.float() * chan.repeat_interleave(H * W, 0)
        kw = dict(aux=av.to(d), aux_chan=chan.to(d), aux_alpha=0.3)
    ln_ref = torch.zeros(P, n)
    ln_ref[:, :c_real] = F.layer_norm(ref[:, :c_real], (c_real,), gamma[:c_real], beta[:c_real], 1e-5)
    stream = res.clone().to(d)
    lno = torch.full((P, n), 7.0, dtype=BF16, device=d)
    o16 = torch.zeros(P, n, dtype=BF16, device=d) if with_bf16 else None
    ops.conv_gemm(_nhwc(x).to(d, BF16), B, H, W, cin, packing.pack_conv(w, n, cin, device=d), kind=kind, n_store=n, bias=bias.to(d),
                  res=stream, out_f32=stream, out_bf16=o16, ln=(gamma.to(d), beta.to(d), 1e-5, c_real, lno), **kw)
    torch.cuda.synchronize()
    assert (stream.cpu() - ref).abs().max().item() < 3e-3
    e = (lno.cpu().float() - ln_ref).abs().max().item()
    assert e < 3e-2, e                                      # bf16 rounding of values up to ~4
    assert (lno.cpu().float()[:, c_real:] == 0).all()
    if with_bf16:
        assert (o16.cpu().float() - ref).abs().max().item() < 4e-2
    # rows wider than one n tile are rejected loudly
    from isr2_b200 import lib
    with pytest.raises(lib.FFError):
        wide = torch.zeros(P, 512, device=d)
        ops.conv_gemm(_nhwc(x).to(d, BF16), B, H, W, cin, packing.pack_conv(torch.zeros(512, cin, k, k), 512, cin, device=d), kind=kind, n_store=512,
                      res=wide, out_f32=wide, ln=(torch.zeros(512, device=d), torch.zeros(512, device=d), 1e-5, 512, torch.zeros(P, 512, dtype=BF16, device=d)))


@pytest.mark.parametrize("kind,cin,n,H,W,B", [(0, 64, 64, 13, 20, 2), (1, 64, 64, 21, 27, 1), (1, 192, 192, 9, 33, 2), (0, 128, 256, 5, 7, 3),
                                               (1, 64, 16, 30, 44, 1), (2, 64, 128, 26, 38, 2), (0, 192, 576, 19, 50, 1)])
@pytest.mark.parametrize("epi", ["store", "gelu", "res", "generic"])
def test_conv_gemm_partial_edge_tiles(kind, cin, n, H, W, B, epi):
    """Sizes that are not multiples of the 8x16 (16x8 halo) output tile: TMA zero-fills the loads and clips the stores of the
    partial edge tiles, the direct-store epilogues mask by coordinates -- nothing may leak into neighbouring rows / samples."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(23)
    k = {0: 1, 1: 3, 2: 2}[kind]
    x = torch.randn(B, cin, H, W, generator=g).to(BF16).float()
    w = (torch.randn(n, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(BF16).float()
    bias = torch.randn(n, generator=g)
    conv = F.conv2d(x, w, bias, stride=2) if kind == 2 else F.conv2d(x, w, bias, padding=k // 2)
    Ho, Wo = conv.shape[-2:]
    conv = _nhwc(conv)
    P = B * Ho * Wo
    d = _dev()
    xd, wd, bd = _nhwc(x).to(d, BF16), packing.pack_conv(w, n, cin, device=d), bias.to(d)
    guard = 64      # rows after the tensor that must stay untouched
    if epi in ("store", "gelu"):
        out = torch.full((P + guard, n), 9.0, dtype=BF16, device=d)
        ops.conv_gemm(xd, B, H, W, cin, wd, kind=kind, n_store=n, bias=bd, act=ops.ACT_GELU if epi == "gelu" else ops.ACT_NONE, out_bf16=out[:P])
        ref = F.gelu(conv) if epi == "gelu" else conv
        torch.cuda.synchronize()
        assert (out[:P].cpu().float() - ref).abs().max().item() < 4e-2
        assert (out[P:] == 9.0).all()
    elif epi == "res":
        res = torch.randn(P, n, generator=g)
        stream = torch.full((P + guard, n), 9.0, device=d)
        stream[:P] = res.to(d)
        ops.conv_gemm(xd, B, H, W, cin, wd, kind=kind, n_store=n, bias=bd, res=stream[:P], out_f32=stream[:P])
        torch.cuda.synchronize()
        assert (stream[:P].cpu() - (conv + res)).abs().max().item() < 4e-3
        assert (stream[P:] == 9.0).all()
    else:
        mul = torch.randn(P, n, generator=g).to(BF16)
        res = torch.randn(P, n, generator=g).to(BF16)
        out = torch.full((P + guard, n), 9.0, device=d)
        ops.conv_gemm(xd, B, H, W, cin, wd, kind=kind, n_store=n, bias=bd, act=ops.ACT_SIGMOID, mul=mul.to(d), res=res.to(d), post_act=ops.ACT_RELU, out_f32=out[:P])
        torch.cuda.synchronize()
        ref = F.relu(torch.sigmoid(conv) * mul.float() + res.float())
        assert (out[:P].cpu() - ref).abs().max().item() < 4e-3
        assert (out[P:] == 9.0).all()


def test_conv_gemm_cropped_narrow_store_and_unaligned_pixel_shuffle():
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(24)
    d = _dev()
    # last conv of an expert run on a padded image: 3 output channels, top-left crop of the output, residual in the padded geometry
    B, cin, H, W, Hc, Wc = 2, 64, 32, 48, 29, 41
    x = torch.randn(B, cin, H, W, generator=g).to(BF16).float()
    w = (torch.randn(3, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(BF16).float()
    bias = torch.randn(3, generator=g)
    res = torch.randn(B * H * W, 4, generator=g)
    ref = (F.conv2d(x, w, bias, padding=1) + _nchw(res[:, :3], B, H, W)).clamp(0, 1)[:, :, :Hc, :Wc]
    out = torch.full((B * Hc * Wc + 32, 12), 9.0, device=d)
    ops.conv_gemm(_nhwc(x).to(d, BF16), B, H, W, cin, packing.pack_conv(w, 16, cin, device=d), kind=1, n_store=3, bias=packing.pack_vector(bias, 16, device=d),
                  res=res.to(d), post_act=ops.ACT_CLAMP01, out_f32=out[:B * Hc * Wc, 3:], out_crop=(Hc, Wc))
    torch.cuda.synchronize()
    got = _nchw(out[:B * Hc * Wc, 3:6].cpu(), B, Hc, Wc)
    assert (got - ref).abs().max().item() < 4e-3
    assert (out[B * Hc * Wc:] == 9.0).all() and (out[:, :3] == 9.0).all() and (out[:, 6:] == 9.0).all()
    # PixelShuffle(2) layer with a residual at a size whose rows do not fill the 8-row tiles, two samples (per-sample launches)
    B, cin, n, H, W = 2, 128, 256, 5, 12
    x = torch.randn(B, cin, H, W, generator=g).to(BF16).float()
    w = (torch.randn(n, cin, 1, 1, generator=g) / math.sqrt(cin)).to(BF16).float()
    skip = torch.randn(B * 4 * H * W, n // 4, generator=g)
    ref = F.pixel_shuffle(F.conv2d(x, w), 2) + _nchw(skip, B, 2 * H, 2 * W)
    rows = packing.pixel_shuffle_rows(n)
    stream = skip.clone().to(d)
    ops.conv_gemm(_nhwc(x).to(d, BF16), B, H, W, cin, packing.pack_matrix(w.reshape(n, cin), n, cin, row_index=rows, device=d), n_store=n, pixel_shuffle=2,
                  res=stream, out_f32=stream)
    torch.cuda.synchronize()
    assert (_nchw(stream.cpu(), B, 2 * H, 2 * W) - ref).abs().max().item() < 4e-3


@pytest.mark.parametrize("cin,cout,k,H,W,B", [(3, 64, 3, 37, 50, 2), (27, 16, 3, 19, 23, 1), (64, 64, 1, 9, 12, 1), (8, 1, 3, 148, 200, 1)])
def test_conv_direct_any_size(cin, cout, k, H, W, B):
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(25)
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    bias = torch.randn(cout, generator=g)
    ref = _nhwc(F.relu(F.conv2d(x, w, bias, padding=k // 2)))
    d = _dev()
    cp = (cout + 7) // 8 * 8
    out = torch.full((B * H * W + 16, cp), 9.0, device=d)
    ops.conv_direct(_nhwc(x).to(d), B, H, W, cin, k, packing.pack_conv_direct(w, cp, d), packing.pack_vector(bias, cp, device=d), n_store=cout, act=ops.ACT_RELU,
                    out_f32=out[:B * H * W])
    torch.cuda.synchronize()
    assert (out[:B * H * W, :cout].cpu() - ref).abs().max().item() < 1e-4
    assert (out[B * H * W:] == 9.0).all()


@pytest.mark.parametrize("cin,cout,k,in_ld,out_ld,in_bf16,out_bf16,act", [
    (8, 1, 3, 8, 1, False, False, "sigmoid"), (16, 1, 3, 16, 1, False, False, "sigmoid"), (6, 16, 3, 8, 16, False, False, "gelu"),
    (6, 16, 3, 6, 16, False, False, "gelu"), (32, 8, 1, 64, 8, True, False, "gelu"), (3, 64, 1, 4, 64, False, True, "none"),
    (64, 64, 1, 64, 64, False, False, "none"), (64, 64, 1, 68, 64, False, False, "gelu")])
@pytest.mark.parametrize("B,H,W", [(2, 37, 50), (1, 64, 96)])
def test_conv_direct_specialised_small_layers(cin, cout, k, in_ld, out_ld, in_bf16, out_bf16, act, B, H, W):
    """The specialised forms ff_conv_direct dispatches to for the edge refiner's small layers at output resolution (3x3 8->1, 16->1,
    6->16 in fp32; 1x1 32 bf16 -> 8; 1x1 3 -> 64 bf16) and for the fp32 64 -> 64 mixers of the routing path, on whole and partial tiles, with channel pitches wider than the channel count,
    against F.conv2d."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(61 + cin)
    x = torch.randn(B, cin, H, W, generator=g)
    if in_bf16:
        x = x.to(BF16).float()
    w = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    bias = torch.randn(cout, generator=g)
    y = F.conv2d(x, w, bias, padding=k // 2)
    y = {"sigmoid": torch.sigmoid, "gelu": F.gelu, "none": lambda v: v}[act](y)
    ref = _nhwc(y)
    d = _dev()
    P = B * H * W
    xin = torch.full((P, in_ld), 3.0)
    xin[:, :cin] = _nhwc(x)
    xin = xin.to(d, BF16 if in_bf16 else F32)
    cp = (cout + 7) // 8 * 8
    out = torch.full((P + 16, out_ld), 9.0, device=d, dtype=BF16 if out_bf16 else F32)
    code = {"sigmoid": ops.ACT_SIGMOID, "gelu": ops.ACT_GELU, "none": ops.ACT_NONE}[act]
    kw = dict(out_bf16=out[:P]) if out_bf16 else dict(out_f32=out[:P])
    ops.conv_direct(xin, B, H, W, cin, k, packing.pack_conv_direct(w, cp, d), packing.pack_vector(bias, cp, device=d), n_store=cout, act=code, **kw)
    torch.cuda.synchronize()
    e = (out[:P, :cout].float().cpu() - ref).abs().max().item()
    assert e < (3e-2 if out_bf16 else 2e-5), e
    assert (out[P:] == 9.0).all()


@pytest.mark.parametrize("B,H,W,with_bf16,with_ln", [(2, 32, 32, False, True), (1, 16, 48, True, False), (3, 24, 16, True, True), (1, 128, 128, False, True)])
def test_mlp_fused(B, H, W, with_bf16, with_ln):
    """ff_mlp_fused: x += fc2(GELU(fc1(t))) with the hidden tile on chip (csrc/mlp_fused.cu), optional bf16 copy and fused LayerNorm
    of the new row, against plain fp32 PyTorch on the bf16-rounded operands (180 / 360 real channels inside 192 / 384 padded)."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(31)
    P, C, HD, CP = B * H * W, 180, 360, 192
    t = torch.zeros(P, CP)
    t[:, :C] = torch.randn(P, C, generator=g)
    t = t.to(BF16)
    w1 = (torch.randn(HD, C, generator=g) / math.sqrt(C)).to(BF16).float()
    w2 = (torch.randn(C, HD, generator=g) / math.sqrt(HD)).to(BF16).float()
    b1, b2 = torch.randn(HD, generator=g) * 0.3, torch.randn(C, generator=g) * 0.3
    x = torch.zeros(P, CP)
    x[:, :C] = torch.randn(P, C, generator=g) * 2 + 0.5
    hid = F.gelu(t.float()[:, :C] @ w1.t() + b1)
    ref = x.clone()
    ref[:, :C] += hid.to(BF16).float() @ w2.t() + b2
    gamma, beta = torch.zeros(CP), torch.zeros(CP)
    gamma[:C], beta[:C] = 1 + 0.2 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    ln_ref = torch.zeros(P, CP)
    ln_ref[:, :C] = F.layer_norm(ref[:, :C], (C,), gamma[:C], beta[:C], 1e-5)
    d = _dev()
    xd = x.clone().to(d)
    o16 = torch.full((P, CP), 7.0, dtype=BF16, device=d) if with_bf16 else None
    lno = torch.full((P, CP), 7.0, dtype=BF16, device=d) if with_ln else None
    ops.mlp_fused(t.to(d), B, H, W, packing.pack_matrix(w1, 2 * CP, CP, device=d), packing.pack_vector(b1, 2 * CP, device=d),
                  packing.pack_matrix(w2, CP, 2 * CP, device=d), packing.pack_vector(b2, CP, device=d), xd, out_bf16=o16,
                  ln=(gamma.to(d), beta.to(d), 1e-5, C, lno) if with_ln else None)
    torch.cuda.synchronize()
    e = (xd.cpu() - ref).abs().max().item()
    assert e < 2e-2, e                                  # bf16 hidden activations (values up to ~3) through a 360-long contraction
    assert (xd.cpu()[:, C:] == 0).all()
    if with_bf16:
        assert (o16.cpu().float() - ref).abs().max().item() < 6e-2
    if with_ln:
        assert (lno.cpu().float() - ln_ref).abs().max().item() < 4e-2
        assert (lno.cpu().float()[:, C:] == 0).all()


@pytest.mark.parametrize("B,H,W,with_cab,with_bf16,with_ln,inplace", [(2, 16, 32, True, False, True, False), (1, 24, 40, True, False, True, True),
                                                                  (3, 64, 64, True, False, True, True), (2, 13, 21, False, True, False, True),
                                                                  (5, 48, 80, False, True, True, False)])
def test_hab_tail(B, H, W, with_cab, with_bf16, with_ln, inplace):
    """ff_hab_tail (csrc/hab_tail.cu): proj + shortcut (+ 0.01 * cab * se through the diagonal K block) -> LayerNorm2 -> MLP -> residual
    (-> next LayerNorm) as one kernel with x1 in TMEM, against plain fp32 PyTorch on the bf16-rounded operands, on whole and partial tiles,
    several tiles per CTA (B=3, 64x64 = 96 tiles... per-SM queues of one; B=5, 48x80 = 150 tiles > 148 SMs)."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(53 + B)
    P, C, HD, CP = B * H * W, 180, 360, 192
    def padded(v):
        o = torch.zeros(P, CP)
        o[:, :C] = v
        return o
    att = padded(torch.randn(P, C, generator=g)).to(BF16)
    cab = padded(torch.randn(P, C, generator=g) * 3).to(BF16)
    wp = (torch.randn(C, C, generator=g) / math.sqrt(C)).to(BF16).float()
    bp = torch.randn(C, generator=g) * 0.3
    se = torch.rand(B, C, generator=g)
    res = padded(torch.randn(P, C, generator=g) * 2 + 0.5)
    w1 = (torch.randn(HD, C, generator=g) / math.sqrt(C)).to(BF16).float()
    w2 = (torch.randn(C, HD, generator=g) / math.sqrt(HD)).to(BF16).float()
    b1, b2 = torch.randn(HD, generator=g) * 0.3, torch.randn(C, generator=g) * 0.3
    def lnp():
        ga, be = torch.zeros(CP), torch.zeros(CP)
        ga[:C], be[:C] = 1 + 0.2 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
        return ga, be
    g2, be2 = lnp()
    gn, ben = lnp()
    x1 = res[:, :C] + att.float()[:, :C] @ wp.t() + bp
    if with_cab:
        x1 = x1 + (0.01 * se).to(BF16).float().repeat_interleave(H * W, 0) * cab.float()[:, :C]
    t = F.layer_norm(x1, (C,), g2[:C], be2[:C], 1e-5).to(BF16).float()
    hid = F.gelu(t @ w1.t() + b1)
    ref = padded(x1 + hid.to(BF16).float() @ w2.t() + b2)
    ln_ref = padded(F.layer_norm(ref[:, :C], (C,), gn[:C], ben[:C], 1e-5))
    d = _dev()
    wpd = packing.pack_matrix(wp, CP, CP, device=d)
    if with_cab:
        wcat = torch.empty(B, CP, 2 * CP, dtype=BF16, device=d)
        sed = torch.zeros(B, CP, device=d)
        sed[:, :C] = se.to(d)
        ops.build_concat_diag_weights(wpd, sed, 0.01, wcat)
        wuse, wbr = wcat.view(B * CP, 2 * CP), CP
    else:
        wuse, wbr = wpd, 0
    resd = res.clone().to(d)
    xd = resd if inplace else torch.full((P, CP), 9.0, device=d)
    o16 = torch.full((P, CP), 7.0, dtype=BF16, device=d) if with_bf16 else None
    lno = torch.full((P, CP), 7.0, dtype=BF16, device=d) if with_ln else None
    ops.hab_tail(att.to(d), B, H, W, wuse, packing.pack_vector(bp, CP, device=d), resd, (g2.to(d), be2.to(d)),
                 packing.pack_matrix(w1, 2 * CP, CP, device=d), packing.pack_vector(b1, 2 * CP, device=d),
                 packing.pack_matrix(w2, CP, 2 * CP, device=d), packing.pack_vector(b2, CP, device=d), xd,
                 a1=cab.to(d) if with_cab else None, wp_batch_rows=wbr, out_bf16=o16, ln=(gn.to(d), ben.to(d), lno) if with_ln else None)
    torch.cuda.synchronize()
    e = (xd.cpu() - ref).abs().max().item()
    assert e < 3e-2, e              # bf16 LayerNorm2 output and hidden activations through 180 / 360-long contractions
    if with_cab:
        # the same block with the diagonal K block generated inside the kernel from the per-sample scale (FFHabTail.a1_diag): identical
        res2 = res.clone().to(d)
        x2 = res2 if inplace else torch.full((P, CP), 9.0, device=d)
        lno2 = torch.full((P, CP), 7.0, dtype=BF16, device=d) if with_ln else None
        ops.hab_tail(att.to(d), B, H, W, wpd, packing.pack_vector(bp, CP, device=d), res2, (g2.to(d), be2.to(d)),
                     packing.pack_matrix(w1, 2 * CP, CP, device=d), packing.pack_vector(b1, 2 * CP, device=d),
                     packing.pack_matrix(w2, CP, 2 * CP, device=d), packing.pack_vector(b2, CP, device=d), x2,
                     a1=cab.to(d), a1_diag=sed, a1_alpha=0.01, ln=(gn.to(d), ben.to(d), lno2) if with_ln else None)
        torch.cuda.synchronize()
        assert torch.equal(x2, xd)
        if with_ln:
            assert torch.equal(lno2, lno)
    assert (xd.cpu()[:, C:] == 0).all()
    if with_bf16:
        assert (o16.cpu().float() - ref).abs().max().item() < 8e-2
    if with_ln:
        assert (lno.cpu().float() - ln_ref).abs().max().item() < 5e-2
        assert (lno.cpu().float()[:, C:] == 0).all()


@pytest.mark.parametrize("B,H,W,batch_w,mode,inplace", [(2, 16, 32, True, "ln", True), (1, 24, 40, False, "ln", False), (3, 64, 64, True, "bf16", True),
                                                         (2, 13, 21, True, "none", True), (5, 96, 160, True, "ln", True)])
def test_naf_tail(B, H, W, batch_w, mode, inplace):
    """ff_naf_tail (csrc/naf_tail.cu): conv3 (per-sample weights) + residual -> LayerNorm2d -> conv4 -> SimpleGate -> conv5 + residual
    (-> next LayerNorm2d / bf16 copy) of a 64-channel NAFBlock as one kernel, against plain fp32 PyTorch on the bf16-rounded operands;
    whole and partial tiles, more tiles than resident CTAs (B=5, 96x160 = 600 tiles > 2 x 148)."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(71 + B)
    P, C = B * H * W, 64
    gt = torch.randn(P, C, generator=g).to(BF16)
    w3 = (torch.randn(B if batch_w else 1, C, C, generator=g) / math.sqrt(C)).to(BF16)
    b3 = torch.randn(C, generator=g) * 0.3
    res = torch.randn(P, C, generator=g) * 2 + 0.5
    w4 = (torch.randn(2 * C, C, generator=g) / math.sqrt(C)).to(BF16)
    b4 = torch.randn(2 * C, generator=g) * 0.3
    w5 = (torch.randn(C, C, generator=g) / math.sqrt(C)).to(BF16)
    b5 = torch.randn(C, generator=g) * 0.3
    g2, be2 = 1 + 0.2 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    gn, ben = 1 + 0.2 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    w3p = w3.float().repeat_interleave(H * W, 0) if batch_w else w3.float().expand(P, C, C)
    y = res + torch.einsum("pk,pnk->pn", gt.float(), w3p) + b3
    t = F.layer_norm(y, (C,), g2, be2, 1e-6).to(BF16).float()
    u = t @ w4.float().t() + b4
    gate = (u[:, :C] * u[:, C:]).to(BF16).float()
    ref = y + gate @ w5.float().t() + b5
    ln_ref = F.layer_norm(ref, (C,), gn, ben, 1e-6)
    d = _dev()
    resd = res.clone().to(d)
    xd = resd if inplace else torch.full((P, C), 9.0, device=d)
    o16 = torch.full((P, C), 7.0, dtype=BF16, device=d) if mode != "none" else None
    ops.naf_tail(gt.to(d), B, H, W, w3.reshape(-1, C).contiguous().to(d), b3.to(d), resd, (g2.to(d), be2.to(d)), w4.to(d), b4.to(d), w5.to(d), b5.to(d), xd,
                 w3_batch_rows=C if batch_w else 0, out_bf16=o16, ln=(gn.to(d), ben.to(d)) if mode == "ln" else None)
    torch.cuda.synchronize()
    e = (xd.cpu() - ref).abs().max().item()
    assert e < 3e-2, e              # bf16 LayerNorm2d output and gate products (values up to ~10) through 64-long contractions
    if not inplace:
        assert torch.equal(resd.cpu(), res)
    if mode == "ln":
        assert (o16.cpu().float() - ln_ref).abs().max().item() < 5e-2
    elif mode == "bf16":
        assert (o16.cpu().float() - ref).abs().max().item() < 8e-2


def test_conv_gemm_k_concatenated_second_operand():
    """FFConvGemm.x2: out = [x | x2] . W^T with per-sample weights [W | diag(alpha * s_b)] (ff_build_concat_diag_weights) reproduces the
    aux epilogue  x.W^T + bias + alpha * s_b[n] * x2[p, n] + res  of HAT's proj layer (hat_arch.py:306) on the tensor pipe."""
    from isr2_b200 import ops, packing
    g = torch.Generator().manual_seed(41)
    B, H, W, C = 3, 16, 32, 192
    P = B * H * W
    x = torch.randn(P, C, generator=g).to(BF16)
    x2 = torch.randn(P, C, generator=g).to(BF16)
    w = (torch.randn(C, C, generator=g) / math.sqrt(C)).to(BF16)
    bias, res = torch.randn(C, generator=g), torch.randn(P, C, generator=g)
    s = torch.rand(B, C, generator=g)
    ref = x.float() @ w.float().t() + bias + (0.01 * s).to(BF16).float().repeat_interleave(H * W, 0) * x2.float() + res
    d = _dev()
    wcat = torch.empty(B, C, 2 * C, dtype=BF16, device=d)
    ops.build_concat_diag_weights(w.to(d), s.to(d), 0.01, wcat)
    assert torch.equal(wcat[:, :, :C].cpu(), w.expand(B, C, C)) and torch.equal(torch.diagonal(wcat[1, :, C:].cpu().float()), (0.01 * s[1]).to(BF16).float())
    stream = res.clone().to(d)
    lno = torch.zeros(P, C, dtype=BF16, device=d)
    gam, bet = torch.ones(C, device=d), torch.zeros(C, device=d)
    ops.conv_gemm(x.to(d), B, H, W, C, wcat.view(B * C, 2 * C), n_store=C, w_batch_rows=C, bias=bias.to(d), x2=x2.to(d), res=stream, out_f32=stream,
                  ln=(gam, bet, 1e-5, C, lno))
    torch.cuda.synchronize()
    assert (stream.cpu() - ref).abs().max().item() < 4e-3
    assert (lno.cpu().float() - F.layer_norm(ref, (C,))).abs().max().item() < 3e-2


@pytest.mark.parametrize("B,nsplit,C_,k1,h1,n_out,out_cols,act1,act2", [(16, 512, 192, 180, 6, 180, 192, 2, 4), (3, 37, 192, 192, 24, 192, 192, 1, 4),
                                                                        (2, 64, 64, 64, 0, 64, 64, 0, 0), (5, 16, 1024, 1024, 0, 1024, 1024, 0, 0)])
def test_gap_finalize_mlp(B, nsplit, C_, k1, h1, n_out, out_cols, act1, act2):
    """ff_gap_finalize_mlp (pool finalise + the squeeze-excite / channel-interaction / SCA layers in one launch, last-block ticket per
    sample) against ff_gap_finalize + ff_vec_linear x 2 and plain PyTorch; repeated launches leave the ticket counters re-armed."""
    from isr2_b200 import ops
    g = torch.Generator().manual_seed(9 + B)
    d = _dev()
    part = torch.randn(B * nsplit, C_, generator=g).to(d)
    w1 = (torch.randn(h1 if h1 else n_out, k1, generator=g) / math.sqrt(k1)).to(d)
    b1 = torch.randn(h1 if h1 else n_out, generator=g).to(d)
    w2 = (torch.randn(n_out, h1, generator=g) / math.sqrt(max(h1, 1))).to(d) if h1 else None
    b2 = torch.randn(n_out, generator=g).to(d) if h1 else None
    inv = 1.0 / (nsplit * 32)
    mean_ref = torch.zeros(B, C_, device=d)
    ops.gap_finalize(part, B, nsplit, C_, inv, mean_ref)
    if h1:
        hid = torch.zeros(B, h1, device=d)
        ops.vec_linear(mean_ref, B, k1, w1, b1, h1, act1, hid)
        ref = torch.full((B, out_cols), 5.0, device=d)
        ops.vec_linear(hid, B, h1, w2, b2, n_out, act2, ref, y_cols=out_cols)
    else:
        ref = torch.full((B, out_cols), 5.0, device=d)
        ops.vec_linear(mean_ref, B, k1, w1, b1, n_out, act1, ref, y_cols=out_cols)
    tickets = torch.zeros(1, 64, dtype=torch.int32, device=d)
    for _ in range(3):
        mean = torch.full((B, C_), 7.0, device=d)
        out = torch.full((B, out_cols), 7.0, device=d)
        ops.gap_finalize_mlp(part, B, nsplit, C_, inv, mean, tickets, w1, b1, k1, act1, out, n_out, w2=w2, b2=b2, h1=h1, act2=act2, out_cols=out_cols)
        torch.cuda.synchronize()
        assert torch.equal(mean, mean_ref)
        assert (out - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())
        assert (tickets == 0).all()
    acts = {0: lambda v: v, 1: F.gelu, 2: F.relu, 4: torch.sigmoid}
    m = part.view(B, nsplit, C_).sum(1).cpu() * inv
    y = acts[act1](m[:, :k1] @ w1.cpu().t() + b1.cpu())
    if h1:
        y = acts[act2](y @ w2.cpu().t() + b2.cpu())
    assert (out.cpu()[:, :n_out] - y).abs().max().item() < 2e-4
    assert (out.cpu()[:, n_out:] == 0).all()
