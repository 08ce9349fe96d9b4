"""DAT x4 expert on the ffb200 kernels.

Host-side mirror of `DAT.forward` (reference src/models/dat/dat_arch.py:1007-1028) as configured by
expert_loader.py:412-420: embed 180, 6 residual groups x 6 DATB, 6 heads, split_size [8, 32], expansion 4.
Even blocks: adaptive spatial attention (two 3-head branches with 8x32 / 32x8 windows, shifted in the
blocks selected by `_should_shift`), odd blocks: adaptive channel attention; every block ends in SGFN.
Consumes the reference state_dict unchanged.  Any image size the reference accepts (reflect padding to 16, window padding to 32).
"""
import ctypes as C_

import torch

from . import lib as L
from . import ops
from .hat import CP, RGB_MEAN, Workspace, _qkv_rows, pack_qkv_bias
from .ops import ACT_CLAMP01, ACT_GELU, ACT_LRELU, ACT_NONE, ACT_SIGMOID, CONV_3X3
from .packing import (BF16, F32, fold_bn, head_pad_index, pack_conv, pack_conv_direct, pack_dw, pack_matrix,
                      pack_vector, pixel_shuffle_rows)

C = 180
HEADS = 6
SPLIT = (8, 32)
HID = 720
HP = 384       # each SGFN half (360) padded to 384


def should_shift(rg, b):   # dat_arch.py:426-429
    return (rg % 2 == 0 and b > 0 and (b - 2) % 4 == 0) or (rg % 2 != 0 and b % 4 == 0)


def _dyn_pos_table(g, p, hs, ws):
    """DynamicPosBias (dat_arch.py:177-212) on the fixed offset grid: input independent, so it is
    evaluated once at load time on the host (fp32) -> [(2hs-1)(2ws-1), 3]."""
    import torch.nn.functional as F
    by, bx = torch.meshgrid(torch.arange(1 - hs, hs), torch.arange(1 - ws, ws), indexing="ij")
    x = torch.stack([by.reshape(-1), bx.reshape(-1)], 1).float()
    x = F.linear(x, g(p + "pos_proj.weight"), g(p + "pos_proj.bias"))
    for n in ("pos1", "pos2", "pos3"):
        x = F.layer_norm(x, (x.shape[-1],), g(p + n + ".0.weight"), g(p + n + ".0.bias"), 1e-5)
        x = F.linear(F.relu(x), g(p + n + ".2.weight"), g(p + n + ".2.bias"))
    return x.contiguous()


class DATRunner:
    def __init__(self, sd, device="cuda", groups=6, blocks=6):
        self.device = device
        self.groups, self.nblocks = groups, blocks
        self.ws = Workspace(device)
        dev = device
        g = lambda k: sd[k].detach().to("cpu", F32)
        hp = head_pad_index(torch.arange(C))
        scale = (C // HEADS) ** -0.5 * 1.4426950408889634   # q * head_dim^-0.5, and log2(e) for the exp2 softmax
        self.mean = torch.tensor(RGB_MEAN, dtype=F32, device=dev)
        self.conv_first_w = pack_conv_direct(g("conv_first.weight"), CP, dev)
        self.conv_first_b = pack_vector(g("conv_first.bias"), CP, device=dev)
        ln = lambda p: (pack_vector(g(p + "weight"), CP, device=dev), pack_vector(g(p + "bias"), CP, device=dev))   # zero-padded to the 192-wide row
        self.before = ln("before_RG.1.")
        half = torch.arange(HID // 2)
        fc1_rows = torch.cat([half, HP + half])

        self.layers = []
        for rg in range(groups):
            blks = []
            for bi in range(blocks):
                p = f"layers.{rg}.blocks.{bi}."
                a = p + "attn."
                d = dict(norm1=ln(p + "norm1."), norm2=ln(p + "norm2."), spatial=(bi % 2 == 0), shift=should_shift(rg, bi))
                wq, bq = g(a + "qkv.weight").clone(), g(a + "qkv.bias").clone()
                if d["spatial"]:
                    wq[:C] *= scale
                    bq[:C] *= scale
                    d["tables"] = [_dyn_pos_table(g, a + f"attns.{br}.pos.", *((SPLIT[0], SPLIT[1]) if br == 0 else (SPLIT[1], SPLIT[0]))).t().contiguous().to(dev)
                                   for br in range(2)]   # [heads][T]
                else:
                    d["temperature"] = g(a + "temperature").reshape(-1).to(dev).contiguous()
                d["qkv_w"] = pack_matrix(wq, 3 * CP, CP, row_index=_qkv_rows(), device=dev)
                d["qkv_b"] = pack_qkv_bias(bq, dev)
                d["proj_w"] = pack_matrix(g(a + "proj.weight"), CP, CP, col_index=hp, device=dev)
                d["proj_b"] = pack_vector(g(a + "proj.bias"), CP, device=dev)
                # depthwise conv branch on v: conv -> BN(eval) -> GELU, BN folded; head-padded channels
                w, b = fold_bn(g(a + "dwconv.0.weight"), g(a + "dwconv.0.bias"), g(a + "dwconv.1.weight"), g(a + "dwconv.1.bias"),
                               g(a + "dwconv.1.running_mean"), g(a + "dwconv.1.running_var"))
                d["dw_w"] = pack_dw(w, CP, index=hp, device=dev)
                d["dw_b"] = pack_vector(b, CP, index=hp, device=dev)
                # channel interaction: GAP -> 1x1 (C -> C/8) -> BN -> GELU -> 1x1 (C/8 -> C)
                w, b = fold_bn(g(a + "channel_interaction.1.weight"), g(a + "channel_interaction.1.bias"), g(a + "channel_interaction.2.weight"),
                               g(a + "channel_interaction.2.bias"), g(a + "channel_interaction.2.running_mean"), g(a + "channel_interaction.2.running_var"))
                d["ci1_w"] = pack_matrix(w.reshape(w.shape[0], C), 24, CP, col_index=hp, dtype=F32, device=dev)
                d["ci1_b"] = pack_vector(b, 24, device=dev)
                d["ci2_w"] = pack_matrix(g(a + "channel_interaction.4.weight").reshape(C, -1), CP, 24, row_index=hp, dtype=F32, device=dev)
                d["ci2_b"] = pack_vector(g(a + "channel_interaction.4.bias"), CP, index=hp, device=dev)
                # spatial interaction: 1x1 (C -> C/16) -> BN -> GELU -> 1x1 (C/16 -> 1)
                w, b = fold_bn(g(a + "spatial_interaction.0.weight"), g(a + "spatial_interaction.0.bias"), g(a + "spatial_interaction.1.weight"),
                               g(a + "spatial_interaction.1.bias"), g(a + "spatial_interaction.1.running_mean"), g(a + "spatial_interaction.1.running_var"))
                d["si_hid"] = w.shape[0]
                d["si1_w"] = pack_matrix(w.reshape(w.shape[0], C), 32, CP, col_index=hp, device=dev)     # tensor-core layer, rows >= hid zero
                d["si1_b"] = pack_vector(b, 32, device=dev)
                d["si2_w"] = g(a + "spatial_interaction.3.weight").reshape(-1).to(dev).contiguous()
                d["si2_b"] = float(g(a + "spatial_interaction.3.bias").item())
                # SGFN: fc1 (180 -> 720) split into two 360-wide halves, each padded to 384
                f = p + "ffn."
                d["fc1_w"] = pack_matrix(g(f + "fc1.weight"), 2 * HP, CP, row_index=fc1_rows, device=dev)
                d["fc1_b"] = pack_vector(g(f + "fc1.bias"), 2 * HP, index=fc1_rows, device=dev)
                d["sg_norm"] = (g(f + "sg.norm.weight").to(dev), g(f + "sg.norm.bias").to(dev))
                d["sg_w"] = pack_dw(g(f + "sg.conv.weight"), HP, device=dev)
                d["sg_b"] = pack_vector(g(f + "sg.conv.bias"), HP, device=dev)
                d["fc2_w"] = pack_matrix(g(f + "fc2.weight"), CP, HP, device=dev)
                d["fc2_b"] = pack_vector(g(f + "fc2.bias"), CP, device=dev)
                blks.append(d)
            self.layers.append(dict(blocks=blks, conv_w=pack_conv(g(f"layers.{rg}.conv.weight"), CP, CP, device=dev),
                                    conv_b=pack_vector(g(f"layers.{rg}.conv.bias"), CP, device=dev)))
        self.norm = ln("norm.")
        self.cab_w = pack_conv(g("conv_after_body.weight"), CP, CP, device=dev)
        self.cab_b = pack_vector(g("conv_after_body.bias"), CP, device=dev)
        self.cbu_w = pack_conv(g("conv_before_upsample.0.weight"), 64, CP, device=dev)
        self.cbu_b = pack_vector(g("conv_before_upsample.0.bias"), 64, device=dev)
        ps = pixel_shuffle_rows(256)
        self.up0_w = pack_conv(g("upsample.0.weight"), 256, 64, row_index=ps, device=dev)
        self.up0_b = pack_vector(g("upsample.0.bias"), 256, index=ps, device=dev)
        self.up2_w = pack_conv(g("upsample.2.weight"), 256, 64, row_index=ps, device=dev)
        self.up2_b = pack_vector(g("upsample.2.bias"), 256, index=ps, device=dev)
        self.last_w = pack_conv(g("conv_last.weight"), 16, 64, device=dev)
        self.last_b = pack_vector(g("conv_last.bias") + torch.tensor(RGB_MEAN), 16, device=dev)

    def forward(self, x, out, out_off=3):
        """x: fp32 NCHW [B,3,h,w].  Writes clamp(DAT(x),0,1) into channels
        out_off..out_off+2 of the fp32 expert stack `out` ([B*4H*4W][ld])  (= forward_dat, expert_loader.py:623-652)."""
        B, _, h0, w0 = x.shape
        # forward_dat (expert_loader.py:623-652): reflect-pad to a multiple of 16, run, crop.  Inside, the spatial attention
        # zero-pads the projected q / k / v to multiples of 32 (dat_arch.py:505-512) -- done by the attention kernel's padded
        # geometry, no copy.
        H, W = -(-h0 // 16) * 16, -(-w0 // 16) * 16
        if H - h0 >= h0 or W - w0 >= w0:
            raise ValueError(f"DATRunner: image {h0}x{w0} is smaller than its reflect padding (the reference's F.pad fails here too)")
        Hp, Wp = -(-H // 32) * 32, -(-W // 32) * 32
        padded = (Hp, Wp) if (Hp, Wp) != (H, W) else None
        M, N = B * H * W, H * W
        ws = self.ws
        lib = L.load()
        st = ops._stream
        img = ws.get("img", M, 4, F32)
        x0 = ws.get("x0", M, CP, F32)
        G = ws.get("G", M, CP, F32)
        X = ws.get("X", M, CP, F32)
        t = ws.get("t", M, CP, BF16)
        qkv = ws.get("qkv", M, 3 * CP, BF16)
        att = ws.get("att", M, CP, BF16)
        convx = ws.get("convx", M, CP, BF16)
        mix = ws.get("mix", M, CP, BF16)
        h = ws.get("h", M, 2 * HP, BF16)
        t2 = ws.get("t2", M, HP, BF16)
        gt = ws.get("gt", M, HP, BF16)
        Xb = ws.get("Xb", M, CP, BF16)
        gapv = ws.get("gap", B, CP, F32)
        ci_h = ws.get("ci_h", B, 24, F32)
        cmap = ws.get("cmap", B, CP, F32)
        pool_mlp = ops.pool_mlp_enabled()
        tickets = ws.get("pool_tickets", 1, max(B, 64), torch.int32)      # per-sample arrival counters of ff_gap_finalize_mlp (self-resetting)
        sih = ws.get("si_hid", M, 32, BF16)
        scratch = ws.get("scratch", 1, max(B * 64 * CP, B * HEADS * ((N + 511) // 512) * 1088), F32)
        wb = ws.get("chan_w", B * CP, CP, BF16)   # block-diagonal channel-attention weights (off-diagonal stays zero)

        if (H, W) == (h0, w0):
            ops.nchw_to_nhwc(x, img, sub=self.mean)
        else:
            ops.nchw_to_nhwc_pad(x, img, H, W, sub=self.mean, reflect=True)
        ops.conv_direct(img, B, H, W, 3, 3, self.conv_first_w, self.conv_first_b, n_store=CP, out_f32=x0)
        ops.layernorm(x0, M, C, self.before[0], self.before[1], 1e-5, out_f32=G, out_cols=CP)

        fused = ops.fused_ln_enabled()      # every LayerNorm after a residual add leaves the producing GEMM's epilogue
        t_ready = False
        for li, layer in enumerate(self.layers):
            src = G
            nb = len(layer["blocks"])
            for bi, d in enumerate(layer["blocks"]):
                if not t_ready:
                    ops.layernorm(src, M, C, d["norm1"][0], d["norm1"][1], 1e-5, out_bf16=t, out_cols=CP)
                ops.conv_gemm(t, B, H, W, CP, d["qkv_w"], n_store=3 * CP, bias=d["qkv_b"], out_bf16=qkv)
                # conv branch on v (image form of v: channels 384..575 of qkv)
                pool_rows = ops.dwconv_pool_rows(H, W, CP, 0) if d["spatial"] else 0
                if pool_rows:
                    # spatial blocks pool the conv branch: the depthwise kernel emits the per-tile sums
                    gpart = ws.get("gpart_dw", B * pool_rows, CP, F32)
                    ops.dwconv_pool(qkv, B, H, W, CP, d["dw_w"], d["dw_b"], convx, gpart, act=ACT_GELU, x_off=2 * CP)
                else:
                    ops.dwconv(qkv, B, H, W, CP, 3, 3, d["dw_w"], d["dw_b"], convx, act=ACT_GELU, x_off=2 * CP)
                if d["spatial"]:
                    for br in range(2):
                        wh, ww = (SPLIT[0], SPLIT[1]) if br == 0 else (SPLIT[1], SPLIT[0])
                        sh = (wh // 2, ww // 2) if d["shift"] else (0, 0)
                        ops.window_attention(qkv, B, H, W, att, bias_table=d["tables"][br], wh=wh, ww=ww, shift=sh, heads=3, head_off=3 * br, padded=padded)
                    gap_src, mode = convx, 0
                else:
                    L.check(lib.ff_dat_channel_attention_weights(C_.c_void_p(qkv.data_ptr()), 3 * CP, 0, CP, B, N, HEADS, C // HEADS,
                                                                 C_.c_void_p(d["temperature"].data_ptr()), C_.c_void_p(wb.data_ptr()),
                                                                 C_.c_void_p(scratch.data_ptr()), C_.c_size_t(scratch.numel() * 4), st()),
                            "ff_dat_channel_attention_weights")
                    # attn @ v with per-sample block-diagonal weights; A = v (channels 384.. of qkv)
                    # the store epilogue emits the pool partials of the attention branch
                    pool_rows = (H // 8) * (W // 16) * 4          # one row of partial sums per (8x16-pixel tile, 32-row quadrant)
                    gpart = ws.get("gpart_att", B * pool_rows, CP, F32)
                    ops.conv_gemm(qkv[:, 2 * CP:], B, H, W, CP, wb, n_store=CP, w_batch_rows=CP, out_bf16=att, x_ld=3 * CP, col_sums=gpart)
                    gap_src, mode = att, 1
                if pool_rows and pool_mlp:      # pool finalise + both channel-interaction layers in one launch
                    ops.gap_finalize_mlp(gpart, B, pool_rows, CP, 1.0 / N, gapv, tickets, d["ci1_w"], d["ci1_b"], CP, ACT_GELU, cmap, CP,
                                         w2=d["ci2_w"], b2=d["ci2_b"], h1=24, act2=ACT_SIGMOID, out_cols=CP)
                else:
                    if pool_rows:
                        ops.gap_finalize(gpart, B, pool_rows, CP, 1.0 / N, gapv)
                    else:
                        ops.gap(gap_src, B, N, CP, gapv, scratch)
                    ops.vec_linear(gapv, B, CP, d["ci1_w"], d["ci1_b"], 24, ACT_GELU, ci_h, y_cols=24)
                    ops.vec_linear(ci_h, B, 24, d["ci2_w"], d["ci2_b"], CP, ACT_SIGMOID, cmap, y_cols=CP)
                # spatial interaction: first layer (C -> C/16, BN folded, GELU) on the tensor cores, second layer inside the gate kernel
                ops.conv_gemm(convx if mode else att, B, H, W, CP, d["si1_w"], n_store=32, bias=d["si1_b"], act=ACT_GELU, out_bf16=sih)
                L.check(lib.ff_dat_aim(C_.c_void_p(att.data_ptr()), CP, C_.c_void_p(convx.data_ptr()), CP, C_.c_void_p(cmap.data_ptr()), CP,
                                       C_.c_void_p(sih.data_ptr()), 32, C_.c_void_p(d["si2_w"].data_ptr()),
                                       C_.c_float(d["si2_b"]), d["si_hid"], mode, C_.c_longlong(M), N, C_.c_void_p(mix.data_ptr()), CP, st()),
                        "ff_dat_aim")
                ops.conv_gemm(mix, B, H, W, CP, d["proj_w"], n_store=CP, bias=d["proj_b"], res=src, out_f32=X,
                              ln=(d["norm2"][0], d["norm2"][1], 1e-5, C, t) if fused else None)
                # SGFN
                if not fused:
                    ops.layernorm(X, M, C, d["norm2"][0], d["norm2"][1], 1e-5, out_bf16=t, out_cols=CP)
                ops.conv_gemm(t, B, H, W, CP, d["fc1_w"], n_store=2 * HP, bias=d["fc1_b"], act=ACT_GELU, out_bf16=h)
                ops.layernorm(h, M, HID // 2, d["sg_norm"][0], d["sg_norm"][1], 1e-5, out_bf16=t2, out_cols=HP, x_off=HP)
                ops.dwconv(t2, B, H, W, HP, 3, 3, d["sg_w"], d["sg_b"], gt, mul=h)
                last = (bi == nb - 1)
                nxt = None if last else layer["blocks"][bi + 1]["norm1"]
                ops.conv_gemm(gt, B, H, W, HP, d["fc2_w"], n_store=CP, bias=d["fc2_b"], res=X, out_f32=X, out_bf16=Xb if last else None,
                              ln=(nxt[0], nxt[1], 1e-5, C, t) if (fused and nxt is not None) else None)
                t_ready = fused and nxt is not None
                src = X
            # group tail conv + residual; its epilogue emits the LayerNorm of the next consumer of G
            nxt = self.layers[li + 1]["blocks"][0]["norm1"] if li + 1 < len(self.layers) else self.norm
            ops.conv_gemm(Xb, B, H, W, CP, layer["conv_w"], kind=CONV_3X3, n_store=CP, bias=layer["conv_b"], res=G, out_f32=G,
                          ln=(nxt[0], nxt[1], 1e-5, C, t) if fused else None)
            t_ready = fused

        if not t_ready:
            ops.layernorm(G, M, C, self.norm[0], self.norm[1], 1e-5, out_bf16=t, out_cols=CP)
        y = Xb
        ops.conv_gemm(t, B, H, W, CP, self.cab_w, kind=CONV_3X3, n_store=CP, bias=self.cab_b, res=x0, out_bf16=y)
        f64 = ws.get("f64", M, 64, BF16)
        ops.conv_gemm(y, B, H, W, CP, self.cbu_w, kind=CONV_3X3, n_store=64, bias=self.cbu_b, act=ACT_LRELU, out_bf16=f64)
        u1 = ws.get("u1", M * 4, 64, BF16)
        ops.conv_gemm(f64, B, H, W, 64, self.up0_w, kind=CONV_3X3, n_store=256, bias=self.up0_b, pixel_shuffle=2, out_bf16=u1)
        u2 = ws.get("u2", M * 16, 64, BF16)
        ops.conv_gemm(u1, B, 2 * H, 2 * W, 64, self.up2_w, kind=CONV_3X3, n_store=256, bias=self.up2_b, pixel_shuffle=2, out_bf16=u2)
        ops.conv_gemm(u2, B, 4 * H, 4 * W, 64, self.last_w, kind=CONV_3X3, n_store=3, bias=self.last_b, post_act=ACT_CLAMP01,
                      out_f32=out[:, out_off:], out_crop=(4 * h0, 4 * w0) if (H, W) != (h0, w0) else None)
        return out
