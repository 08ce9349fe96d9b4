"""NAFNet-SIDD-width64 wrapped for x4 SR on the ffb200 kernels.

Host-side mirror of `NAFNetSR.forward` (reference src/models/nafnet/__init__.py:117-139: bicubic x4 -> NAFNet ->
clamp) and `NAFNet.forward` / `NAFBlock.forward` (src/models/nafnet/nafnet_arch.py:195-225, :110-131) with
enc [2,2,4,8], middle 12, dec [2,2,2,2], width 64 (expert_loader.py:481-487).  State-dict keys are those of
the inner `nafnet.` module (the format `load_nafnet_weights` ingests).

NHWC throughout, so LayerNorm2d is a plain per-pixel layer norm; SimpleGate is folded into the depthwise
kernel (first gate) and into the conv4 epilogue (second gate); both PixelShuffle(2) are folded into stores.
"""
import ctypes as C_

import torch

from . import lib as L
from . import ops
from .hat import Workspace
from .ops import ACT_CLAMP01, ACT_NONE, CONV_1X1, CONV_2X2S2, CONV_3X3
from .packing import BF16, F32, pack_conv, pack_conv_im2col2, pack_dw, pack_matrix, pack_vector, pixel_shuffle_rows

WIDTH = 64
ENC = (2, 2, 4, 8)
DEC = (2, 2, 2, 2)
MID = 12


def _gate_perm(c):
    """Row order for conv4 so each 16-row chunk holds 8 x1 channels followed by their 8 x2 partners."""
    return torch.cat([torch.cat([torch.arange(8 * i, 8 * i + 8), c + torch.arange(8 * i, 8 * i + 8)]) for i in range(c // 8)])


class NAFNetRunner:
    def __init__(self, sd, device="cuda", enc=ENC, dec=DEC, mid=MID):
        self.device = device
        self.enc, self.dec, self.mid = enc, dec, mid
        self.ws = Workspace(device)
        dev = device
        g = lambda k: sd[k].detach().to("cpu", F32)

        def block(p, c):
            perm = _gate_perm(c)
            d = dict(
                c=c,
                n1=(g(p + "norm1.weight").to(dev), g(p + "norm1.bias").to(dev)),
                n2=(g(p + "norm2.weight").to(dev), g(p + "norm2.bias").to(dev)),
                w1=pack_matrix(g(p + "conv1.weight").reshape(2 * c, c), 2 * c, c, device=dev), b1=g(p + "conv1.bias").to(dev),
                dw=pack_dw(g(p + "conv2.weight"), 2 * c, device=dev), dwb=g(p + "conv2.bias").to(dev),
                w3=pack_matrix(g(p + "conv3.weight").reshape(c, c), c, c, device=dev), b3=g(p + "conv3.bias").to(dev),
                w3_f32=g(p + "conv3.weight").reshape(c, c).to(dev).contiguous(),
                sca_w=g(p + "sca.1.weight").reshape(c, c).to(dev).contiguous(), sca_b=g(p + "sca.1.bias").to(dev),
                w4=pack_matrix(g(p + "conv4.weight").reshape(2 * c, c)[perm], 2 * c, c, device=dev), b4=g(p + "conv4.bias")[perm].contiguous().to(dev),
                w5=pack_matrix(g(p + "conv5.weight").reshape(c, c), c, c, device=dev), b5=g(p + "conv5.bias").to(dev),
                beta=g(p + "beta").reshape(-1).to(dev).contiguous(), gamma=g(p + "gamma").reshape(-1).to(dev).contiguous(),
            )
            if c == 64:
                # operands of ff_naf_tail (the block's second half as one kernel): beta / gamma folded into conv3 / conv5, conv4 in the
                # reference's row order (the kernel pairs column j with column 64 + j)
                beta, gamma = g(p + "beta").reshape(-1), g(p + "gamma").reshape(-1)
                w3, w5 = g(p + "conv3.weight").reshape(c, c), g(p + "conv5.weight").reshape(c, c)
                d.update(w3_beta_f32=(beta[:, None] * w3).to(dev).contiguous(), w3_beta=pack_matrix(beta[:, None] * w3, c, c, device=dev),
                         b3_beta=(beta * g(p + "conv3.bias")).to(dev),
                         w4n=pack_matrix(g(p + "conv4.weight").reshape(2 * c, c), 2 * c, c, device=dev), b4n=g(p + "conv4.bias").to(dev),
                         w5_gamma=pack_matrix(gamma[:, None] * w5, c, c, device=dev), b5_gamma=(gamma * g(p + "conv5.bias")).to(dev))
            return d

        self.intro_w = pack_conv_im2col2(g("intro.weight"), WIDTH, device=dev)      # 3 -> 64 3x3 as an im2col GEMM (ops.pack_taps)
        self.intro_b = g("intro.bias").to(dev)
        self.end_w = pack_conv(g("ending.weight"), 16, WIDTH, device=dev)
        self.end_b = pack_vector(g("ending.bias"), 16, device=dev)
        c = WIDTH
        self.encoders, self.downs = [], []
        for s, n in enumerate(enc):
            self.encoders.append([block(f"encoders.{s}.{k}.", c) for k in range(n)])
            self.downs.append((pack_conv(g(f"downs.{s}.weight"), 2 * c, c, device=dev), g(f"downs.{s}.bias").to(dev)))
            c *= 2
        self.middle = [block(f"middle_blks.{k}.", c) for k in range(mid)]
        self.ups, self.decoders = [], []
        for s, n in enumerate(dec):
            rows = pixel_shuffle_rows(2 * c)
            self.ups.append(pack_matrix(g(f"ups.{s}.0.weight").reshape(2 * c, c), 2 * c, c, row_index=rows, device=dev))
            c //= 2
            self.decoders.append([block(f"decoders.{s}.{k}.", c) for k in range(n)])

    def _block(self, d, S, Sb, B, H, W, bufs, want_bf16, t_ready=False, next_norm=None):
        """One NAFBlock in place on the fp32 stream S [P, c]; Sb receives a bf16 copy when want_bf16.
        t_ready: LayerNorm2d(norm1)(S) already sits in `t` (emitted by the previous block's conv5 epilogue); next_norm: norm1 of the
        following block at this level, emitted the same way.  Rows up to 256 channels fit one n tile and take the fused path."""
        c = d["c"]
        P = B * H * W
        t, a, gt, gapv, sca, scratch = bufs
        fused = ops.fused_ln_enabled() and c <= 256
        if not t_ready:
            ops.layernorm(S, P, c, d["n1"][0], d["n1"][1], 1e-6, out_bf16=t, out_cols=c)
        ops.conv_gemm(t, B, H, W, c, d["w1"], n_store=2 * c, bias=d["b1"], out_bf16=a)
        rows = ops.dwconv_pool_rows(H, W, c, 1)
        if rows:
            # the SimpleGate depthwise kernel also emits the per-tile sums of the SCA average pool
            gpart = self.ws.get(f"gpart{c}", B * rows, c, F32)
            ops.dwconv_pool(a, B, H, W, 2 * c, d["dw"], d["dwb"], gt, gpart, mode=1)
            if ops.pool_mlp_enabled():      # pool finalise + the SCA 1x1 conv in one launch
                tickets = self.ws.get("pool_tickets", 1, max(B, 64), torch.int32)
                ops.gap_finalize_mlp(gpart, B, rows, c, 1.0 / (H * W), gapv, tickets, d["sca_w"], d["sca_b"], c, ACT_NONE, sca, c)
            else:
                ops.gap_finalize(gpart, B, rows, c, 1.0 / (H * W), gapv)
                ops.vec_linear(gapv, B, c, d["sca_w"], d["sca_b"], c, ACT_NONE, sca)
        else:
            ops.dwconv(a, B, H, W, 2 * c, 3, 3, d["dw"], d["dwb"], gt, mode=1)
            ops.gap(gt, B, H * W, c, gapv, scratch)
            ops.vec_linear(gapv, B, c, d["sca_w"], d["sca_b"], c, ACT_NONE, sca)
        if fused and c == 64 and ops.naf_tail_enabled():
            # conv3 + residual + norm2 + conv4 + SimpleGate + conv5 + residual + the next block's norm1 in one pass over the fp32 stream
            if c < H * W:
                w3b = self.ws.get(f"w3b{c}", B * c, c, BF16)
                ops.scale_weight_cols(d["w3_beta_f32"], sca, w3b.view(B, c, c))
                rows = c
            else:
                ops.scale_channels(gt, B, H * W, c, sca)
                w3b, rows = d["w3_beta"], 0
            ops.naf_tail(gt, B, H, W, w3b, d["b3_beta"], S, d["n2"], d["w4n"], d["b4n"], d["w5_gamma"], d["b5_gamma"], S, w3_batch_rows=rows,
                         out_bf16=t if next_norm is not None else (Sb if want_bf16 else None), ln=next_norm)
            return next_norm is not None
        ln2 = (d["n2"][0], d["n2"][1], 1e-6, c, t) if fused else None
        if c < H * W:
            # x * sca folded into per-sample conv3 weights (c*c per sample instead of a pass over H*W*c activations)
            w3b = self.ws.get(f"w3b{c}", B * c, c, BF16)
            ops.scale_weight_cols(d["w3_f32"], sca, w3b.view(B, c, c))
            ops.conv_gemm(gt, B, H, W, c, w3b, n_store=c, w_batch_rows=c, bias=d["b3"], col_scale=d["beta"], res=S, out_f32=S, ln=ln2)
        else:
            ops.scale_channels(gt, B, H * W, c, sca)
            ops.conv_gemm(gt, B, H, W, c, d["w3"], n_store=c, bias=d["b3"], col_scale=d["beta"], res=S, out_f32=S, ln=ln2)
        if not fused:
            ops.layernorm(S, P, c, d["n2"][0], d["n2"][1], 1e-6, out_bf16=t, out_cols=c)
        ops.conv_gemm(t, B, H, W, c, d["w4"], n_store=2 * c, bias=d["b4"], gate_pairs=1, out_bf16=gt)
        ln1 = (next_norm[0], next_norm[1], 1e-6, c, t) if (fused and next_norm is not None) else None
        ops.conv_gemm(gt, B, H, W, c, d["w5"], n_store=c, bias=d["b5"], col_scale=d["gamma"], res=S, out_f32=S,
                      out_bf16=Sb if want_bf16 else None, ln=ln1)
        return ln1 is not None

    def _run_blocks(self, blks, l, B, ready=False):
        """A run of NAFBlocks at one UNet level; each block's conv5 epilogue emits the next block's norm1 (ready: the first block's
        norm1 was already emitted by the producer of the level's stream)."""
        for k, d in enumerate(blks):
            nxt = blks[k + 1]["n1"] if k + 1 < len(blks) else None
            ready = self._block(d, l["S"], l["Sb"], B, l["H"], l["W"], l["bufs"], want_bf16=(k == len(blks) - 1), t_ready=ready, next_norm=nxt)

    def forward(self, x, out, out_off=6):
        """x: fp32 NCHW [B,3,h,w], any size.  Writes clamp(NAFNetSR(x), 0, 1) into channels out_off..out_off+2 of the fp32 expert
        stack [B*4h*4w][ld]."""
        B, _, h, w = x.shape
        nlev = len(self.enc)
        ps = 1 << nlev
        # NAFNet.check_image_size (nafnet_arch.py:219-225): the bicubic image is zero-padded on the right / bottom to a multiple
        # of 2^levels, the network runs on the padded image (its pooled statistics include the padding) and the result is
        # cropped (:216).  The padding is written by the bicubic kernel, the crop is fused into the last conv's store.
        H, W = -(-4 * h // ps) * ps, -(-4 * w // ps) * ps
        ws = self.ws
        up = ws.get("up", B * H * W, 4, F32)
        L.check(L.load().ff_bicubic_up_pad(C_.c_void_p(x.data_ptr()), B, 3, h, w, 4, C_.c_void_p(up.data_ptr()), 4, H, W, ops._stream()), "ff_bicubic_up")
        # per-level buffers
        lv = []
        c, Hc, Wc = WIDTH, H, W
        for _ in range(nlev + 1):
            P = B * Hc * Wc
            lv.append(dict(c=c, H=Hc, W=Wc, S=ws.get(f"S{c}", P, c, F32), Sb=ws.get(f"Sb{c}", P, c, BF16),
                           bufs=(ws.get(f"t{c}", P, c, BF16), ws.get(f"a{c}", P, 2 * c, BF16), ws.get(f"g{c}", P, c, BF16),
                                 ws.get(f"gap{c}", B, c, F32), ws.get(f"sca{c}", B, c, F32), ws.get("gscratch", 1, B * 64 * 1024, F32))))
            c, Hc, Wc = 2 * c, Hc // 2, Wc // 2
        l0 = lv[0]
        P0 = B * H * W
        im = l0["bufs"][1].view(-1)[:P0 * 64].view(P0, 64)        # [P, 64] bf16 scratch of level 0 (the conv1 output buffer, free until the first block)
        ops.pack_taps(up, B, H, W, 3, 3, 2, im)
        intro_fused = ops.fused_ln_enabled() and WIDTH <= 256
        if intro_fused:
            # intro conv through the fp32-residual epilogue (a never-written all-zero "residual"): one pass emits the fp32 stream AND
            # the first block's LayerNorm2d, instead of the generic fp32-store epilogue followed by a LayerNorm pass
            zero = ws.get("zero_stream", P0, WIDTH, F32)
            n1 = self.encoders[0][0]["n1"]
            ops.conv_gemm(im, B, H, W, 64, self.intro_w, n_store=WIDTH, bias=self.intro_b, res=zero, out_f32=l0["S"],
                          ln=(n1[0], n1[1], 1e-6, WIDTH, l0["bufs"][0]))
        else:
            ops.conv_gemm(im, B, H, W, 64, self.intro_w, n_store=WIDTH, bias=self.intro_b, out_f32=l0["S"])
        first_ready = {0: intro_fused}      # level -> the first block's norm1 already sits in the level's t buffer
        for s in range(nlev):
            l = lv[s]
            blks = self.encoders[s]
            self._run_blocks(blks, l, B, ready=first_ready.get(s, False))
            dw_, db_ = self.downs[s]
            nl = lv[s + 1]
            nblks = self.encoders[s + 1] if s + 1 < nlev else self.middle
            down_fused = ops.fused_ln_enabled() and nl["c"] <= 256
            if down_fused:      # as the intro conv: the level's fp32 stream and its first LayerNorm2d in one pass
                zero = ws.get(f"zero_stream{nl['c']}", B * nl["H"] * nl["W"], nl["c"], F32)
                n1 = nblks[0]["n1"]
                ops.conv_gemm(l["Sb"], B, l["H"], l["W"], l["c"], dw_, kind=CONV_2X2S2, n_store=2 * l["c"], bias=db_, res=zero, out_f32=nl["S"],
                              ln=(n1[0], n1[1], 1e-6, nl["c"], nl["bufs"][0]))
            else:
                ops.conv_gemm(l["Sb"], B, l["H"], l["W"], l["c"], dw_, kind=CONV_2X2S2, n_store=2 * l["c"], bias=db_, out_f32=nl["S"])
            first_ready[s + 1] = down_fused
        l = lv[nlev]
        self._run_blocks(self.middle, l, B, ready=first_ready.get(nlev, False))
        for s in range(len(self.dec)):
            src, dst = lv[nlev - s], lv[nlev - s - 1]
            # 1x1 conv c -> 2c (no bias) + PixelShuffle(2) + encoder skip, written in place over the skip buffer
            ops.conv_gemm(src["Sb"], B, src["H"], src["W"], src["c"], self.ups[s], n_store=2 * src["c"], pixel_shuffle=2, res=dst["S"], out_f32=dst["S"])
            self._run_blocks(self.decoders[s], dst, B)
        ops.conv_gemm(l0["Sb"], B, H, W, WIDTH, self.end_w, kind=CONV_3X3, n_store=3, bias=self.end_b, res=up, post_act=ACT_CLAMP01,
                      out_f32=out[:, out_off:], out_crop=(4 * h, 4 * w) if (H, W) != (4 * h, 4 * w) else None)
        return out
