"""HAT-L x4 expert on the ffb200 kernels.

Host-side mirror of `HAT.forward` (reference src/models/hat/hat_arch.py:971-984) for the HAT-L
configuration built by `create_hat_model` (hat/__init__.py:63-118): embed 180, 12 RHAG x (6 HAB + OCAB),
6 heads, window 16, shift 8, mlp_ratio 2, compress 3, squeeze 30, conv_scale 0.01, overlap 0.5.
Consumes the reference state_dict unchanged (keys of Appendix A of SURVEY.md).

Data layout: tokens are NHWC rows [B*H*W][192]; the fp32 residual stream stays fp32, every GEMM
operand is bf16.  q/k/v and the attention output use the head-padded channel layout (30 -> 32).
"""
import torch

from . import ops
from .ops import ACT_CLAMP01, ACT_GELU, ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, CONV_1X1, CONV_3X3
from .packing import (BF16, F32, head_pad_index, pack_conv, pack_conv_direct, pack_matrix, pack_vector,
                      pixel_shuffle_rows)

C = 180
CP = 192
HEADS = 6
WS = 16
RGB_MEAN = (0.4488, 0.4371, 0.4040)   # hat_arch.py:775 (plain attribute, not in the state dict)


class Workspace:
    """Named device buffers cached per shape (keeps addresses stable for CUDA-graph capture).  Every entry remembers the
    forward (`epoch`) that last used it, so the owner can evict the buffers of shapes that have not been seen for a while
    (FreqFusionB200._trim_workspaces) -- the cache is bounded, not monotonically growing."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}
        self.epoch = 0

    def get(self, name, rows, cols, dtype, zero=False):
        key = (name, rows, cols, dtype)
        ent = self.bufs.get(key)
        if ent is None:
            ent = self.bufs[key] = [torch.zeros(rows, cols, dtype=dtype, device=self.device), self.epoch]
        else:
            ent[1] = self.epoch
            if zero:
                ent[0].zero_()
        return ent[0]

    def nbytes(self):
        return sum(e[0].numel() * e[0].element_size() for e in self.bufs.values())

    def evict_unused_since(self, epoch):
        """Drops every buffer last used before `epoch`; returns the bytes released."""
        dead = [k for k, e in self.bufs.items() if e[1] < epoch]
        freed = sum(self.bufs[k][0].numel() * self.bufs[k][0].element_size() for k in dead)
        for k in dead:
            del self.bufs[k]
        return freed


def _qkv_rows():
    o = torch.arange(3 * C)
    return (o // C) * CP + head_pad_index(o % C)


def pack_qkv_bias(bq, device):
    """qkv bias in the head-padded layout; padding dim 31 of every v head is set to 1.0 (its weight row is zero) so that
    ff_window_attention gets the softmax row sums out of the P.V MMA (see csrc/window_attention.cu)."""
    b = torch.zeros(3 * CP, dtype=F32)
    b[_qkv_rows()] = bq.to(F32)
    b[2 * CP + 31::32] = 1.0
    return b.to(device).contiguous()


class HATRunner:
    def __init__(self, sd, device="cuda", depths=12, blocks=6):
        self.device = device
        self.depths, self.blocks = depths, blocks
        self.ws = Workspace(device)
        g = lambda k: sd[k].detach().to("cpu", F32)
        dev = device
        hp = head_pad_index(torch.arange(C))
        scale = (C // HEADS) ** -0.5 * 1.4426950408889634   # q * head_dim^-0.5, and log2(e) for the exp2 softmax
        self.mean = torch.tensor(RGB_MEAN, dtype=F32, device=dev)
        self.conv_first_w = pack_conv_direct(g("conv_first.weight"), CP, dev)
        self.conv_first_b = pack_vector(g("conv_first.bias"), CP, device=dev)
        self.pe_norm = (g("patch_embed.norm.weight").to(dev), g("patch_embed.norm.bias").to(dev))

        def attn_pack(prefix):
            wq = g(prefix + "qkv.weight").clone()
            bq = g(prefix + "qkv.bias").clone()
            wq[:C] *= scale   # q = q * scale (hat_arch.py:175) folded into the projection
            bq[:C] *= scale
            return dict(
                qkv_w=pack_matrix(wq, 3 * CP, CP, row_index=_qkv_rows(), device=dev),
                qkv_b=pack_qkv_bias(bq, dev),
                proj_w=pack_matrix(g(prefix + "proj.weight"), CP, CP, col_index=hp, device=dev),
                proj_b=pack_vector(g(prefix + "proj.bias"), CP, device=dev),
                table=g(prefix + "relative_position_bias_table").t().contiguous().to(dev),   # [heads][T]
            )

        def mlp_pack(prefix):
            return dict(
                fc1_w=pack_matrix(g(prefix + "fc1.weight"), 2 * CP, CP, device=dev),
                fc1_b=pack_vector(g(prefix + "fc1.bias"), 2 * CP, device=dev),
                fc2_w=pack_matrix(g(prefix + "fc2.weight"), CP, 2 * CP, device=dev),
                fc2_b=pack_vector(g(prefix + "fc2.bias"), CP, device=dev),
            )

        def ln(prefix):
            # [gamma, beta] padded to the 192-wide row with zeros (the fused epilogue normalises whole rows; padding stays 0)
            return (pack_vector(g(prefix + "weight"), CP, device=dev), pack_vector(g(prefix + "bias"), CP, device=dev))

        self.layers = []
        for i in range(depths):
            pre = f"layers.{i}.residual_group."
            habs = []
            for j in range(blocks):
                bp = pre + f"blocks.{j}."
                d = dict(norm1=ln(bp + "norm1."), norm2=ln(bp + "norm2."))
                d.update(attn_pack(bp + "attn."))
                d.update(mlp_pack(bp + "mlp."))
                d["cab1_w"] = pack_conv(g(bp + "conv_block.cab.0.weight"), 64, CP, device=dev)
                d["cab1_b"] = pack_vector(g(bp + "conv_block.cab.0.bias"), 64, device=dev)
                d["cab2_w"] = pack_conv(g(bp + "conv_block.cab.2.weight"), CP, 64, device=dev)
                d["cab2_b"] = pack_vector(g(bp + "conv_block.cab.2.bias"), CP, device=dev)
                d["se1_w"] = g(bp + "conv_block.cab.3.attention.1.weight").reshape(6, C).to(dev).contiguous()
                d["se1_b"] = g(bp + "conv_block.cab.3.attention.1.bias").to(dev)
                d["se2_w"] = g(bp + "conv_block.cab.3.attention.3.weight").reshape(C, 6).to(dev).contiguous()
                d["se2_b"] = g(bp + "conv_block.cab.3.attention.3.bias").to(dev)
                habs.append(d)
            op = pre + "overlap_attn."
            oc = dict(norm1=ln(op + "norm1."), norm2=ln(op + "norm2."))
            oc.update(attn_pack(op))
            oc.update(mlp_pack(op + "mlp."))
            conv_w = pack_conv(g(f"layers.{i}.conv.weight"), CP, CP, device=dev)
            conv_b = pack_vector(g(f"layers.{i}.conv.bias"), CP, device=dev)
            self.layers.append(dict(habs=habs, ocab=oc, conv_w=conv_w, conv_b=conv_b))
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
        # x / img_range + mean (hat_arch.py:982) folded into the bias (img_range == 1)
        self.last_b = pack_vector(g("conv_last.bias") + torch.tensor(RGB_MEAN), 16, device=dev)

    # ------------------------------------------------------------------------------------------
    def _mlp(self, d, X, t, h, B, H, W, M, extra_bf16=None, t_ready=False, next_norm=None):
        """x += fc2(GELU(fc1(LN2 x))) (hat_arch.py:308).  t_ready: LN2(x) was already emitted into `t` by the producer of X;
        next_norm: (gamma, beta) of the LayerNorm that consumes the new X -- emitted into `t` by the fc2 epilogue."""
        if not t_ready:
            ops.layernorm(X, M, C, d["norm2"][0], d["norm2"][1], 1e-5, out_bf16=t, out_cols=CP)
        ln = (next_norm[0], next_norm[1], 1e-5, C, t) if next_norm is not None else None
        if ops.mlp_fused_enabled():
            # one kernel, hidden tile on chip.  The next LayerNorm may overwrite t in place: a CTA stores the rows of a tile only
            # after that tile's A operand (the same rows of t) has been consumed, and no other CTA reads them
            ops.mlp_fused(t, B, H, W, d["fc1_w"], d["fc1_b"], d["fc2_w"], d["fc2_b"], X, out_bf16=extra_bf16, ln=ln)
        else:
            ops.conv_gemm(t, B, H, W, CP, d["fc1_w"], n_store=2 * CP, bias=d["fc1_b"], act=ACT_GELU, out_bf16=h)
            ops.conv_gemm(h, B, H, W, 2 * CP, d["fc2_w"], n_store=CP, bias=d["fc2_b"], res=X, out_f32=X, out_bf16=extra_bf16, ln=ln)

    def forward(self, x, out, out_off=0):
        """x: fp32 NCHW [B,3,h,w] (any size the reference's reflect padding accepts) on the GPU.
        out: fp32 [B*4h*4w][ld] expert stack; channels out_off..out_off+2 receive clamp(SR, 0, 1)
        (= ExpertEnsemble.forward_hat, expert_loader.py:592-621)."""
        B, _, h0, w0 = x.shape
        # ExpertEnsemble.forward_hat (expert_loader.py:592-621): reflect-pad right / bottom to a multiple of the window size,
        # run, crop to 4h x 4w.  The pad is fused into the NCHW -> NHWC conversion, the crop into the last conv's store.
        H, W = -(-h0 // WS) * WS, -(-w0 // WS) * WS
        if H - h0 >= h0 or W - w0 >= w0:
            raise ValueError(f"HATRunner: image {h0}x{w0} is smaller than its reflect padding (the reference's F.pad fails here too)")
        M = B * H * W
        ws = self.ws
        img = ws.get("img", M, 4, F32)
        x0 = ws.get("x0", M, CP, F32)
        G = ws.get("G", M, CP, F32)
        X = ws.get("X", M, CP, F32)
        t = ws.get("t", M, CP, BF16)
        qkv = ws.get("qkv", M, 3 * CP, BF16)
        att = ws.get("att", M, CP, BF16)
        cab1 = ws.get("cab1", M, 64, BF16)
        cab2 = ws.get("cab2", M, CP, BF16)
        h = ws.get("h", M, 2 * CP, BF16)
        Xb = ws.get("Xb", M, CP, BF16)
        gapv = ws.get("gap", B, CP, F32)
        gpart = ws.get("gap_part", B * (H * W // 32), CP, F32)
        se_h = ws.get("se_h", B, 8, F32)
        se = ws.get("se", B, CP, F32)
        scratch = ws.get("gap_scratch", 1, B * 64 * CP, F32)
        wcat = ws.get("proj_cat_w", B * CP, 2 * CP, BF16)      # per-sample [W_proj | diag(0.01 se)] of the current block
        fused = ops.fused_ln_enabled()      # every LayerNorm after a residual add leaves the producing GEMM's epilogue
        tail = fused and ops.hab_tail_enabled()      # proj + shortcut + LN2 + MLP + residual + next LN as one kernel
        pool_mlp = ops.pool_mlp_enabled()
        tickets = ws.get("pool_tickets", 1, max(B, 64), torch.int32)      # per-sample arrival counters of ff_gap_finalize_mlp (self-resetting)

        if (H, W) == (h0, w0):
            ops.nchw_to_nhwc(x, img, sub=self.mean)
        else:
            ops.nchw_to_nhwc_pad(x, img, H, W, sub=self.mean, reflect=True)
        ops.conv_direct(img, B, H, W, 3, 3, self.conv_first_w, self.conv_first_b, n_store=CP, out_f32=x0)
        ops.layernorm(x0, M, C, self.pe_norm[0], self.pe_norm[1], 1e-5, out_f32=G, out_cols=CP)

        t_ready = False      # `t` already holds the LayerNorm the next consumer needs
        for li, layer in enumerate(self.layers):
            src = G
            nhab = len(layer["habs"])
            for j, d in enumerate(layer["habs"]):
                shift = WS // 2 if (j % 2 == 1) else 0
                if not t_ready:
                    ops.layernorm(src, M, C, d["norm1"][0], d["norm1"][1], 1e-5, out_bf16=t, out_cols=CP)
                # CAB on the LN1 output (hat_arch.py:272-277)
                ops.conv_gemm(t, B, H, W, CP, d["cab1_w"], kind=CONV_3X3, n_store=64, bias=d["cab1_b"], act=ACT_GELU, out_bf16=cab1)
                # the conv's store epilogue also emits the per-tile column sums of the squeeze-excite average pool
                ops.conv_gemm(cab1, B, H, W, 64, d["cab2_w"], kind=CONV_3X3, n_store=CP, bias=d["cab2_b"], out_bf16=cab2, col_sums=gpart)
                if pool_mlp:      # pool finalise + the two squeeze-excite layers in one launch
                    ops.gap_finalize_mlp(gpart, B, H * W // 32, CP, 1.0 / (H * W), gapv, tickets, d["se1_w"], d["se1_b"], C, ACT_RELU, se, C,
                                         w2=d["se2_w"], b2=d["se2_b"], h1=6, act2=ACT_SIGMOID, out_cols=CP)
                else:
                    ops.gap_finalize(gpart, B, H * W // 32, CP, 1.0 / (H * W), gapv)
                    ops.vec_linear(gapv, B, C, d["se1_w"], d["se1_b"], 6, ACT_RELU, se_h, y_cols=8)
                    ops.vec_linear(se_h, B, 6, d["se2_w"], d["se2_b"], C, ACT_SIGMOID, se, y_cols=CP)
                # (S)W-MSA
                ops.conv_gemm(t, B, H, W, CP, d["qkv_w"], n_store=3 * CP, bias=d["qkv_b"], out_bf16=qkv)
                ops.window_attention(qkv, B, H, W, att, bias_table=d["table"], wh=WS, ww=WS, shift=(shift, shift))
                # x = shortcut + attn + 0.01 * cab   (hat_arch.py:306); the epilogue also emits LN2(x) into t
                nxt = layer["habs"][j + 1]["norm1"] if j + 1 < nhab else layer["ocab"]["norm1"]
                if tail:
                    # everything after the attention as one kernel: x1 stays in TMEM, LN2(x1) in shared memory (csrc/hab_tail.cu)
                    # (the 0.01 * cab * se term is a diagonal K block that the kernel generates from `se`: no per-sample weights)
                    ops.hab_tail(att, B, H, W, d["proj_w"], d["proj_b"], src, d["norm2"], d["fc1_w"], d["fc1_b"], d["fc2_w"], d["fc2_b"], X,
                                 a1=cab2, a1_diag=se, a1_alpha=0.01, ln=(nxt[0], nxt[1], t))
                    t_ready = True
                    src = X
                    continue
                ln2 = (d["norm2"][0], d["norm2"][1], 1e-5, C, t) if fused else None
                if ops.concat_aux_enabled():
                    # the + 0.01 * cab * se term rides on the tensor pipe: [att | cab] . [W_proj ; diag(0.01 * se_b)] per sample,
                    # so the layer keeps the plain (faster) residual epilogue
                    ops.build_concat_diag_weights(d["proj_w"], se, 0.01, wcat.view(B, CP, 2 * CP))
                    ops.conv_gemm(att, B, H, W, CP, wcat, n_store=CP, w_batch_rows=CP, bias=d["proj_b"], x2=cab2, res=src, out_f32=X, ln=ln2)
                else:
                    ops.conv_gemm(att, B, H, W, CP, d["proj_w"], n_store=CP, bias=d["proj_b"], aux=cab2, aux_chan=se, aux_alpha=0.01, res=src, out_f32=X, ln=ln2)
                self._mlp(d, X, t, h, B, H, W, M, t_ready=fused, next_norm=nxt if fused else None)
                t_ready = fused
                src = X
            d = layer["ocab"]
            if not t_ready:
                ops.layernorm(X, M, C, d["norm1"][0], d["norm1"][1], 1e-5, out_bf16=t, out_cols=CP)
            ops.conv_gemm(t, B, H, W, CP, d["qkv_w"], n_store=3 * CP, bias=d["qkv_b"], out_bf16=qkv)
            ops.window_attention(qkv, B, H, W, att, bias_table=d["table"], wh=WS, ww=WS, kh=24, kw=24, kpad=(4, 4),
                                 rel_sign=-1, rel_off=(-7, -7), rel_stride=39)
            if tail:
                ops.hab_tail(att, B, H, W, d["proj_w"], d["proj_b"], X, d["norm2"], d["fc1_w"], d["fc1_b"], d["fc2_w"], d["fc2_b"], X, out_bf16=Xb)
            else:
                ops.conv_gemm(att, B, H, W, CP, d["proj_w"], n_store=CP, bias=d["proj_b"], res=X, out_f32=X,
                              ln=(d["norm2"][0], d["norm2"][1], 1e-5, C, t) if fused else None)
                self._mlp(d, X, t, h, B, H, W, M, extra_bf16=Xb, t_ready=fused)
            # RHAG tail: conv3x3 + group residual (hat_arch.py:618-619); its epilogue emits the LayerNorm of the next consumer
            # of G: norm1 of the next group's first block, or the final `norm`
            nxt = self.layers[li + 1]["habs"][0]["norm1"] if li + 1 < len(self.layers) else self.norm
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
