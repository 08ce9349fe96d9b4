"""FreqFusion fusion head (eval path) on the ffb200 kernels.

Host-side mirror of phases 2-7 of `CompleteEnhancedFusionSR.forward` (reference src/models/enhanced_fusion.py:694-754;
process_frequency_bands :397-460, fuse_experts :502-591, apply_dynamic_selection :593-647, refine_output :653-688)
for the shipped MODEL_CONFIG (models/team29_FreqFusion/io.py:40-58).  Consumes the reference fusion state_dict unchanged;
the tensors that are dead at inference (collaborative.*, freq_router.*, expert_weights, band_importance) are ignored.

Precision plan: the LR-resolution routing path (frequency bands, band fusion, multiscale, selector, blend) is fp32;
conv chains run bf16 x bf16 -> fp32 on tensor cores.
"""
import ctypes as C_
import math

import torch
import torch.nn.functional as F

from . import lib as L
from . import ops
from .hat import Workspace
from .ops import ACT_CLAMP01, ACT_GELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, CONV_1X1, CONV_3X3
from .packing import (BF16, F32, pack_conv, pack_conv_direct, pack_conv_im2col2, pack_conv_split3, pack_dw, pack_matrix,
                      pack_vector)

DB4_LO = [-0.010597401784997278, 0.032883011666982945, 0.030841381835986965, -0.18703481171888114,
          -0.027983769416983849, 0.63088076792959036, 0.71484657055291582, 0.23037781330885523]
DB4_HI = [-0.23037781330885523, 0.71484657055291582, -0.63088076792959036, -0.027983769416983849,
          0.18703481171888114, 0.030841381835986965, -0.032883011666982945, -0.010597401784997278]
NBANDS = 9


def _dct_tables():
    n = 8
    d = torch.zeros(n, n, dtype=torch.float64)
    for k in range(n):
        for i in range(n):
            d[k, i] = math.sqrt(1.0 / n) if k == 0 else math.sqrt(2.0 / n) * math.cos(math.pi * k * (2 * i + 1) / (2 * n))
    # zigzag rank of each coefficient; low: rank < 21, mid: rank < 42, high: rest (multi_domain_frequency.py:105-116)
    rank = [[0] * n for _ in range(n)]
    idx = 0
    for s in range(2 * n - 1):
        lo, hi = max(0, s - n + 1), min(s, n - 1)
        order = range(hi, lo - 1, -1) if s % 2 == 0 else range(lo, hi + 1)
        for i in order:
            rank[i][s - i] = idx
            idx += 1
    band = torch.tensor([[0 if rank[i][j] < (n * n) // 3 else (1 if rank[i][j] < 2 * (n * n) // 3 else 2) for j in range(n)] for i in range(n)], dtype=torch.int32)
    return d.float().reshape(-1), band.reshape(-1)


def _bn_affine(g, p, eps=1e-5):
    a = g(p + "weight") / torch.sqrt(g(p + "running_var") + eps)
    return a, g(p + "bias") - g(p + "running_mean") * a


def _ptr(t):
    return C_.c_void_p(t.data_ptr())


class HeadRunner:
    def __init__(self, sd, device="cuda"):
        self.device = dev = device
        self.ws = Workspace(dev)
        g = lambda k: sd[k].detach().to("cpu", F32)
        self._g = g
        d = lambda t: t.to(dev).contiguous()
        # ---- frequency decomposition constants / parameters
        dm, bo = _dct_tables()
        self.dct_mat, self.dct_band = d(dm), d(bo)
        self.dct_scale = d(g("multi_domain_freq.dct.band_scale"))
        self.dwt_lo, self.dwt_hi = d(torch.tensor(DB4_LO)), d(torch.tensor(DB4_HI))
        self.dwt_scale = d(g("multi_domain_freq.dwt.subband_scale"))
        self.fft_scale = d(g("multi_domain_freq.fft.band_scale"))
        self._fft_logits = g("multi_domain_freq.fft.freq_mask_logits")
        self._fft_temp = g("multi_domain_freq.fft.temperature")
        self._fft_masks = {}
        # ---- cross-band attention + LKA
        p = "cross_band_attn."
        self.cb_proj_w, self.cb_proj_b = d(g(p + "band_proj.weight").reshape(64, 3)), d(g(p + "band_proj.bias"))
        self.cb_ln = (d(g(p + "norm.weight")), d(g(p + "norm.bias")))
        wi, bi = g(p + "band_attention.in_proj_weight").clone(), g(p + "band_attention.in_proj_bias").clone()
        wi[:64] *= 0.25   # head_dim 16 ** -0.5, applied to q by nn.MultiheadAttention
        bi[:64] *= 0.25
        self.cb_in_w, self.cb_in_b = pack_matrix(wi, 192, 64, device=dev), d(bi)
        self.cb_out_w, self.cb_out_b = pack_matrix(g(p + "band_attention.out_proj.weight"), 64, 64, device=dev), d(g(p + "band_attention.out_proj.bias"))
        q = p + "lka_block."
        a1, b1 = _bn_affine(g, q + "norm1.")
        self.lka_n1 = (d(a1), d(b1))
        tile9 = lambda w: pack_dw(w.repeat(NBANDS, 1, 1, 1), 64 * NBANDS, device=dev)
        self.lka_dw5, self.lka_dwh, self.lka_dwv = tile9(g(q + "lka.local_conv.weight")), tile9(g(q + "lka.h_conv.weight")), tile9(g(q + "lka.v_conv.weight"))
        ab, bb = _bn_affine(g, q + "lka.bn.")
        self.lka_pw_w = pack_matrix(g(q + "lka.pw_conv.weight").reshape(64, 64) * ab[:, None], 64, 64, device=dev)
        self.lka_pw_b = d(bb)
        self.lka_s1, self.lka_s2 = float(g(q + "scale1")), float(g(q + "scale2"))
        a2, b2 = _bn_affine(g, q + "norm2.")
        w0 = g(q + "ffn.0.weight").reshape(128, 64)
        self.lka_f0_w, self.lka_f0_b = pack_matrix(w0 * a2[None, :], 128, 64, device=dev), d(g(q + "ffn.0.bias") + w0 @ b2)
        self.lka_f2_w, self.lka_f2_b = pack_matrix(g(q + "ffn.2.weight").reshape(64, 128), 64, 128, device=dev), d(g(q + "ffn.2.bias"))
        self.cb_o_w, self.cb_o_b = pack_matrix(g(p + "out_proj.weight").reshape(3, 64), 16, 64, device=dev), pack_vector(g(p + "out_proj.bias"), 16, device=dev)
        # ---- adaptive band fusion
        p = "multi_domain_freq.band_fusion."
        wa = torch.zeros(16, 9, 27)
        ba = torch.zeros(16)
        for i in range(NBANDS):
            wa[i, :, 3 * i:3 * i + 3] = g(p + f"band_attention.{i}.conv.0.weight")[0].permute(1, 2, 0).reshape(9, 3)
            ba[i] = g(p + f"band_attention.{i}.conv.0.bias")[0]
        self.ba_w, self.ba_b = d(wa.reshape(16, 243)), d(ba)
        imp = torch.cat([F.softplus(g(p + "dct_importance")), F.softplus(g(p + "dwt_importance")), F.softplus(g(p + "fft_importance"))])
        imp = imp / (imp.sum() + 1e-8)
        blob = [imp]
        for name in ("fusion_transform", "fusion_gate"):
            blob += [g(p + name + ".0.weight").reshape(64, 27).reshape(-1), g(p + name + ".0.bias"), g(p + name + ".2.weight").reshape(9, 64).reshape(-1), g(p + name + ".2.bias")]
        blob += [g(p + "dct_residual.weight").reshape(81), g(p + "dct_residual.bias")]
        self.bf_blob = d(torch.cat(blob))
        # ---- multiscale (BN after ReLU folded forward into the 1x1 fusion) + selector, all fp32
        p = "multiscale."
        wf = g(p + "fusion.weight").reshape(64, 192)
        self.ms_conv, self.ms_mix = [], []
        bias_total = torch.zeros(64)
        for i, name in enumerate(("conv_1x", "conv_2x", "conv_4x")):
            self.ms_conv.append(pack_conv_direct(g(p + name + ".0.weight"), 64, dev))
            a, b = _bn_affine(g, p + name + ".2.")
            wpart = wf[:, 64 * i:64 * (i + 1)]
            self.ms_mix.append(d(wpart * a[None, :]))
            bias_total += wpart @ b
        self.ms_bias = d(bias_total)
        p = "dynamic_selector."
        self.de = [(pack_conv_direct(g(p + f"difficulty_estimator.{i}.weight"), co, dev), pack_vector(g(p + f"difficulty_estimator.{i}.bias"), co, device=dev))
                   for i, co in ((0, 64), (2, 32), (4, 8))]
        self.eg0 = (pack_conv_direct(g(p + "expert_gate.0.weight"), 64, dev), d(g(p + "expert_gate.0.bias")))
        self.eg2 = (pack_conv_direct(g(p + "expert_gate.2.weight"), 8, dev), pack_vector(g(p + "expert_gate.2.bias"), 8, device=dev))
        # the two 64-channel 3x3 layers of the selector run on the tensor cores with split-bf16 operands (hi/lo terms of the fp32
        # activations and weights, fp32 accumulation: ~16 mantissa bits, the hard >= 0.99*max mask downstream stays stable)
        self.de1_tc = (pack_conv_split3(g(p + "difficulty_estimator.2.weight"), 32, device=dev), pack_vector(g(p + "difficulty_estimator.2.bias"), 32, device=dev))
        self.eg0_tc = (pack_conv_split3(g(p + "expert_gate.0.weight"), 64, device=dev), d(g(p + "expert_gate.0.bias")))
        # ---- hierarchical fusion
        p = "multi_res_fusion."
        self.hier = []
        for n, cin_pad, c2 in ((1, 64, 64), (2, 128, 64), (3, 128, 32)):
            st = dict(
                c0_w=pack_conv(g(p + f"stage{n}_conv.0.weight"), 64, cin_pad, device=dev), c0_b=d(g(p + f"stage{n}_conv.0.bias")),
                c2_w=pack_conv(g(p + f"stage{n}_conv.2.weight"), 64, 64, device=dev), c2_b=pack_vector(g(p + f"stage{n}_conv.2.bias"), 64, device=dev),
                g_w1=d(g(p + f"stage{n}_gate.gate.0.weight").reshape(c2 // 4, c2)), g_b1=d(g(p + f"stage{n}_gate.gate.0.bias")),
                g_w2=d(g(p + f"stage{n}_gate.gate.2.weight").reshape(-1)), g_b2=float(g(p + f"stage{n}_gate.gate.2.bias")),
                r0_w=pack_conv(g(p + f"stage{n}_res.block.0.weight"), 64, 64, device=dev), r2_w=pack_conv(g(p + f"stage{n}_res.block.2.weight"), 64, 64, device=dev),
                scale=float(g(p + f"stage{n}_res.scale")), c=c2)
            self.hier.append(st)
        self.w12, self.w23 = float(g(p + "residual_weight_1_2")), float(g(p + "residual_weight_2_3"))
        self.rgb0_w, self.rgb0_b = pack_conv(g(p + "to_rgb.0.weight"), 64, 64, device=dev), pack_vector(g(p + "to_rgb.0.bias"), 64, device=dev)
        self.rgb2_w, self.rgb2_b = pack_conv(g(p + "to_rgb.2.weight"), 16, 64, device=dev), pack_vector(g(p + "to_rgb.2.bias"), 16, device=dev)
        # ---- refine net
        # 3 -> 64 first layers at HR: im2col of the 3x3 neighbourhood with split-bf16 activations (54 of one 64-wide k-block)
        self.rf0 = (pack_conv_im2col2(g("refine_net.0.weight"), 64, device=dev), d(g("refine_net.0.bias")))
        self.rf2 = (pack_conv(g("refine_net.2.weight"), 64, 64, device=dev), d(g("refine_net.2.bias")))
        self.rf4 = (pack_conv(g("refine_net.4.weight"), 64, 64, device=dev), d(g("refine_net.4.bias")))
        self.rf6 = (pack_conv(g("refine_net.6.weight"), 16, 64, device=dev), pack_vector(g("refine_net.6.bias"), 16, device=dev))
        self.residual_scale = float(g("residual_scale"))
        # ---- Laplacian edge refinement
        p = "edge_refine."
        c = torch.arange(5, dtype=F32) - 2
        k1 = torch.exp(-(c ** 2) / (2 * 1.5 ** 2))
        self.gauss = d(k1 / k1.sum())
        self.edge_levels = []
        for l in range(3):
            q = p + f"edge_refiners.{l}."
            self.edge_levels.append(dict(
                c1=(pack_conv_im2col2(g(q + "conv1.weight"), 64, device=dev), pack_vector(g(q + "conv1.bias"), 64, device=dev)),
                pj=(pack_conv_direct(g(q + "proj.weight"), 64, dev), pack_vector(g(q + "proj.bias"), 64, device=dev)),
                c2=(pack_conv(g(q + "conv2.weight"), 64, 64, device=dev), pack_vector(g(q + "conv2.bias"), 64, device=dev)),
                c3=(pack_conv(g(q + "conv3.weight"), 64, 64, device=dev), pack_vector(g(q + "conv3.bias"), 64, device=dev)),
                a0=(pack_conv_direct(g(q + "attn.attn.0.weight"), 8, dev), d(g(q + "attn.attn.0.bias"))),
                a2=(pack_conv_direct(g(q + "attn.attn.2.weight"), 8, dev), pack_vector(g(q + "attn.attn.2.bias"), 8, device=dev))))
        self.level_w = [float(v) for v in F.softmax(g(p + "level_weights"), 0)]
        self.ef0 = (pack_conv(g(p + "fusion.0.weight"), 64, 128, device=dev), pack_vector(g(p + "fusion.0.bias"), 64, device=dev))
        self.ef2 = (pack_conv(g(p + "fusion.2.weight"), 16, 64, device=dev), pack_vector(g(p + "fusion.2.bias"), 16, device=dev))
        self.egate0 = (pack_conv_direct(g(p + "edge_gate.0.weight"), 16, dev), d(g(p + "edge_gate.0.bias")))
        self.egate2 = (pack_conv_direct(g(p + "edge_gate.2.weight"), 8, dev), pack_vector(g(p + "edge_gate.2.bias"), 8, device=dev))
        self.edge_strength = float(g(p + "edge_strength"))

    # ------------------------------------------------------------------------------------------
    def _pack_collaborative(self):
        """Weights of `collaborative.*` (EnhancedCollaborativeWithLKA, large_kernel_attention.py:251-325); packed on first use --
        the branch only runs when expert features are passed to forward_with_precomputed."""
        if getattr(self, "_collab", None) is not None:
            return self._collab
        g, dev = self._g, self.device
        d = lambda t: t.to(dev).contiguous()
        p = "collaborative."
        c = {}
        c["align"] = [(pack_matrix(g(p + f"align_layers.{n}.weight").reshape(128, cin), 128, cp, device=dev), d(g(p + f"align_layers.{n}.bias")), cin, cp)
                      for n, cin, cp in (("hat", 180, 192), ("dat", 180, 192), ("nafnet", 64, 64))]
        c["ln1"] = (d(g(p + "norm1.weight")), d(g(p + "norm1.bias")))
        c["ln2"] = (d(g(p + "norm2.weight")), d(g(p + "norm2.bias")))
        wi, bi = g(p + "cross_attn.in_proj_weight").clone(), g(p + "cross_attn.in_proj_bias").clone()
        wi[:128] *= 0.25      # head_dim 16 ** -0.5, applied to q by nn.MultiheadAttention
        bi[:128] *= 0.25
        c["in"] = (pack_matrix(wi, 384, 128, device=dev), d(bi))
        c["out"] = (pack_matrix(g(p + "cross_attn.out_proj.weight"), 128, 128, device=dev), d(g(p + "cross_attn.out_proj.bias")))
        c["f0"] = (pack_matrix(g(p + "ffn.0.weight"), 256, 128, device=dev), d(g(p + "ffn.0.bias")))
        c["f2"] = (pack_matrix(g(p + "ffn.2.weight"), 128, 256, device=dev), d(g(p + "ffn.2.bias")))
        q = p + "lka_global."
        a1, b1 = _bn_affine(g, q + "norm1.")
        c["n1"] = (d(a1), d(b1))
        tile3 = lambda w: pack_dw(w.repeat(3, 1, 1, 1), 128 * 3, device=dev)
        c["dw5"], c["dwh"], c["dwv"] = tile3(g(q + "lka.local_conv.weight")), tile3(g(q + "lka.h_conv.weight")), tile3(g(q + "lka.v_conv.weight"))
        ab, bb = _bn_affine(g, q + "lka.bn.")
        c["pw"] = (pack_matrix(g(q + "lka.pw_conv.weight").reshape(128, 128) * ab[:, None], 128, 128, device=dev), d(bb))
        c["s1"], c["s2"] = float(g(q + "scale1")), float(g(q + "scale2"))
        a2, b2 = _bn_affine(g, q + "norm2.")
        w0 = g(q + "ffn.0.weight").reshape(256, 128)
        c["lf0"] = (pack_matrix(w0 * a2[None, :], 256, 128, device=dev), d(g(q + "ffn.0.bias") + w0 @ b2))
        c["lf2"] = (pack_matrix(g(q + "ffn.2.weight").reshape(128, 256), 128, 256, device=dev), d(g(q + "ffn.2.bias")))
        c["mod"] = [(pack_matrix(g(p + f"modulation.{i}.0.weight").reshape(32, 128), 32, 128, device=dev), d(g(p + f"modulation.{i}.0.bias")),
                     d(g(p + f"modulation.{i}.3.weight").reshape(3, 32)), d(g(p + f"modulation.{i}.3.bias"))) for i in range(3)]
        self._collab = c
        return c

    def collaborative(self, feats, stack, B, h, w, intermediates=None):
        """EnhancedCollaborativeWithLKA.forward (large_kernel_attention.py:327-419):
        feats = {'hat': [B,180,h,w], 'dat': [B,180,h,w], 'nafnet': [B,64,h,w]} fp32 NCHW on the device; as in the reference a map with
        other channel counts is truncated / zero padded, larger maps are resized to the smallest (which must be the LR size) and a
        missing expert contributes zeros.  Modulates channels 0-8 of the
        expert stack in place: out_e <- clamp(out_e * (1 + 0.2 (mod_e - 0.5)), 0, 1) with mod_e [B,3] from the e-th modulation head."""
        c = self._pack_collaborative()
        ws, lib, st, ck = self.ws, L.load(), ops._stream, L.check
        P = B * h * w
        T = 3 * P
        stk = ws.get("co_stk", P, 384, BF16)
        present = [n for n in ("hat", "dat", "nafnet") if feats.get(n) is not None]
        if not present:
            return stack                    # no features at all: the expert outputs pass through (large_kernel_attention.py:358-359)
        for n in present:
            f = feats[n]
            if not (torch.is_tensor(f) and f.is_cuda and f.dim() == 4 and f.shape[0] == B):
                raise L.FFError(f"expert_features['{n}'] must be a CUDA tensor [B={B}, C, H, W], got {tuple(f.shape) if torch.is_tensor(f) else type(f)}")
        # aligned maps are brought to the smallest spatial size among them (:362-369); the modulation heads run at 4x that size here
        mh, mw = min(feats[n].shape[2] for n in present), min(feats[n].shape[3] for n in present)
        if (mh, mw) != (h, w):
            raise L.FFError(f"the smallest expert feature map is {mh}x{mw}; it must have the LR size {h}x{w} (the SR outputs are 4x that)")
        for e, name in enumerate(("hat", "dat", "nafnet")):
            wgt, bias, cin, cp = c["align"][e]
            if name not in present:         # an expert without features contributes zeros (:374-377)
                stk[:, 128 * e:128 * (e + 1)].zero_()
                continue
            f = feats[name].float()
            if f.shape[1] > cin:            # too many channels: truncated; too few: zero padded (:346-355)
                f = f[:, :cin]
            elif f.shape[1] < cin:
                f = F.pad(f, (0, 0, 0, 0, 0, cin - f.shape[1]))
            f = f.contiguous()
            if tuple(f.shape[2:]) != (h, w):
                # the reference resizes the ALIGNED map; a 1x1 conv and a bilinear resize commute (the interpolation weights sum
                # to one, so the bias passes through), so the features are resized instead: every channel plane as a 1-channel image
                hf, wf = f.shape[2], f.shape[3]
                fr = ws.get(f"co_fr{e}", B * cin, h * w, F32)
                ck(lib.ff_bilinear_f32(_ptr(f), B * cin, hf, wf, 1, 1, _ptr(fr), h, w, 1, 0, None, st()), "ff_bilinear_f32")
                f = fr
            fb = ws.get(f"co_f{e}", P, cp, BF16)
            ck(lib.ff_nchw_to_nhwc_bf16(_ptr(f), B, cin, h, w, _ptr(fb), cp, st()), "ff_nchw_to_nhwc_bf16")
            ops.conv_gemm(fb, B, h, w, cp, wgt, n_store=128, bias=bias, out_bf16=stk[:, 128 * e:])
        Wt = 3 * w
        stk_t = stk.view(T, 128)
        nrm = ws.get("co_nrm", T, 128, BF16)
        qkv = ws.get("co_qkv", T, 384, BF16)
        att = ws.get("co_att", T, 128, BF16)
        A = ws.get("co_A", T, 128, BF16)
        A2 = ws.get("co_A2", T, 128, BF16)
        hid = ws.get("co_hid", T, 256, BF16)
        n1 = ws.get("co_n1", T, 128, BF16)
        d1 = ws.get("co_d1", T, 128, BF16)
        d2 = ws.get("co_d2", T, 128, BF16)
        ops.layernorm(stk_t, T, 128, c["ln1"][0], c["ln1"][1], 1e-5, out_bf16=nrm, out_cols=128)
        ops.conv_gemm(nrm, B, h, Wt, 128, c["in"][0], n_store=384, bias=c["in"][1], out_bf16=qkv)
        ck(lib.ff_token_attention(_ptr(qkv), C_.c_longlong(T), 3, 128, _ptr(att), st()), "ff_token_attention")
        ops.conv_gemm(att, B, h, Wt, 128, c["out"][0], n_store=128, bias=c["out"][1], res=stk_t, out_bf16=A)
        ops.layernorm(A, T, 128, c["ln2"][0], c["ln2"][1], 1e-5, out_bf16=nrm, out_cols=128)
        ops.conv_gemm(nrm, B, h, Wt, 128, c["f0"][0], n_store=256, bias=c["f0"][1], act=ACT_GELU, out_bf16=hid)
        ops.conv_gemm(hid, B, h, Wt, 256, c["f2"][0], n_store=128, bias=c["f2"][1], res=A, out_bf16=A2)
        # shared LKA block on the [B, h, w, 3*128] view (weights tiled over the three experts)
        ck(lib.ff_affine_rows(_ptr(A2), C_.c_longlong(T), 128, _ptr(c["n1"][0]), _ptr(c["n1"][1]), _ptr(n1), st()), "ff_affine_rows")
        CB = 384
        ops.dwconv(n1, B, h, w, CB, 5, 5, c["dw5"], None, d1.view(P, CB), x_ld=CB)
        ops.dwconv(d1, B, h, w, CB, 1, 21, c["dwh"], None, d2.view(P, CB), x_ld=CB)
        ops.dwconv(d2, B, h, w, CB, 21, 1, c["dwv"], None, d1.view(P, CB), x_ld=CB)
        ops.conv_gemm(d1, B, h, Wt, 128, c["pw"][0], n_store=128, bias=c["pw"][1], act=ACT_SIGMOID, alpha=c["s1"], mul=n1, res=A2, out_bf16=d2)
        ops.conv_gemm(d2, B, h, Wt, 128, c["lf0"][0], n_store=256, bias=c["lf0"][1], act=ACT_GELU, out_bf16=hid)
        ops.conv_gemm(hid, B, h, Wt, 128 * 2, c["lf2"][0], n_store=128, bias=c["lf2"][1], alpha=c["s2"], res=d2, out_bf16=A)
        # per-expert modulation heads: 1x1 conv at LR (it commutes with the bilinear x4), GELU + global mean at HR, 1x1 + sigmoid
        A3 = A.view(P, 384)
        gbuf = ws.get("co_g", P, 32, F32)
        NBLK = 64
        part = ws.get("co_part", B * NBLK, 32, F32)
        pooled = ws.get("co_pool", B, 32, F32)
        mods = ws.get("co_mod", B, 12, F32)
        for e in range(3):
            w0, b0, w3, b3 = c["mod"][e]
            ops.conv_gemm(A3[:, 128 * e:], B, h, w, 128, w0, n_store=32, bias=b0, out_f32=gbuf)
            ck(lib.ff_up_gelu_pool(_ptr(gbuf), 32, B, h, w, 32, 4, NBLK, _ptr(part), st()), "ff_up_gelu_pool")
            ops.gap_finalize(part, B, NBLK, 32, 1.0 / (16 * h * w), pooled)
            me = mods[:, 4 * e:]
            ops.vec_linear(pooled, B, 32, w3, b3, 3, ACT_SIGMOID, me, y_cols=3)
            ck(lib.ff_scale_clamp_channels(_ptr(stack), stack.stride(0), B, C_.c_longlong(16 * h * w), 3 * e, 3, _ptr(me), mods.stride(0), C_.c_float(0.9), C_.c_float(0.2), st()),
               "ff_scale_clamp_channels")
        if intermediates is not None:
            intermediates["modulation"] = mods.view(B, 3, 4)[:, :, :3].clone()
        return stack

    def _fft_mask(self, H, W):
        key = (H, W)
        if key not in self._fft_masks:
            m = F.interpolate(self._fft_logits, size=(H, W // 2 + 1), mode="bilinear", align_corners=False)
            m = torch.sigmoid(m * self._fft_temp.clamp(min=1.0)).reshape(H, W // 2 + 1)
            self._fft_masks[key] = m.to(self.device).contiguous()
        return self._fft_masks[key]

    # ------------------------------------------------------------------------------------------
    def forward(self, lr, stack, out=None, intermediates=None):
        """lr: fp32 NCHW [B,3,h,w] (any size); stack: fp32 [B*16*h*w][12] expert outputs (hat 0-2, dat 3-5, nafnet 6-8).
        Returns fp32 NCHW [B,3,4h,4w].  `intermediates` (dict) receives band_features [P][9] and fused_before_refine [P_hr][4]."""
        B, _, h, w = lr.shape
        if min(h, w) < 8:
            raise ValueError(f"HeadRunner: image {h}x{w} is smaller than the DWT's reflect padding (the reference's F.pad fails here too)")
        H, W = 4 * h, 4 * w
        P, PH = B * h * w, B * H * W
        ws, lib, st = self.ws, L.load(), ops._stream
        ck = L.check
        if out is None:
            out = torch.empty(B, 3, H, W, dtype=F32, device=self.device)

        lrn = ws.get("lrn", P, 4, F32)
        ops.nchw_to_nhwc(lr, lrn)
        # ---------------- phase 2: 9 frequency bands
        bands = ws.get("bands", P, 27, F32)
        scratch = ws.get("fscratch", 1, B * 3 * (h * (w // 2 + 1) * 4 + 4 * ((h + 6) // 2 + 1) * ((w + 6) // 2 + 1)), F32)
        ck(lib.ff_freq_decompose(_ptr(lr), B, h, w, _ptr(self.dct_mat), _ptr(self.dct_band), _ptr(self.dct_scale), _ptr(self.dwt_lo), _ptr(self.dwt_hi),
                                 _ptr(self.dwt_scale), _ptr(self._fft_mask(h, w)), _ptr(self.fft_scale), _ptr(bands), _ptr(scratch),
                                 C_.c_size_t(scratch.numel() * 4), st()), "ff_freq_decompose")
        # ---------------- phase 3: cross-band attention + shared LKA block (tokens = pixel x band)
        T = P * NBANDS
        Wt = w * NBANDS
        stk = ws.get("cb_stk", T, 64, BF16)
        nrm = ws.get("cb_nrm", T, 64, BF16)
        qkv = ws.get("cb_qkv", T, 192, BF16)
        att = ws.get("cb_att", T, 64, BF16)
        A = ws.get("cb_A", T, 64, BF16)
        n1 = ws.get("cb_n1", T, 64, BF16)
        d1 = ws.get("cb_d1", T, 64, BF16)
        d2 = ws.get("cb_d2", T, 64, BF16)
        hid = ws.get("cb_hid", T, 128, BF16)
        ck(lib.ff_cb_embed_ln(_ptr(bands), C_.c_longlong(T), _ptr(self.cb_proj_w), _ptr(self.cb_proj_b), _ptr(self.cb_ln[0]), _ptr(self.cb_ln[1]),
                              _ptr(stk), _ptr(nrm), st()), "ff_cb_embed_ln")
        ops.conv_gemm(nrm, B, h, Wt, 64, self.cb_in_w, n_store=192, bias=self.cb_in_b, out_bf16=qkv)
        ck(lib.ff_cb_attention(_ptr(qkv), C_.c_longlong(T), NBANDS, _ptr(att), st()), "ff_cb_attention")
        ops.conv_gemm(att, B, h, Wt, 64, self.cb_out_w, n_store=64, bias=self.cb_out_b, res=stk, out_bf16=A)
        ck(lib.ff_affine_rows(_ptr(A), C_.c_longlong(T), 64, _ptr(self.lka_n1[0]), _ptr(self.lka_n1[1]), _ptr(n1), st()), "ff_affine_rows")
        # depthwise chain on the [B, h, w, 9*64] view (weights tiled over the 9 bands)
        CB = 64 * NBANDS
        ops.dwconv(n1, B, h, w, CB, 5, 5, self.lka_dw5, None, d1.view(P, CB), x_ld=CB)
        ops.dwconv(d1, B, h, w, CB, 1, 21, self.lka_dwh, None, d2.view(P, CB), x_ld=CB)
        ops.dwconv(d2, B, h, w, CB, 21, 1, self.lka_dwv, None, d1.view(P, CB), x_ld=CB)
        # x1 = x + scale1 * n1 * sigmoid(bn(pw(.)))
        ops.conv_gemm(d1, B, h, Wt, 64, self.lka_pw_w, n_store=64, bias=self.lka_pw_b, act=ACT_SIGMOID, alpha=self.lka_s1, mul=n1, res=A, out_bf16=d2)
        ops.conv_gemm(d2, B, h, Wt, 64, self.lka_f0_w, n_store=128, bias=self.lka_f0_b, act=ACT_GELU, out_bf16=hid)
        ops.conv_gemm(hid, B, h, Wt, 128, self.lka_f2_w, n_store=64, bias=self.lka_f2_b, alpha=self.lka_s2, res=d2, out_bf16=A)
        bands2 = ws.get("bands2", P, 27, F32)
        ops.conv_gemm(A, B, h, Wt, 64, self.cb_o_w, n_store=3, bias=self.cb_o_b, res=bands.view(T, 3), out_f32=bands2.view(T, 3))
        # ---------------- adaptive band fusion 9 -> 3 and frequency guidance
        batt = ws.get("batt", P, 16, F32)
        ops.conv_direct(bands2, B, h, w, 27, 3, self.ba_w, self.ba_b, n_store=16, act=ACT_SIGMOID, out_f32=batt)
        feats = ws.get("bfeat", P, 9, F32)
        guid = ws.get("guid", P, 4, F32)
        ck(lib.ff_band_fuse(_ptr(bands2), _ptr(batt), 16, C_.c_longlong(P), _ptr(self.bf_blob), self.bf_blob.numel(), _ptr(feats), _ptr(guid), st()), "ff_band_fuse")
        # ---------------- multiscale routing features + dynamic expert selector (fp32)
        ms = ws.get("ms", P, 64, F32)
        r = ws.get("ms_r", P, 64, F32)
        ops.conv_direct(lrn, B, h, w, 3, 3, self.ms_conv[0], None, n_store=64, act=ACT_RELU, out_f32=r)
        ops.conv_direct(r, B, h, w, 64, 1, self.ms_mix[0], self.ms_bias, n_store=64, out_f32=ms)
        for i, f in ((1, 2), (2, 4)):
            # F.interpolate(scale_factor=1/f) (fusion_network.py:594,599): output floor(size / f), source coordinates scaled by
            # exactly f -- not by the size ratio, which differs on sizes f does not divide
            hs, wsz = h // f, w // f
            lrs = ws.get(f"lr_d{f}", B * hs * wsz, 4, F32)
            ck(lib.ff_bilinear_f32_scaled(_ptr(lrn), B, h, w, 4, 3, _ptr(lrs), hs, wsz, 4, C_.c_float(float(f)), C_.c_float(float(f)), st()), "ff_bilinear_f32_scaled")
            rs = ws.get(f"ms_r{f}", B * hs * wsz, 64, F32)
            mx = ws.get(f"ms_m{f}", B * hs * wsz, 64, F32)
            ops.conv_direct(lrs, B, hs, wsz, 3, 3, self.ms_conv[i], None, n_store=64, act=ACT_RELU, out_f32=rs)
            ops.conv_direct(rs, B, hs, wsz, 64, 1, self.ms_mix[i], None, n_store=64, out_f32=mx)
            ck(lib.ff_bilinear_f32(_ptr(mx), B, hs, wsz, 64, 64, _ptr(ms), h, w, 64, 1, None, st()), "ff_bilinear_f32")
        gd = ws.get("gd", P, 4, F32)
        t64 = ws.get("sel64", P, 64, F32)
        t32 = ws.get("sel32", P, 32, F32)
        ops.conv_direct(lrn, B, h, w, 3, 3, self.de[0][0], self.de[0][1], n_store=64, act=ACT_RELU, out_f32=t64)
        sp = ws.get("split192", P, 192, BF16)
        ops.pack_taps(t64, B, h, w, 64, 1, 3, sp)
        ops.conv_gemm(sp, B, h, w, 192, self.de1_tc[0], kind=CONV_3X3, n_store=32, bias=self.de1_tc[1], act=ACT_RELU, out_f32=t32)
        ops.conv_direct(t32, B, h, w, 32, 3, self.de[2][0], self.de[2][1], n_store=1, act=ACT_SIGMOID, out_f32=gd, out_f32_off=3)
        ops.pack_taps(ms, B, h, w, 64, 1, 3, sp)
        ops.conv_gemm(sp, B, h, w, 192, self.eg0_tc[0], kind=CONV_3X3, n_store=64, bias=self.eg0_tc[1], act=ACT_RELU, out_f32=t64)
        ops.conv_direct(t64, B, h, w, 64, 1, self.eg2[0], self.eg2[1], n_store=3, act=ACT_SIGMOID, out_f32=gd)
        ck(lib.ff_selector_tail(_ptr(gd), C_.c_longlong(P), st()), "ff_selector_tail")
        # ---------------- hierarchical multi-resolution fusion (1/4 -> 1/2 -> 1x)
        in1 = ws.get("h_in1", PH // 16, 64, BF16)
        in2 = ws.get("h_in2", PH // 4, 128, BF16)
        in3 = ws.get("h_in3", PH, 128, BF16)
        ck(lib.ff_experts_resize(_ptr(stack), stack.stride(0), B, H, W, 4, _ptr(in1), 64, 0, st()), "ff_experts_resize")
        ck(lib.ff_experts_resize(_ptr(stack), stack.stride(0), B, H, W, 2, _ptr(in2), 128, 64, st()), "ff_experts_resize")
        ck(lib.ff_experts_resize(_ptr(stack), stack.stride(0), B, H, W, 1, _ptr(in3), 128, 64, st()), "ff_experts_resize")
        prev_up = None
        f_prev = None
        for n, (xin, cin, div) in enumerate(((in1, 64, 4), (in2, 128, 2), (in3, 128, 1))):
            sp = self.hier[n]
            Hs, Ws_, Ps = H // div, W // div, PH // (div * div)
            fa = ws.get(f"h_fa{n}", Ps, 64, BF16)
            fb = ws.get(f"h_fb{n}", Ps, 64, BF16)
            fc = ws.get(f"h_fc{n}", Ps, 64, BF16)
            ops.conv_gemm(xin, B, Hs, Ws_, cin, sp["c0_w"], kind=CONV_3X3, n_store=64, bias=sp["c0_b"], act=ACT_GELU, out_bf16=fa)
            ops.conv_gemm(fa, B, Hs, Ws_, 64, sp["c2_w"], kind=CONV_3X3, n_store=64, bias=sp["c2_b"], act=ACT_GELU, out_bf16=fb)
            ck(lib.ff_pixel_gate(_ptr(fb), 64, C_.c_longlong(Ps), sp["c"], _ptr(sp["g_w1"]), _ptr(sp["g_b1"]), _ptr(sp["g_w2"]), C_.c_float(sp["g_b2"]), st()), "ff_pixel_gate")
            ops.conv_gemm(fb, B, Hs, Ws_, 64, sp["r0_w"], kind=CONV_3X3, n_store=64, act=ACT_GELU, out_bf16=fa)
            if n == 0:
                ops.conv_gemm(fa, B, Hs, Ws_, 64, sp["r2_w"], kind=CONV_3X3, n_store=64, alpha=sp["scale"], res=fb, out_bf16=fc)
            else:
                ops.conv_gemm(fa, B, Hs, Ws_, 64, sp["r2_w"], kind=CONV_3X3, n_store=64, alpha=sp["scale"], res=fb, aux=xin,
                              aux_alpha=self.w12 if n == 1 else self.w23, out_bf16=fc)
            if n < 2:
                nxt = in2 if n == 0 else in3
                ck(lib.ff_bilinear_up2_bf16(_ptr(fc), B, Hs, Ws_, 64, 64, _ptr(nxt), 128, st()), "ff_bilinear_up2_bf16")
        rgb = ws.get("h_rgb", PH, 64, BF16)
        hier = ws.get("hier", PH, 4, F32)
        ops.conv_gemm(fc, B, H, W, 64, self.rgb0_w, kind=CONV_3X3, n_store=64, bias=self.rgb0_b, act=ACT_GELU, out_bf16=rgb)
        ops.conv_gemm(rgb, B, H, W, 64, self.rgb2_w, kind=CONV_3X3, n_store=3, bias=self.rgb2_b, act=ACT_SIGMOID, out_f32=hier)
        # ---------------- blend (frequency guidance + dynamic selection) at HR
        fused = ws.get("fused", PH, 4, F32)
        base = ws.get("base", PH, 4, F32)
        ck(lib.ff_blend(_ptr(stack), stack.stride(0), _ptr(hier), _ptr(guid), _ptr(gd), _ptr(lrn), B, h, w, C_.c_float(self.residual_scale),
                        _ptr(fused), _ptr(base), st()), "ff_blend")
        # ---------------- refine net + bilinear LR residual + clamp
        ra = ws.get("rf_a", PH, 64, BF16)
        rb = ws.get("rf_b", PH, 64, BF16)
        se = ws.get("se", PH, 8, F32)
        im = ws.get("im2col", PH, 64, BF16)
        ops.pack_taps(fused, B, H, W, 3, 3, 2, im)
        ops.conv_gemm(im, B, H, W, 64, self.rf0[0], kind=CONV_1X1, n_store=64, bias=self.rf0[1], act=ACT_GELU, out_bf16=ra)
        ops.conv_gemm(ra, B, H, W, 64, self.rf2[0], kind=CONV_3X3, n_store=64, bias=self.rf2[1], act=ACT_GELU, out_bf16=rb)
        ops.conv_gemm(rb, B, H, W, 64, self.rf4[0], kind=CONV_3X3, n_store=64, bias=self.rf4[1], act=ACT_GELU, out_bf16=ra)
        ops.conv_gemm(ra, B, H, W, 64, self.rf6[0], kind=CONV_3X3, n_store=3, bias=self.rf6[1], alpha=0.1, res=base, post_act=ACT_CLAMP01, out_f32=se)
        # ---------------- Laplacian pyramid edge refinement
        dn1 = ws.get("e_dn1", PH // 4, 4, F32)
        dn2 = ws.get("e_dn2", PH // 16, 4, F32)
        lap0 = ws.get("e_lap0", PH, 4, F32)
        lap1 = ws.get("e_lap1", PH // 4, 4, F32)
        ck(lib.ff_gauss_down(_ptr(se), 8, B, H, W, _ptr(self.gauss), _ptr(dn1), 4, st()), "ff_gauss_down")
        ck(lib.ff_lap_sub(_ptr(se), 8, _ptr(dn1), 4, B, H, W, _ptr(lap0), 4, st()), "ff_lap_sub")
        ck(lib.ff_gauss_down(_ptr(dn1), 4, B, H // 2, W // 2, _ptr(self.gauss), _ptr(dn2), 4, st()), "ff_gauss_down")
        ck(lib.ff_lap_sub(_ptr(dn1), 4, _ptr(dn2), 4, B, H // 2, W // 2, _ptr(lap1), 4, st()), "ff_lap_sub")
        allf = ws.get("e_all", PH, 128, BF16)
        for l, lap in enumerate((lap0, lap1, dn2)):
            e = self.edge_levels[l]
            Hl, Wl, Pl = H >> l, W >> l, PH >> (2 * l)
            c1 = ws.get(f"e_c1_{l}", Pl, 64, BF16)
            idn = ws.get(f"e_id_{l}", Pl, 64, BF16)
            c2 = ws.get(f"e_c2_{l}", Pl, 64, BF16)
            a8 = ws.get(f"e_a8_{l}", Pl, 8, F32)
            am = ws.get(f"e_am_{l}", Pl, 1, F32)
            ops.pack_taps(lap, B, Hl, Wl, 3, 3, 2, im)
            ops.conv_gemm(im, B, Hl, Wl, 64, e["c1"][0], kind=CONV_1X1, n_store=64, bias=e["c1"][1], act=ACT_GELU, out_bf16=c1)
            ops.conv_direct(lap, B, Hl, Wl, 3, 1, e["pj"][0], e["pj"][1], n_store=64, out_bf16=idn)
            ops.conv_gemm(c1, B, Hl, Wl, 64, e["c2"][0], kind=CONV_3X3, n_store=64, bias=e["c2"][1], act=ACT_GELU, out_bf16=c2)
            ops.conv_gemm(c2, B, Hl, Wl, 64, e["c3"][0], kind=CONV_3X3, n_store=64, bias=e["c3"][1], res=idn, out_bf16=c1)
            ops.conv_direct(c1, B, Hl, Wl, 32, 1, e["a0"][0], e["a0"][1], n_store=8, act=ACT_GELU, out_f32=a8)
            ops.conv_direct(a8, B, Hl, Wl, 8, 3, e["a2"][0], e["a2"][1], n_store=1, act=ACT_SIGMOID, out_f32=am)
            ck(lib.ff_edge_merge(_ptr(c1), 64, _ptr(am), B, Hl, Wl, 32, C_.c_float(self.level_w[l]), _ptr(allf), H, W, 128, 32 * l, st()), "ff_edge_merge")
        ops.conv_gemm(allf, B, H, W, 128, self.ef0[0], kind=CONV_3X3, n_store=64, bias=self.ef0[1], act=ACT_GELU, out_bf16=ra)
        ops.conv_gemm(ra, B, H, W, 64, self.ef2[0], kind=CONV_3X3, n_store=3, bias=self.ef2[1], out_f32=se[:, 3:])
        g16 = ws.get("e_g16", PH, 16, F32)
        gate = ws.get("e_gate", PH, 1, F32)
        ops.conv_direct(se, B, H, W, 6, 3, self.egate0[0], self.egate0[1], n_store=16, act=ACT_GELU, out_f32=g16)
        ops.conv_direct(g16, B, H, W, 16, 3, self.egate2[0], self.egate2[1], n_store=1, act=ACT_SIGMOID, out_f32=gate)
        ck(lib.ff_edge_final(_ptr(se), _ptr(gate), 1, B, H, W, C_.c_float(self.edge_strength), _ptr(out), st()), "ff_edge_final")
        if intermediates is not None:
            intermediates["band_features"] = feats
            intermediates["fused_before_refine"] = fused
            intermediates["bands_raw"] = bands
            intermediates["bands_attended"] = bands2
        return out
