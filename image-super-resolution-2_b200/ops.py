"""Thin tensor-level wrappers over the C ABI (include/ffb200.h).

torch is used for device memory and streams only; every arithmetic op on the hot path is one of
the hand-written kernels in csrc/.  All activations are NHWC; `ld` is the channel pitch.
"""
import ctypes as C

import torch

from . import lib as L
from .lib import (ACT_CLAMP01, ACT_GELU, ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, CONV_1X1, CONV_2X2S2, CONV_3X3)

_BF16 = torch.bfloat16
_F32 = torch.float32


def fused_ln_enabled():
    """The LayerNorm that follows a residual add is emitted by the producing GEMM's epilogue (FFConvGemm.ln_*);
    FFB200_FUSED_LN=0 restores the separate ff_layernorm passes (A/B measurements)."""
    import os
    return os.environ.get("FFB200_FUSED_LN", "1") != "0"


class KernelProfile:
    """Optional per-launch CUDA-event timing of ff_conv_gemm (bench.py's roofline leg; off on the hot path)."""

    def __init__(self):
        self.records = []   # (start_event, end_event, algorithmic_flops, executed_flops, algorithmic_bytes)

    def summary(self):
        torch.cuda.synchronize()
        ms = sum(r[0].elapsed_time(r[1]) for r in self.records)
        return dict(launches=len(self.records), ms=ms, algo_flops=sum(r[2] for r in self.records), exec_flops=sum(r[3] for r in self.records),
                    algo_bytes=sum(r[4] for r in self.records))


PROFILE = None   # set to a KernelProfile instance to record


# torch.cuda.current_stream() costs ~4 us of Python per call (device-index plumbing) and every launch asks for it: 1 358 times per
# forward.  The raw accessors below return the same cudaStream_t (they are what torch.cuda.current_stream wraps).
_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_RAW_DEVICE = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    if _RAW_STREAM is not None and _RAW_DEVICE is not None:
        return C.c_void_p(_RAW_STREAM(_RAW_DEVICE()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.FFError("ffb200 kernels need CUDA tensors: there is no CPU fallback")


def conv_gemm(x, B, H, W, cin, w, *, kind=CONV_1X1, n_store, bias=None, act=ACT_NONE, alpha=1.0, col_scale=None,
              mul=None, aux=None, aux_chan=None, aux_alpha=1.0, res=None, post_act=ACT_NONE, out_bf16=None,
              out_f32=None, pixel_shuffle=0, gate_pairs=0, w_batch_rows=0, x_ld=None, debug_simt=0, col_sums=None, ln=None, out_crop=None, x2=None):
    """Implicit-GEMM conv / linear on tcgen05 (see ff_conv_gemm in include/ffb200.h).

    x: bf16 tensor whose last dim is the channel pitch (or pass x_ld); w: packed bf16 [n_pad, taps*cin].
    Operand tensors (mul/aux/res/out_*) are 2-D-viewable [pixels, ld]; their ld is the last-dim stride owner.
    """
    _req_cuda(x, w, bias, col_scale, mul, aux, aux_chan, res, out_bf16, out_f32)
    p = L.FFConvGemm()
    p.x = x.data_ptr(); p.B, p.H, p.W = B, H, W
    p.x_ld = x_ld if x_ld is not None else x.stride(-2)
    p.cin = cin; p.kind = kind
    p.w = w.data_ptr(); p.n_pad = w.shape[0] if not w_batch_rows else w_batch_rows; p.n_store = n_store
    p.bias = bias.data_ptr() if bias is not None else None
    p.act = act; p.alpha = alpha
    p.col_scale = col_scale.data_ptr() if col_scale is not None else None
    if mul is not None:
        p.mul = mul.data_ptr(); p.mul_ld = mul.stride(-2)
    if aux is not None:
        p.aux = aux.data_ptr(); p.aux_ld = aux.stride(-2)
    if aux_chan is not None:
        p.aux_chan = aux_chan.data_ptr(); p.aux_chan_ld = aux_chan.stride(0)
    p.aux_alpha = aux_alpha
    if res is not None:
        p.res = res.data_ptr(); p.res_ld = res.stride(-2); p.res_is_f32 = 1 if res.dtype == _F32 else 0
    p.post_act = post_act
    if out_bf16 is not None:
        p.out_bf16 = out_bf16.data_ptr(); p.out_ld = out_bf16.stride(-2)
    if out_f32 is not None:
        p.out_f32 = out_f32.data_ptr(); p.out_f32_ld = out_f32.stride(-2)
    p.pixel_shuffle = pixel_shuffle; p.gate_pairs = gate_pairs; p.w_batch_rows = w_batch_rows
    p.debug_simt = debug_simt
    if col_sums is not None:
        _req_cuda(col_sums)
        p.col_sums = col_sums.data_ptr()
    if out_crop is not None:
        p.out_crop_h, p.out_crop_w = out_crop
    if x2 is not None:
        _req_cuda(x2)
        p.x2 = x2.data_ptr(); p.x2_ld = x2.stride(-2); p.cin2 = x2.shape[-1]
    if ln is not None:
        # fused LayerNorm of the updated residual row: ln = (gamma [n_store], beta [n_store], eps, real channel count, bf16 out)
        g_, b_, eps_, cols_, lo_ = ln
        _req_cuda(g_, b_, lo_)
        p.ln_gamma = g_.data_ptr(); p.ln_beta = b_.data_ptr(); p.ln_eps = eps_; p.ln_cols = cols_
        p.ln_out = lo_.data_ptr(); p.ln_out_ld = lo_.stride(-2)
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.load().ff_conv_gemm(C.byref(p), _stream()), "ff_conv_gemm")
        e1.record()
        taps = {CONV_1X1: 1, CONV_3X3: 9, CONV_2X2S2: 4}[kind]
        Mo = B * H * W // (4 if kind == CONV_2X2S2 else 1)
        n_real, k_real = getattr(w, "ff_real", (p.n_pad, taps * cin))
        # compulsory HBM bytes of this launch: A once (its real channels), weights, every epilogue operand / output at its dtype
        cin_real = k_real // taps
        width = n_real // (2 if gate_pairs else 1)
        if x2 is not None:
            byts_x2 = B * H * W * n_real * 2
        byts = (byts_x2 if x2 is not None else 0) + B * H * W * cin_real * 2 + n_real * k_real * 2 + Mo * width * ((2 if out_bf16 is not None else 0) + (4 if out_f32 is not None else 0))
        byts += Mo * width * ((4 if res.dtype == _F32 else 2) if res is not None else 0) + Mo * width * (2 if mul is not None else 0) + Mo * width * (2 if aux is not None else 0)
        byts += Mo * ln[3] * 2 if ln is not None else 0
        PROFILE.records.append((e0, e1, 2.0 * Mo * n_real * k_real, 2.0 * Mo * p.n_pad * taps * cin, float(byts),
                                (kind, cin, p.n_pad, B, H, W, act, res is not None, aux is not None, mul is not None, gate_pairs, pixel_shuffle, out_f32 is not None)))
        return
    L.check(L.load().ff_conv_gemm(C.byref(p), _stream()), "ff_conv_gemm")


def concat_aux_enabled():
    """HAT's proj + 0.01 * cab * se as one K-concatenated GEMM (FFConvGemm.x2); FFB200_CONCAT_AUX=0 restores the aux epilogue."""
    import os
    return os.environ.get("FFB200_CONCAT_AUX", "1") != "0"


def mlp_fused_enabled():
    """fc1 + GELU + fc2 (+ residual, + next LayerNorm) as one kernel with the hidden tile on chip (ff_mlp_fused);
    FFB200_FUSED_MLP=0 restores the two conv_gemm launches."""
    import os
    return os.environ.get("FFB200_FUSED_MLP", "1") != "0"


def mlp_fused(t, B, H, W, w1, b1, w2, b2, x, *, out_bf16=None, ln=None):
    """x += fc2(GELU(fc1(t))) in place on the fp32 residual stream x; ln = (gamma, beta, eps, cols, bf16 out) as in conv_gemm."""
    _req_cuda(t, w1, b1, w2, b2, x, out_bf16)
    p = L.FFMlpFused()
    p.t = t.data_ptr(); p.t_ld = t.stride(-2)
    p.B, p.H, p.W = B, H, W
    p.w1 = w1.data_ptr(); p.b1 = b1.data_ptr(); p.w2 = w2.data_ptr(); p.b2 = b2.data_ptr()
    p.x = x.data_ptr(); p.x_ld = x.stride(-2)
    if out_bf16 is not None:
        p.out_bf16 = out_bf16.data_ptr(); p.out_ld = out_bf16.stride(-2)
    if ln is not None:
        g_, b_, eps_, cols_, lo_ = ln
        _req_cuda(g_, b_, lo_)
        p.ln_gamma = g_.data_ptr(); p.ln_beta = b_.data_ptr(); p.ln_eps = eps_; p.ln_cols = cols_
        p.ln_out = lo_.data_ptr(); p.ln_out_ld = lo_.stride(-2)
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.load().ff_mlp_fused(C.byref(p), _stream()), "ff_mlp_fused")
        e1.record()
        M = B * H * W
        n1, k1 = getattr(w1, "ff_real", tuple(w1.shape))
        n2, k2 = getattr(w2, "ff_real", tuple(w2.shape))
        flops = 2.0 * M * (n1 * k1 + n2 * k2)
        byts = M * k1 * 2 + (n1 * k1 + n2 * k2) * 2 + M * n2 * 8 + (M * n2 * 2 if out_bf16 is not None else 0) + (M * ln[3] * 2 if ln is not None else 0)
        PROFILE.records.append((e0, e1, flops, 2.0 * M * (w1.shape[0] * w1.shape[1] + w2.shape[0] * w2.shape[1]), float(byts), ("mlp_fused", B, H, W)))
        return
    L.check(L.load().ff_mlp_fused(C.byref(p), _stream()), "ff_mlp_fused")


def hab_tail_enabled():
    """proj + shortcut + LayerNorm2 + MLP + residual (+ next LayerNorm) of a HAT block as one kernel (ff_hab_tail);
    FFB200_HAB_TAIL=0 restores the residual GEMM followed by ff_mlp_fused."""
    import os
    return os.environ.get("FFB200_HAB_TAIL", "1") != "0"


def hab_tail(a0, B, H, W, wp, bp, res, ln2, w1, b1, w2, b2, x, *, a1=None, a1_diag=None, a1_alpha=1.0, wp_batch_rows=0, out_bf16=None, ln=None, eps=1e-5, cols=180):
    """x = x1 + fc2(GELU(fc1(LN2(x1)))) with x1 = res + [a0 | a1] . wp^T + bp; ln = (gamma, beta, bf16 out) of the next LayerNorm."""
    _req_cuda(a0, a1, wp, bp, res, ln2[0], ln2[1], w1, b1, w2, b2, x, out_bf16)
    p = L.FFHabTail()
    p.a0 = a0.data_ptr(); p.a0_ld = a0.stride(-2)
    if a1 is not None:
        p.a1 = a1.data_ptr(); p.a1_ld = a1.stride(-2)
        if a1_diag is not None:      # per-sample channel scale of the a1 term: the diagonal K block is generated inside the kernel
            _req_cuda(a1_diag)
            p.a1_diag = a1_diag.data_ptr(); p.a1_diag_ld = a1_diag.stride(0); p.a1_alpha = a1_alpha
    p.B, p.H, p.W = B, H, W
    p.wp = wp.data_ptr(); p.wp_batch_rows = wp_batch_rows; p.bp = bp.data_ptr()
    p.res = res.data_ptr(); p.res_ld = res.stride(-2)
    p.ln2_gamma = ln2[0].data_ptr(); p.ln2_beta = ln2[1].data_ptr()
    p.w1 = w1.data_ptr(); p.b1 = b1.data_ptr(); p.w2 = w2.data_ptr(); p.b2 = b2.data_ptr()
    p.x = x.data_ptr(); p.x_ld = x.stride(-2)
    if out_bf16 is not None:
        p.out_bf16 = out_bf16.data_ptr(); p.out_ld = out_bf16.stride(-2)
    if ln is not None:
        g_, b_, lo_ = ln
        _req_cuda(g_, b_, lo_)
        p.ln_gamma = g_.data_ptr(); p.ln_beta = b_.data_ptr(); p.ln_out = lo_.data_ptr(); p.ln_out_ld = lo_.stride(-2)
    p.ln_eps = eps; p.ln_cols = cols
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.load().ff_hab_tail(C.byref(p), _stream()), "ff_hab_tail")
        e1.record()
        M = B * H * W
        n0, k0 = getattr(wp, "ff_real", (cols, cols))
        n1, k1 = getattr(w1, "ff_real", tuple(w1.shape))
        n2, k2 = getattr(w2, "ff_real", tuple(w2.shape))
        flops = 2.0 * M * (n0 * k0 + n1 * k1 + n2 * k2 + (cols if a1 is not None else 0))
        executed = 2.0 * M * (192 * 192 + (3 * 64 * 64 if a1 is not None else 0) + w1.shape[0] * w1.shape[1] + w2.shape[0] * w2.shape[1])
        byts = M * k0 * 2 * (2 if a1 is not None else 1) + (n0 * k0 + n1 * k1 + n2 * k2) * 2 + M * n2 * 8 + (M * n2 * 2 if out_bf16 is not None else 0) + (M * cols * 2 if ln is not None else 0)
        PROFILE.records.append((e0, e1, flops, executed, float(byts), ("hab_tail", B, H, W)))
        return
    L.check(L.load().ff_hab_tail(C.byref(p), _stream()), "ff_hab_tail")


def naf_tail_enabled():
    """conv3 + residual + norm2 + conv4 + SimpleGate + conv5 + residual (+ next norm1) of a 64-channel NAFBlock as one kernel
    (ff_naf_tail); FFB200_NAF_TAIL=0 restores the three ff_conv_gemm passes."""
    import os
    return os.environ.get("FFB200_NAF_TAIL", "1") != "0"


def naf_tail(a0, B, H, W, w3, b3, res, ln2, w4, b4, w5, b5, x, *, w3_batch_rows=0, out_bf16=None, ln=None, eps=1e-6):
    """x = y + (u1 * u2) . w5^T + b5 with y = res + a0 . w3^T + b3 and u = LN2(y) . w4^T + b4 (64 channels; beta / gamma / sca folded
    into w3 / w5 by the caller); ln = (gamma, beta) of the next LayerNorm2d, written to out_bf16 instead of a plain bf16 copy."""
    _req_cuda(a0, w3, b3, res, ln2[0], ln2[1], w4, b4, w5, b5, x, out_bf16)
    p = L.FFNafTail()
    p.a0 = a0.data_ptr(); p.a0_ld = a0.stride(-2)
    p.B, p.H, p.W = B, H, W
    p.w3 = w3.data_ptr(); p.w3_batch_rows = w3_batch_rows; p.b3 = b3.data_ptr()
    p.res = res.data_ptr(); p.res_ld = res.stride(-2)
    p.ln2_gamma = ln2[0].data_ptr(); p.ln2_beta = ln2[1].data_ptr()
    p.w4 = w4.data_ptr(); p.b4 = b4.data_ptr(); p.w5 = w5.data_ptr(); p.b5 = b5.data_ptr()
    p.x = x.data_ptr(); p.x_ld = x.stride(-2)
    if out_bf16 is not None:
        p.out_bf16 = out_bf16.data_ptr(); p.out_ld = out_bf16.stride(-2)
    if ln is not None:
        _req_cuda(ln[0], ln[1])
        p.ln_gamma = ln[0].data_ptr(); p.ln_beta = ln[1].data_ptr()
    p.ln_eps = eps
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.load().ff_naf_tail(C.byref(p), _stream()), "ff_naf_tail")
        e1.record()
        M = B * H * W
        flops = 2.0 * M * (64 * 64 + 128 * 64 + 64 * 64)
        byts = M * 64 * 2 + (64 * 64 * (B if w3_batch_rows else 1) + 128 * 64 + 64 * 64) * 2 + M * 64 * 8 + (M * 64 * 2 if out_bf16 is not None else 0)
        PROFILE.records.append((e0, e1, flops, flops, float(byts), ("naf_tail", B, H, W)))
        return
    L.check(L.load().ff_naf_tail(C.byref(p), _stream()), "ff_naf_tail")


def window_attention(qkv, B, H, W, out, *, bias_table, wh, ww, kh=None, kw=None, kpad=(0, 0), shift=(0, 0), heads=6,
                     head_off=0, bias_head_off=0, rel_sign=1, rel_off=None, rel_stride=None, q_off=0, k_off=192,
                     v_off=384, out_off=0, padded=None):
    _req_cuda(qkv, out, bias_table)
    kh = wh if kh is None else kh
    kw = ww if kw is None else kw
    p = L.FFWinAttn()
    p.qkv = qkv.data_ptr(); p.ld = qkv.stride(-2)
    p.q_off, p.k_off, p.v_off = q_off, k_off, v_off
    p.B, p.H, p.W = B, H, W
    p.wh, p.ww, p.kh, p.kw = wh, ww, kh, kw
    p.kpad_y, p.kpad_x = kpad
    p.shift_y, p.shift_x = shift
    p.heads, p.head_off = heads, head_off
    p.bias_table = bias_table.data_ptr(); p.T = bias_table.shape[1]; p.bias_heads = bias_table.shape[0]   # [heads][T]
    p.bias_head_off = bias_head_off
    p.rel_sign = rel_sign
    p.rel_off_y, p.rel_off_x = rel_off if rel_off is not None else (wh - 1, ww - 1)
    p.rel_stride = rel_stride if rel_stride is not None else (2 * ww - 1)
    p.out = out.data_ptr(); p.out_ld = out.stride(-2); p.out_off = out_off
    if padded is not None:
        p.Hp, p.Wp = padded
    L.check(L.load().ff_window_attention(C.byref(p), _stream()), "ff_window_attention")


def layernorm(x, rows, C_, gamma, beta, eps, *, out_bf16=None, out_cols=None, out_f32=None, x_ld=None, x_off=0):
    _req_cuda(x, gamma, beta, out_bf16, out_f32)
    is_bf16 = 1 if x.dtype == _BF16 else 0
    esz = 2 if is_bf16 else 4
    out_cols = out_cols if out_cols is not None else (out_bf16.stride(-2) if out_bf16 is not None else C_)
    L.check(L.load().ff_layernorm(C.c_void_p(x.data_ptr() + x_off * esz), is_bf16, x_ld if x_ld is not None else x.stride(-2),
                                  C.c_longlong(rows), C_, _ptr(gamma), _ptr(beta), C.c_float(eps), _ptr(out_bf16),
                                  out_bf16.stride(-2) if out_bf16 is not None else 0, out_cols, _ptr(out_f32),
                                  out_f32.stride(-2) if out_f32 is not None else 0, _stream()), "ff_layernorm")


def gap(x, B, P, C_, out, scratch, *, x_off=0):
    _req_cuda(x, out, scratch)
    is_bf16 = 1 if x.dtype == _BF16 else 0
    esz = 2 if is_bf16 else 4
    L.check(L.load().ff_gap(C.c_void_p(x.data_ptr() + x_off * esz), is_bf16, x.stride(-2), B, P, C_, _ptr(out), out.stride(0), _ptr(scratch),
                            C.c_size_t(scratch.numel() * 4), _stream()), "ff_gap")


def gap_finalize(partial, B, nsplit, C_, inv, out):
    """out[b][c] = inv * sum_s partial[b][s][c] (second phase of the pool; partials from conv_gemm(col_sums=...))."""
    _req_cuda(partial, out)
    L.check(L.load().ff_gap_finalize(_ptr(partial), B, nsplit, C_, C.c_float(inv), _ptr(out), out.stride(0), _stream()), "ff_gap_finalize")


def pool_mlp_enabled():
    """Pool finalise + the per-sample MLP behind it as one launch (ff_gap_finalize_mlp); FFB200_POOL_MLP=0 restores the three launches."""
    import os
    return os.environ.get("FFB200_POOL_MLP", "1") != "0"


def gap_finalize_mlp(partial, B, nsplit, C_, inv, mean, counters, w1, b1, k1, act1, out, n_out, *, w2=None, b2=None, h1=0, act2=ACT_NONE, out_cols=None):
    """mean = pooled partials; out = act2(w2 . act1(w1 . mean + b1) + b2) (h1 > 0) or act1(w1 . mean + b1) (h1 == 0), see ff_gap_finalize_mlp."""
    _req_cuda(partial, mean, counters, w1, b1, w2, b2, out)
    L.check(L.load().ff_gap_finalize_mlp(_ptr(partial), B, nsplit, C_, C.c_float(inv), _ptr(mean), mean.stride(0), _ptr(counters), _ptr(w1), _ptr(b1), k1, h1, act1,
                                         _ptr(w2), _ptr(b2), n_out, act2, _ptr(out), out.stride(0), out_cols if out_cols is not None else n_out, _stream()),
            "ff_gap_finalize_mlp")


def vec_linear(x, R, K, W, bias, N, act, y, y_cols=None):
    _req_cuda(x, W, bias, y)
    L.check(L.load().ff_vec_linear(_ptr(x), x.stride(0), R, K, _ptr(W), _ptr(bias), N, act, _ptr(y), y.stride(0),
                                   y_cols if y_cols is not None else N, _stream()), "ff_vec_linear")


def dwconv(x, B, H, W, C_, kh, kw, w, bias, out, *, act=ACT_NONE, mode=0, mul=None, x_off=0, mul_off=0, x_ld=None):
    _req_cuda(x, w, bias, out, mul)
    L.check(L.load().ff_dwconv(C.c_void_p(x.data_ptr() + 2 * x_off), x_ld if x_ld is not None else x.stride(-2), B, H, W, C_, kh, kw,
                               _ptr(w), _ptr(bias), act, mode,
                               C.c_void_p(mul.data_ptr() + 2 * mul_off) if mul is not None else None,
                               mul.stride(-2) if mul is not None else 0, _ptr(out), out.stride(-2), _stream()), "ff_dwconv")


def dwconv_pool_rows(H, W, cout, mode=0):
    """Partial rows per sample of dwconv_pool, 0 when the shape does not tile for the fused pool."""
    return int(L.load().ff_dwconv_pool_rows(H, W, cout, mode))


def dwconv_pool(x, B, H, W, C_, w, bias, out, col_sums, *, act=ACT_NONE, mode=0, mul=None, x_off=0, x_ld=None):
    """3x3 depthwise conv + per-tile column sums of its output (finish with gap_finalize)."""
    _req_cuda(x, w, bias, out, mul, col_sums)
    L.check(L.load().ff_dwconv_pool(C.c_void_p(x.data_ptr() + 2 * x_off), x_ld if x_ld is not None else x.stride(-2), B, H, W, C_,
                                    _ptr(w), _ptr(bias), act, mode, _ptr(mul), mul.stride(-2) if mul is not None else 0,
                                    _ptr(out), out.stride(-2), _ptr(col_sums), _stream()), "ff_dwconv_pool")


def scale_channels(x, B, pixels_per_sample, C_, s):
    _req_cuda(x, s)
    L.check(L.load().ff_scale_channels(_ptr(x), x.stride(-2), B, C.c_longlong(pixels_per_sample), C_, _ptr(s), s.stride(0), _stream()),
            "ff_scale_channels")


def build_concat_diag_weights(w, s, alpha, out):
    """out[b] = [w | diag(alpha * s[b])] (bf16 [B][n_pad][k1 + n_pad]) -- the per-sample weights of a K-concatenated layer (conv_gemm x2=)."""
    _req_cuda(w, s, out)
    n_pad, k1 = w.shape
    B = out.shape[0]
    L.check(L.load().ff_build_concat_diag_weights(_ptr(w), n_pad, k1, _ptr(s), s.stride(0), C.c_float(alpha), B, _ptr(out), _stream()), "ff_build_concat_diag_weights")


def scale_weight_cols(w_f32, s, out):
    """out[b] = bf16(w * s[b][None, :]) -- see ff_scale_weight_cols."""
    _req_cuda(w_f32, s, out)
    N, K = w_f32.shape
    B, n_pad, k_pad = out.shape
    L.check(L.load().ff_scale_weight_cols(_ptr(w_f32), N, K, _ptr(s), s.stride(0), B, _ptr(out), n_pad, k_pad, _stream()), "ff_scale_weight_cols")


def pack_taps(x, B, H, W, Cin, k, terms, out):
    """fp32 NHWC rows -> split-bf16 (hi, lo, hi) operand rows, optionally with the 3x3 neighbourhood gathered (ff_pack_taps)."""
    _req_cuda(x, out)
    L.check(L.load().ff_pack_taps(_ptr(x), x.stride(-2), B, H, W, Cin, k, terms, _ptr(out), out.stride(-2), _stream()), "ff_pack_taps")


def conv_direct(x, B, H, W, Cin, k, w, bias, *, n_store, act=ACT_NONE, mul_f32=None, out_bf16=None, out_f32=None,
                x_ld=None, x_off=0, out_f32_off=0, out_bf16_off=0):
    _req_cuda(x, w, bias, mul_f32, out_bf16, out_f32)
    is_bf16 = 1 if x.dtype == _BF16 else 0
    esz = 2 if is_bf16 else 4
    L.check(L.load().ff_conv_direct(C.c_void_p(x.data_ptr() + x_off * esz), is_bf16, x_ld if x_ld is not None else x.stride(-2), B, H, W, Cin, k,
                                    _ptr(w), _ptr(bias), w.shape[0], n_store, act, _ptr(mul_f32),
                                    mul_f32.stride(-2) if mul_f32 is not None else 0,
                                    C.c_void_p(out_bf16.data_ptr() + 2 * out_bf16_off) if out_bf16 is not None else None,
                                    out_bf16.stride(-2) if out_bf16 is not None else 0,
                                    C.c_void_p(out_f32.data_ptr() + 4 * out_f32_off) if out_f32 is not None else None,
                                    out_f32.stride(-2) if out_f32 is not None else 0, _stream()), "ff_conv_direct")


def nchw_to_nhwc(x, out, sub=None):
    _req_cuda(x, out, sub)
    B, C_, H, W = x.shape
    L.check(L.load().ff_nchw_to_nhwc(_ptr(x), B, C_, H, W, _ptr(sub), _ptr(out), out.stride(-2), _stream()), "ff_nchw_to_nhwc")


def nchw_to_nhwc_pad(x, out, Hp, Wp, sub=None, reflect=True):
    """NCHW image -> NHWC rows of the right / bottom padded Hp x Wp image (reflect = pad_to_window_size, else zeros)."""
    _req_cuda(x, out, sub)
    B, C_, H, W = x.shape
    L.check(L.load().ff_nchw_to_nhwc_pad(_ptr(x), B, C_, H, W, _ptr(sub), _ptr(out), out.stride(-2), Hp, Wp, 1 if reflect else 0, _stream()), "ff_nchw_to_nhwc_pad")


def nhwc_to_nchw(x, coff, C_, out):
    _req_cuda(x, out)
    B, _, H, W = out.shape
    L.check(L.load().ff_nhwc_to_nchw(_ptr(x), x.stride(-2), coff, B, C_, H, W, _ptr(out), _stream()), "ff_nhwc_to_nchw")


def psnr_y(a, b, crop=4):
    """PSNR on BT.601 Y, border crop (reference src/utils/metrics.py:76-126); a, b: fp32 NCHW [B,3,H,W] on the GPU -> fp32 [B]."""
    _req_cuda(a, b)
    B, _, H, W = a.shape
    out = torch.empty(B, dtype=torch.float32, device=a.device)
    scratch = torch.empty(B * 64, dtype=torch.float64, device=a.device)
    L.check(L.load().ff_psnr_y(_ptr(a.contiguous()), _ptr(b.contiguous()), B, H, W, crop, _ptr(out), _ptr(scratch), C.c_size_t(scratch.numel() * 8), _stream()),
            "ff_psnr_y")
    return out


def ssim_y(a, b, crop=4):
    """SSIM on BT.601 Y, border crop (reference src/utils/metrics.py:189-246 -> :129-186, 11x11 Gaussian window, zero padding);
    a, b: fp32 NCHW [B,3,H,W] on the GPU -> fp32 [B]."""
    _req_cuda(a, b)
    B, _, H, W = a.shape
    lib = L.load()
    nbytes = int(lib.ff_ssim_y_scratch_bytes(B, H, W, crop))
    if nbytes == 0:
        raise ValueError(f"ssim_y: image {H}x{W} does not survive a crop of {crop}")
    out = torch.empty(B, dtype=torch.float32, device=a.device)
    scratch = torch.empty(nbytes // 8, dtype=torch.float64, device=a.device)
    L.check(lib.ff_ssim_y(_ptr(a.contiguous()), _ptr(b.contiguous()), B, H, W, crop, _ptr(out), _ptr(scratch), C.c_size_t(nbytes), _stream()),
            "ff_ssim_y")
    return out


def eval_psnr_ssim_u8(a, b, border=4):
    """(psnr, ssim) of the reference's evaluation harness (utils/utils_image.py:287-312 cal_psnr_ssim: OpenCV's 8-bit luma,
    scikit-image's 7x7 uniform-window SSIM) for two uint8 HWC RGB images on the GPU -> float64 [2] on the GPU."""
    _req_cuda(a, b)
    if a.dtype != torch.uint8 or b.dtype != torch.uint8 or a.dim() != 3 or a.shape[2] != 3 or a.shape != b.shape:
        raise ValueError(f"eval_psnr_ssim_u8: two uint8 [H, W, 3] images of the same size are needed, got {tuple(a.shape)} {a.dtype} / {tuple(b.shape)} {b.dtype}")
    H, W, _ = a.shape
    lib = L.load()
    nbytes = int(lib.ff_eval_scratch_bytes(H, W, border))
    if nbytes == 0:
        raise ValueError(f"eval_psnr_ssim_u8: image {H}x{W} does not hold a 7x7 window after a crop of {border}")
    out = torch.empty(2, dtype=torch.float64, device=a.device)
    scratch = torch.empty(nbytes // 8, dtype=torch.float64, device=a.device)
    L.check(lib.ff_eval_psnr_ssim_u8(_ptr(a.contiguous()), _ptr(b.contiguous()), H, W, border, _ptr(out), _ptr(scratch), C.c_size_t(nbytes), _stream()),
            "ff_eval_psnr_ssim_u8")
    return out
