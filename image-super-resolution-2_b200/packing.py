"""Load-time weight repacking (host side, runs once per checkpoint).

Layouts produced here are what the kernels in csrc/ expect:
  * GEMM / conv weights: bf16 [n_pad][taps*cin_pad], K index = tap*cin_pad + c, tap = ky*kw + kx;
  * 180-channel transformer tensors live in 192-wide buffers, either "plain" (c -> c) or
    "head-padded" (c -> (c//30)*32 + c%30) so every 30-wide head starts on a 32-channel boundary;
  * eval-mode BatchNorm is folded into the adjacent convolution.
"""
import torch

BF16 = torch.bfloat16
F32 = torch.float32


def rup(x, m):
    return (x + m - 1) // m * m


def head_pad_index(c, head_dim=30, pad=32):
    """plain channel index (tensor of ints) -> head-padded index"""
    return (c // head_dim) * pad + c % head_dim


def pack_matrix(W, n_pad, k_pad, row_index=None, col_index=None, dtype=BF16, device="cuda"):
    """W [N, K] fp32 -> [n_pad, k_pad] with optional scatter of rows/cols to new positions."""
    N, K = W.shape
    out = torch.zeros(n_pad, k_pad, dtype=F32)
    r = torch.arange(N) if row_index is None else row_index
    c = torch.arange(K) if col_index is None else col_index
    out[r[:, None], c[None, :]] = W.to(F32)
    t = out.to(dtype).to(device).contiguous()
    t.ff_real = (N, K)      # un-padded (n, k): used only for FLOP accounting in bench.py
    return t


def pack_vector(v, n_pad, index=None, device="cuda", fill=0.0):
    out = torch.full((n_pad,), fill, dtype=F32)
    idx = torch.arange(v.numel()) if index is None else index
    out[idx] = v.reshape(-1).to(F32)
    return out.to(device).contiguous()


def pack_conv(W, n_pad, cin_pad, row_index=None, col_index=None, device="cuda", dtype=BF16):
    """Conv weight [Cout, Cin, kh, kw] -> [n_pad, kh*kw*cin_pad] (tap-major K)."""
    Cout, Cin, kh, kw = W.shape
    out = torch.zeros(n_pad, kh * kw, cin_pad, dtype=F32)
    r = torch.arange(Cout) if row_index is None else row_index
    c = torch.arange(Cin) if col_index is None else col_index
    Wt = W.to(F32).permute(0, 2, 3, 1).reshape(Cout, kh * kw, Cin)
    out[r[:, None], :, c[None, :]] = Wt.permute(0, 2, 1)
    t = out.reshape(n_pad, kh * kw * cin_pad).to(dtype).to(device).contiguous()
    t.ff_real = (Cout, kh * kw * Cin)
    return t


def _split_bf16(W):
    hi = W.to(F32).to(BF16).to(F32)
    lo = (W.to(F32) - hi).to(BF16).to(F32)
    return hi, lo


def pack_conv_split3(W, n_pad, device="cuda"):
    """fp32 conv weight [Cout, Cin, k, k] for the split-bf16 tensor-core path: input channels become [hi | lo | hi] blocks
    (ops.pack_taps with terms = 3), so the weight channels are [w_hi ; w_hi ; w_lo] -> [n_pad, k*k*3*Cin]."""
    hi, lo = _split_bf16(W)
    t = pack_conv(torch.cat([hi, hi, lo], dim=1), n_pad, 3 * W.shape[1], device=device)
    t.ff_real = (W.shape[0], W.shape[2] * W.shape[3] * W.shape[1])
    return t


def pack_conv_im2col2(W, n_pad, device="cuda"):
    """fp32 conv weight [Cout, Cin, 3, 3] (Cin * 18 <= 64) for ops.pack_taps(k=3, terms=2): K index (t*9 + tap)*Cin + c holds
    bf16(W[:, c, tap]) for both activation terms -> [n_pad, 64]."""
    Cout, Cin, kh, kw = W.shape
    hi, _ = _split_bf16(W)
    row = hi.permute(0, 2, 3, 1).reshape(Cout, kh * kw * Cin)
    out = torch.zeros(n_pad, 64, dtype=F32)
    out[:Cout, :2 * kh * kw * Cin] = torch.cat([row, row], dim=1)
    t = out.to(BF16).to(device).contiguous()
    t.ff_real = (Cout, kh * kw * Cin)
    return t


def pack_conv_direct(W, cout_pad, device="cuda"):
    """Conv weight [Cout, Cin, k, k] -> fp32 [cout_pad, k*k*Cin] for ff_conv_direct."""
    Cout, Cin, kh, kw = W.shape
    out = torch.zeros(cout_pad, kh * kw * Cin, dtype=F32)
    out[:Cout] = W.to(F32).permute(0, 2, 3, 1).reshape(Cout, -1)
    return out.to(device).contiguous()


def pack_dw(W, c_pad, index=None, device="cuda"):
    """Depthwise weight [C, 1, kh, kw] -> fp32 [kh*kw][c_pad] (tap-major)."""
    C_, _, kh, kw = W.shape
    out = torch.zeros(kh * kw, c_pad, dtype=F32)
    idx = torch.arange(C_) if index is None else index
    out[:, idx] = W.to(F32).reshape(C_, kh * kw).t()
    return out.to(device).contiguous()


def pixel_shuffle_rows(cout, r=2):
    """Row permutation that lets the conv epilogue do nn.PixelShuffle(r): original output channel
    n = c*r*r + i*r + j  ->  packed row (i*r+j)*(cout/r^2) + c."""
    n = torch.arange(cout)
    c, sub = n // (r * r), n % (r * r)
    return sub * (cout // (r * r)) + c


def fold_bn(weight, bias, bn_w, bn_b, bn_mean, bn_var, eps=1e-5):
    """conv -> BN(eval)  ==  conv with scaled weights/bias."""
    s = bn_w / torch.sqrt(bn_var + eps)
    w = weight * s.reshape(-1, *([1] * (weight.dim() - 1)))
    b = (bias if bias is not None else torch.zeros_like(bn_mean)) * s + (bn_b - bn_mean * s)
    return w, b
