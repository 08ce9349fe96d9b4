// Fusion-head kernels that are not plain convolutions (reference src/models/enhanced_fusion.py and friends).
// All NHWC; image-space tensors are fp32 with small channel pitches, feature maps bf16.
#include "ff_common.cuh"
#include "../../include/ffb200.h"

extern long long g_ff_launches;

namespace {

struct Bilin { int i0, i1; float l; };
// PyTorch area_pixel_compute_source_index (align_corners=False, non-cubic): src = ratio*(dst+0.5)-0.5, clamped at 0
__device__ __forceinline__ Bilin bilin(int o, float ratio, int in_size) {
  float s = ratio * (o + 0.5f) - 0.5f;
  s = fmaxf(s, 0.f);
  Bilin r;
  r.i0 = min((int)s, in_size - 1);
  r.i1 = min(r.i0 + 1, in_size - 1);
  r.l = s - r.i0;
  return r;
}

// ------------------------------------------------------------------------------------------------
// Cross-band attention front end (large_kernel_attention.py:207-229): per pixel and band,
//   proj = band_proj(band) (1x1, 3 -> 64);  stacked (bf16, the residual) and LN(proj) (bf16, MHA input).
// One warp per (pixel, band) token; lane handles channels lane and lane+32.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cb_embed_ln_kernel(const float* __restrict__ bands, long long tokens,
                                                         const float* __restrict__ pw, const float* __restrict__ pb,
                                                         const float* __restrict__ g, const float* __restrict__ bt,
                                                         bf16* __restrict__ stacked, bf16* __restrict__ normed) {
  const long long tok = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tok >= tokens) return;
  const int lane = threadIdx.x & 31;
  const float b0 = bands[tok * 3], b1 = bands[tok * 3 + 1], b2 = bands[tok * 3 + 2];
  float v[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = lane + 32 * i;
    v[i] = pw[c * 3] * b0 + pw[c * 3 + 1] * b1 + pw[c * 3 + 2] * b2 + pb[c];
  }
  const float mean = warp_sum(v[0] + v[1]) * (1.f / 64.f);
  const float d0 = v[0] - mean, d1 = v[1] - mean;
  const float rstd = rsqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.f / 64.f) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = lane + 32 * i;
    stacked[tok * 64 + c] = __float2bfloat16_rn(v[i]);
    normed[tok * 64 + c] = __float2bfloat16_rn((v[i] - mean) * rstd * g[c] + bt[c]);
  }
}

// The same front end with eight lanes per token: a lane owns eight channels (its 24 projection weights, bias, gamma, beta live in
// registers across a grid-stride loop), the LayerNorm statistics take three shuffles, and a token's two 128-byte rows leave as
// 16-byte stores (a warp writes 512 contiguous bytes per store instruction).
__global__ void __launch_bounds__(256) cb_embed_ln8_kernel(const float* __restrict__ bands, long long tokens,
                                                          const float* __restrict__ pw, const float* __restrict__ pb,
                                                          const float* __restrict__ g, const float* __restrict__ bt,
                                                          bf16* __restrict__ stacked, bf16* __restrict__ normed) {
  const int oct = threadIdx.x & 7;
  float w[8][3], bias[8], ga[8], be[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = oct * 8 + i;
    w[i][0] = pw[c * 3]; w[i][1] = pw[c * 3 + 1]; w[i][2] = pw[c * 3 + 2];
    bias[i] = pb[c]; ga[i] = g[c]; be[i] = bt[c];
  }
  const long long stride = ((long long)gridDim.x * 256) >> 3;
  const long long tok_end = (tokens + 3) & ~3LL;      // whole warps stay in the loop (shuffles), stores are masked
  for (long long tok = ((long long)blockIdx.x * 256 + threadIdx.x) >> 3; tok < tok_end; tok += stride) {
    const bool live = tok < tokens;
    const long long t = live ? tok : tokens - 1;
    const float b0 = __ldg(bands + t * 3), b1 = __ldg(bands + t * 3 + 1), b2 = __ldg(bands + t * 3 + 2);
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = w[i][0] * b0 + w[i][1] * b1 + w[i][2] * b2 + bias[i]; s += v[i]; }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / 64.f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q += d * d; }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / 64.f) + 1e-5f);
    if (live) {
      float n[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) n[i] = (v[i] - mean) * rstd * ga[i] + be[i];
      auto pk = [](float lo, float hi) { __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); };
      *reinterpret_cast<uint4*>(stacked + tok * 64 + oct * 8) = make_uint4(pk(v[0], v[1]), pk(v[2], v[3]), pk(v[4], v[5]), pk(v[6], v[7]));
      *reinterpret_cast<uint4*>(normed + tok * 64 + oct * 8) = make_uint4(pk(n[0], n[1]), pk(n[2], n[3]), pk(n[4], n[5]), pk(n[6], n[7]));
    }
  }
}

// 9-token multi-head attention per pixel (nn.MultiheadAttention core, 4 heads x 16; q pre-scaled by 1/4 in the packed in_proj).
// qkv: bf16 [tokens][192] (q | k | v); one thread per (token, head).
// DIM = embedding width (64: cross-band attention, 4 heads; 128: the collaborative branch's cross-expert attention, 8 heads).
template <int DIM>
__global__ void __launch_bounds__(256) cb_attn_kernel(const bf16* __restrict__ qkv, long long tokens, int nb, bf16* __restrict__ out) {
  constexpr int NH = DIM / 16;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= tokens * NH) return;
  const int h = (int)(idx % NH);
  const long long tok = idx / NH;
  const long long pix = tok / nb;
  float q[16];
  {
    const uint4* p = reinterpret_cast<const uint4*>(qkv + tok * (3 * DIM) + h * 16);
    const uint4 a = p[0], b = p[1];
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) { q[2 * i] = __uint_as_float(w[i] << 16); q[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  float s[9];
  float m = -1e30f;
  for (int j = 0; j < nb; ++j) {
    const uint4* p = reinterpret_cast<const uint4*>(qkv + (pix * nb + j) * (3 * DIM) + DIM + h * 16);
    const uint4 a = p[0], b = p[1];
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) d += q[2 * i] * __uint_as_float(w[i] << 16) + q[2 * i + 1] * __uint_as_float(w[i] & 0xffff0000u);
    s[j] = d;
    m = fmaxf(m, d);
  }
  float l = 0.f;
  for (int j = 0; j < nb; ++j) { s[j] = __expf(s[j] - m); l += s[j]; }
  const float inv = 1.f / l;
  float o[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = 0.f;
  for (int j = 0; j < nb; ++j) {
    const uint4* p = reinterpret_cast<const uint4*>(qkv + (pix * nb + j) * (3 * DIM) + 2 * DIM + h * 16);
    const uint4 a = p[0], b = p[1];
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const float pj = s[j] * inv;
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[2 * i] += pj * __uint_as_float(w[i] << 16); o[2 * i + 1] += pj * __uint_as_float(w[i] & 0xffff0000u); }
  }
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { __nv_bfloat162 hh = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]); w[i] = *reinterpret_cast<uint32_t*>(&hh); }
  uint4* dst = reinterpret_cast<uint4*>(out + tok * DIM + h * 16);
  dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
  dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// y = x * a[c] + b[c]  (eval BatchNorm as a per-channel affine), bf16 rows [rows][C], C a power-of-two multiple of 8
__global__ void __launch_bounds__(256) affine_rows_kernel(const bf16* __restrict__ x, long long rows, int C, const float* __restrict__ a,
                                                         const float* __restrict__ b, bf16* __restrict__ y) {
  const int groups = C >> 3;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= rows * groups) return;
  const int g = (int)(idx % groups);
  const uint4 q = reinterpret_cast<const uint4*>(x)[idx];
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = g * 8 + 2 * i;
    __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(w[i] << 16) * a[c] + b[c], __uint_as_float(w[i] & 0xffff0000u) * a[c + 1] + b[c + 1]);
    o[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  reinterpret_cast<uint4*>(y)[idx] = make_uint4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------------------------------------
// Adaptive band fusion 9 -> 3 + frequency guidance (multi_domain_frequency.py:478-526, enhanced_fusion.py:533-542).
// bands: fp32 [P][27] (after cross-band attention); att: fp32 [P][att_ld] = sigmoid(conv3x3(band_i)) per band.
// weights blob (fp32): imp[9] | Wt1[64][27] bt1[64] Wt2[9][64] bt2[9] | Wg1[64][27] bg1[64] Wg2[9][64] bg2[9] | Wr[9][9] br[9]
// Outputs: band_features fp32 [P][9] (3 guidance bands x 3 ch), guidance fp32 [P][4] = [high, mid, low]/sum (HAT, DAT, NAFNet).
// ------------------------------------------------------------------------------------------------
constexpr int BF_BLOB = 9 + 2 * (64 * 27 + 64 + 9 * 64 + 9) + 81 + 9;
__global__ void __launch_bounds__(128) band_fuse_kernel(const float* __restrict__ bands, const float* __restrict__ att, int att_ld,
                                                       long long P, const float* __restrict__ blob, float* __restrict__ feats,
                                                       float* __restrict__ guidance) {
  __shared__ float sw[BF_BLOB];
  for (int i = threadIdx.x; i < BF_BLOB; i += 128) sw[i] = blob[i];
  __syncthreads();
  const long long p = (long long)blockIdx.x * 128 + threadIdx.x;
  if (p >= P) return;
  const float* imp = sw;
  const float* Wt1 = sw + 9; const float* bt1 = Wt1 + 64 * 27; const float* Wt2 = bt1 + 64; const float* bt2 = Wt2 + 9 * 64;
  const float* Wg1 = bt2 + 9; const float* bg1 = Wg1 + 64 * 27; const float* Wg2 = bg1 + 64; const float* bg2 = Wg2 + 9 * 64;
  const float* Wr = bg2 + 9; const float* br = Wr + 81;
  float raw[27], x[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) raw[i] = bands[p * 27 + i];
#pragma unroll
  for (int b = 0; b < 9; ++b) {
    const float a = att[p * att_ld + b] * imp[b];
    x[3 * b] = raw[3 * b] * a; x[3 * b + 1] = raw[3 * b + 1] * a; x[3 * b + 2] = raw[3 * b + 2] * a;
  }
  float t[9], g[9];
#pragma unroll
  for (int o = 0; o < 9; ++o) { t[o] = bt2[o]; g[o] = bg2[o]; }
  for (int h = 0; h < 64; ++h) {
    float a = bt1[h], c = bg1[h];
#pragma unroll
    for (int i = 0; i < 27; ++i) { a += Wt1[h * 27 + i] * x[i]; c += Wg1[h * 27 + i] * x[i]; }
    a = gelu_erf(a); c = gelu_erf(c);
#pragma unroll
    for (int o = 0; o < 9; ++o) { t[o] += Wt2[o * 64 + h] * a; g[o] += Wg2[o * 64 + h] * c; }
  }
  float f[9];
#pragma unroll
  for (int o = 0; o < 9; ++o) {
    float r = br[o];
#pragma unroll
    for (int i = 0; i < 9; ++i) r += Wr[o * 9 + i] * raw[i];
    f[o] = t[o] * sigmoidf_(g[o]) + 0.3f * r;
    feats[p * 9 + o] = f[o];
  }
  const float lo = (fabsf(f[0]) + fabsf(f[1]) + fabsf(f[2])) * (1.f / 3.f);
  const float mi = (fabsf(f[3]) + fabsf(f[4]) + fabsf(f[5])) * (1.f / 3.f);
  const float hi = (fabsf(f[6]) + fabsf(f[7]) + fabsf(f[8])) * (1.f / 3.f);
  const float s = lo + mi + hi + 1e-8f;
  guidance[p * 4] = hi / s; guidance[p * 4 + 1] = mi / s; guidance[p * 4 + 2] = lo / s; guidance[p * 4 + 3] = 0.f;
}

// ------------------------------------------------------------------------------------------------
// Generic fp32 NHWC bilinear resize (align_corners=False): in [B][Hi][Wi][ld_in] (C channels) -> out [B][Ho][Wo][ld_out]
// (+ optional accumulate:  out = (acc ? out : 0) + resized + bias[c]).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bilinear_f32_kernel(const float* __restrict__ in, int B, int Hi, int Wi, int ld_in, int C,
                                                          float* __restrict__ out, int Ho, int Wo, int ld_out, int accumulate,
                                                          const float* __restrict__ bias, float ratio_y, float ratio_x) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const int cg = (C + 3) / 4;
  if (idx >= (long long)B * Ho * Wo * cg) return;
  const int g = (int)(idx % cg);
  const long long pix = idx / cg;
  const int xo = (int)(pix % Wo), yo = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
  const Bilin by = bilin(yo, ratio_y, Hi), bx = bilin(xo, ratio_x, Wi);
  const float* base = in + (long long)b * Hi * Wi * ld_in;
  for (int c = g * 4; c < min(C, g * 4 + 4); ++c) {
    const float v00 = base[((long long)by.i0 * Wi + bx.i0) * ld_in + c], v01 = base[((long long)by.i0 * Wi + bx.i1) * ld_in + c];
    const float v10 = base[((long long)by.i1 * Wi + bx.i0) * ld_in + c], v11 = base[((long long)by.i1 * Wi + bx.i1) * ld_in + c];
    float v = (1.f - by.l) * ((1.f - bx.l) * v00 + bx.l * v01) + by.l * ((1.f - bx.l) * v10 + bx.l * v11);
    if (bias) v += bias[c];
    float* o = out + pix * ld_out + c;
    *o = accumulate ? (*o + v) : v;
  }
}

// bf16 NHWC bilinear x2 up-sampling of C channels (C % 8 == 0) into a wider output row (pitch ld_out, channel offset 0)
__global__ void __launch_bounds__(256) bilinear_up2_bf16_kernel(const bf16* __restrict__ in, int B, int Hi, int Wi, int ld_in, int C,
                                                               bf16* __restrict__ out, int ld_out) {
  const int Ho = 2 * Hi, Wo = 2 * Wi, groups = C >> 3;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * Ho * Wo * groups) return;
  const int g = (int)(idx % groups);
  const long long pix = idx / groups;
  const int xo = (int)(pix % Wo), yo = (int)((pix / Wo) % Ho), b = (int)(pix / ((long long)Wo * Ho));
  const Bilin by = bilin(yo, 0.5f, Hi), bx = bilin(xo, 0.5f, Wi);
  const bf16* base = in + (long long)b * Hi * Wi * ld_in + g * 8;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const int ys[2] = {by.i0, by.i1}, xs[2] = {bx.i0, bx.i1};
  const float wy[2] = {1.f - by.l, by.l}, wx[2] = {1.f - bx.l, bx.l};
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint4 q = *reinterpret_cast<const uint4*>(base + ((long long)ys[a] * Wi + xs[c]) * ld_in);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
      const float ww = wy[a] * wx[c];
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[2 * i] += ww * __uint_as_float(w[i] << 16); acc[2 * i + 1] += ww * __uint_as_float(w[i] & 0xffff0000u); }
    }
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]); o[i] = *reinterpret_cast<uint32_t*>(&h); }
  *reinterpret_cast<uint4*>(out + pix * ld_out + g * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}

// DynamicExpertSelector tail (fusion_network.py:221-234): g = sigmoid(10 (g - (0.7 - 0.4 d))); mask = g >= 0.99 max g; g = max(g, 0.9 mask)
// gd: fp32 [P][4] = (gate0, gate1, gate2, difficulty) with the raw sigmoid gates on input; updated in place.
__global__ void __launch_bounds__(256) selector_tail_kernel(float* __restrict__ gd, long long P) {
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= P) return;
  float4 v = reinterpret_cast<float4*>(gd)[p];
  const float thr = 0.7f - 0.4f * v.w;
  float g0 = 1.f / (1.f + expf(-10.f * (v.x - thr))), g1 = 1.f / (1.f + expf(-10.f * (v.y - thr))), g2 = 1.f / (1.f + expf(-10.f * (v.z - thr)));
  const float mx = fmaxf(g0, fmaxf(g1, g2)) * 0.99f;
  g0 = fmaxf(g0, (g0 >= mx) ? 0.9f : 0.f);
  g1 = fmaxf(g1, (g1 >= mx) ? 0.9f : 0.f);
  g2 = fmaxf(g2, (g2 >= mx) ? 0.9f : 0.f);
  reinterpret_cast<float4*>(gd)[p] = make_float4(g0, g1, g2, v.w);
}

// ------------------------------------------------------------------------------------------------
// Hierarchical-fusion inputs (hierarchical_fusion.py:140-185): bilinear resize of the 9 stacked expert channels to
// 1/4, 1/2 and full resolution, written as bf16 into the (zero padded) conv input buffers at a channel offset.
//   factor 4: mean of the centre 2x2 of each 4x4 block;  factor 2: 2x2 mean;  factor 1: copy.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) experts_resize_kernel(const float* __restrict__ stack, int ld, int B, int H, int W, int factor,
                                                            bf16* __restrict__ out, int out_ld, int out_off) {
  const int Ho = H / factor, Wo = W / factor;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * Ho * Wo) return;
  const int xo = (int)(idx % Wo), yo = (int)((idx / Wo) % Ho), b = (int)(idx / ((long long)Wo * Ho));
  float v[9];
  if (factor == 1) {
    const float* p = stack + idx * ld;
#pragma unroll
    for (int c = 0; c < 9; ++c) v[c] = p[c];
  } else {
    const int y0 = factor == 2 ? 2 * yo : 4 * yo + 1, x0 = factor == 2 ? 2 * xo : 4 * xo + 1;
#pragma unroll
    for (int c = 0; c < 9; ++c) v[c] = 0.f;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        const float* p = stack + ((long long)(b * H + y0 + dy) * W + x0 + dx) * ld;
#pragma unroll
        for (int c = 0; c < 9; ++c) v[c] += 0.25f * p[c];
      }
  }
  bf16* o = out + idx * out_ld + out_off;
#pragma unroll
  for (int c = 0; c < 9; ++c) o[c] = __float2bfloat16_rn(v[c]);
}

// The same resize with 16-byte accesses: three float4 loads per source pixel (pitch 12) and two 16-byte stores per output pixel --
// channels out_off .. out_off + 8 and seven zeros behind them (the zero padding of the conv input rows, which nothing else writes).
__global__ void __launch_bounds__(256) experts_resize_vec_kernel(const float* __restrict__ stack, int ld, int B, int H, int W, int factor,
                                                                bf16* __restrict__ out, int out_ld, int out_off) {
  const int Ho = H / factor, Wo = W / factor;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * Ho * Wo) return;
  const int xo = (int)(idx % Wo), yo = (int)((idx / Wo) % Ho), b = (int)(idx / ((long long)Wo * Ho));
  float v[12];
  if (factor == 1) {
    const float4* p = reinterpret_cast<const float4*>(stack + idx * ld);
    const float4 a = __ldg(p), c = __ldg(p + 1), d = __ldg(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w; v[8] = d.x;
  } else {
    const int y0 = factor == 2 ? 2 * yo : 4 * yo + 1, x0 = factor == 2 ? 2 * xo : 4 * xo + 1;
#pragma unroll
    for (int c = 0; c < 9; ++c) v[c] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {      // (same summation order as the scalar kernel)
        const float4* p = reinterpret_cast<const float4*>(stack + ((long long)(b * H + y0 + dy) * W + x0 + dx) * ld);
        const float4 a = __ldg(p), c = __ldg(p + 1), d = __ldg(p + 2);
        v[0] += 0.25f * a.x; v[1] += 0.25f * a.y; v[2] += 0.25f * a.z; v[3] += 0.25f * a.w;
        v[4] += 0.25f * c.x; v[5] += 0.25f * c.y; v[6] += 0.25f * c.z; v[7] += 0.25f * c.w; v[8] += 0.25f * d.x;
      }
  }
  uint4* o = reinterpret_cast<uint4*>(out + idx * out_ld + out_off);
  auto pk = [](float lo, float hi) { __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); };
  o[0] = make_uint4(pk(v[0], v[1]), pk(v[2], v[3]), pk(v[4], v[5]), pk(v[6], v[7]));
  o[1] = make_uint4(pk(v[8], 0.f), 0u, 0u, 0u);
}

// SpatialGate (hierarchical_fusion.py:25-43): x *= sigmoid(w2 . gelu(W1 x + b1) + b2), per pixel, in place (bf16 rows).
// One warp per pixel; C in {32, 64}, hidden = C/4.
template <int C>
__global__ void __launch_bounds__(256) pixel_gate_kernel(bf16* __restrict__ x, int ld, long long P, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ w2, float b2) {
  // one thread per pixel: C values in registers, weights broadcast from shared memory
  constexpr int HID = C / 4;
  __shared__ float sW[HID * C], sB[HID], sW2[HID];
  for (int i = threadIdx.x; i < HID * C; i += 256) sW[i] = w1[i];
  if (threadIdx.x < HID) { sB[threadIdx.x] = b1[threadIdx.x]; sW2[threadIdx.x] = w2[threadIdx.x]; }
  __syncthreads();
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= P) return;
  float v[C];
  uint4* row = reinterpret_cast<uint4*>(x + p * ld);
#pragma unroll
  for (int i = 0; i < C / 8; ++i) {
    const uint4 q = row[i];
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[i * 8 + 2 * j] = __uint_as_float(w[j] << 16); v[i * 8 + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
  }
  float s = b2;
#pragma unroll 2
  for (int h = 0; h < HID; ++h) {
    float d = sB[h];
#pragma unroll
    for (int c = 0; c < C; ++c) d += v[c] * sW[h * C + c];
    s += gelu_erf(d) * sW2[h];
  }
  const float g = sigmoidf_(s);
#pragma unroll
  for (int i = 0; i < C / 8; ++i) {
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { __nv_bfloat162 hh = __floats2bfloat162_rn(v[i * 8 + 2 * j] * g, v[i * 8 + 2 * j + 1] * g); o[j] = *reinterpret_cast<uint32_t*>(&hh); }
    row[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// HR blend (enhanced_fusion.py:550-556 and 593-647) in one pass over the HR pixels:
//   fw    = sum_e expert_e * up4(guidance)_e;            fused = 0.7 * hier + 0.3 * fw
//   dyn   = sum_e expert_e * up4(gate)_e / (sum_e up4(gate)_e + 1e-8);   d = up4(difficulty)
//   fused = fused * (1 - 0.3 d) + dyn * 0.3 d                         -> fused [P][4]  (fused_before_refine)
//   base  = fused + residual_scale * up4(lr)                          -> base  [P][4]  (residual operand of refine_net's last conv)
// stack: fp32 [P][ld_s] experts (hat 0..2, dat 3..5, nafnet 6..8); hier fp32 [P][4]; guidance / gates fp32 LR [p][4]; lr NHWC fp32 [p][4]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) blend_kernel(const float* __restrict__ stack, int ld_s, const float* __restrict__ hier,
                                                   const float* __restrict__ guidance, const float* __restrict__ gates,
                                                   const float* __restrict__ lr, int B, int h, int w, float residual_scale,
                                                   float* __restrict__ fused, float* __restrict__ base) {
  const int H = 4 * h, W = 4 * w;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * H * W) return;
  const int xo = (int)(idx % W), yo = (int)((idx / W) % H), b = (int)(idx / ((long long)W * H));
  const Bilin by = bilin(yo, 0.25f, h), bx = bilin(xo, 0.25f, w);
  const long long o00 = ((long long)(b * h + by.i0) * w + bx.i0), o01 = ((long long)(b * h + by.i0) * w + bx.i1);
  const long long o10 = ((long long)(b * h + by.i1) * w + bx.i0), o11 = ((long long)(b * h + by.i1) * w + bx.i1);
  const float w00 = (1.f - by.l) * (1.f - bx.l), w01 = (1.f - by.l) * bx.l, w10 = by.l * (1.f - bx.l), w11 = by.l * bx.l;
  auto up = [&](const float* t) {
    const float4 a = reinterpret_cast<const float4*>(t)[o00], bq = reinterpret_cast<const float4*>(t)[o01];
    const float4 c = reinterpret_cast<const float4*>(t)[o10], d = reinterpret_cast<const float4*>(t)[o11];
    // same association as PyTorch's upsample_bilinear2d: (1-ly)*((1-lx)*v00 + lx*v01) + ly*((1-lx)*v10 + lx*v11)
    float4 r;
    r.x = (1.f - by.l) * ((1.f - bx.l) * a.x + bx.l * bq.x) + by.l * ((1.f - bx.l) * c.x + bx.l * d.x);
    r.y = (1.f - by.l) * ((1.f - bx.l) * a.y + bx.l * bq.y) + by.l * ((1.f - bx.l) * c.y + bx.l * d.y);
    r.z = (1.f - by.l) * ((1.f - bx.l) * a.z + bx.l * bq.z) + by.l * ((1.f - bx.l) * c.z + bx.l * d.z);
    r.w = (1.f - by.l) * ((1.f - bx.l) * a.w + bx.l * bq.w) + by.l * ((1.f - bx.l) * c.w + bx.l * d.w);
    return r;
  };
  (void)w00; (void)w01; (void)w10; (void)w11;
  const float4 gu = up(guidance), ga = up(gates), lv = up(lr);
  const float* e = stack + idx * ld_s;
  const float4 hr = reinterpret_cast<const float4*>(hier)[idx];
  const float hv[3] = {hr.x, hr.y, hr.z};
  const float lrv[3] = {lv.x, lv.y, lv.z};
  const float gsum = ga.x + ga.y + ga.z + 1e-8f;
  const float dd = 0.3f * ga.w;
  float f[3], bs[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float eh = e[c], ed = e[3 + c], en = e[6 + c];
    const float fw = (eh * gu.x + ed * gu.y) + en * gu.z;
    float v = hv[c] * 0.7f + fw * 0.3f;
    const float dyn = ((eh * ga.x + ed * ga.y) + en * ga.z) / gsum;
    v = v * (1.f - dd) + dyn * dd;
    f[c] = v;
    bs[c] = v + residual_scale * lrv[c];
  }
  reinterpret_cast<float4*>(fused)[idx] = make_float4(f[0], f[1], f[2], 0.f);
  reinterpret_cast<float4*>(base)[idx] = make_float4(bs[0], bs[1], bs[2], 0.f);
}

// ------------------------------------------------------------------------------------------------
// Laplacian pyramid pieces (edge_enhancement.py:182-220), fp32 NHWC with pitch ld (3 channels used).
//   gauss_down: down = avg_pool2(gaussian5x5(cur))  (zero padding)
//   lap_sub:    lap  = cur - bilinear_up2(down)
// ------------------------------------------------------------------------------------------------
// blur5x5 (zero pad) followed by the 2x2 mean is one separable 6-tap stencil per axis, k6[t] = k[t] + k[t-1] (t = 0..5,
// input offset 2*o - 2 + t): 36 taps per output instead of 4 x 25; each tap is one 16-byte load when the rows allow it.
__global__ void __launch_bounds__(256) gauss_down_kernel(const float* __restrict__ cur, int ld, int B, int H, int W,
                                                        const float* __restrict__ k1d, float* __restrict__ down, int ld_o) {
  const int Ho = H / 2, Wo = W / 2;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * Ho * Wo) return;
  const int xo = (int)(idx % Wo), yo = (int)((idx / Wo) % Ho), b = (int)(idx / ((long long)Wo * Ho));
  float k6[6];
#pragma unroll
  for (int t = 0; t < 6; ++t) k6[t] = (t < 5 ? k1d[t] : 0.f) + (t > 0 ? k1d[t - 1] : 0.f);
  const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(cur) & 15) == 0);
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int ty = 0; ty < 6; ++ty) {
    const int yy = 2 * yo - 2 + ty;
    if (yy < 0 || yy >= H) continue;
    float r[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int tx = 0; tx < 6; ++tx) {
      const int xx = 2 * xo - 2 + tx;
      if (xx < 0 || xx >= W) continue;
      const float* p = cur + ((long long)(b * H + yy) * W + xx) * ld;
      float v0, v1, v2;
      if (vec) { const float4 q = __ldg(reinterpret_cast<const float4*>(p)); v0 = q.x; v1 = q.y; v2 = q.z; }
      else { v0 = __ldg(p); v1 = __ldg(p + 1); v2 = __ldg(p + 2); }
      r[0] = fmaf(k6[tx], v0, r[0]); r[1] = fmaf(k6[tx], v1, r[1]); r[2] = fmaf(k6[tx], v2, r[2]);
    }
    acc[0] = fmaf(k6[ty], r[0], acc[0]); acc[1] = fmaf(k6[ty], r[1], acc[1]); acc[2] = fmaf(k6[ty], r[2], acc[2]);
  }
  float* o = down + idx * ld_o;
  o[0] = acc[0] * 0.25f; o[1] = acc[1] * 0.25f; o[2] = acc[2] * 0.25f;
  for (int c = 3; c < ld_o; ++c) o[c] = 0.f;
}
__global__ void __launch_bounds__(256) lap_sub_kernel(const float* __restrict__ cur, int ld, const float* __restrict__ down, int ld_d,
                                                     int B, int H, int W, float* __restrict__ lap, int ld_l) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * H * W) return;
  const int xo = (int)(idx % W), yo = (int)((idx / W) % H), b = (int)(idx / ((long long)W * H));
  const int Hd = H / 2, Wd = W / 2;
  const Bilin by = bilin(yo, 0.5f, Hd), bx = bilin(xo, 0.5f, Wd);
  const float* base = down + (long long)b * Hd * Wd * ld_d;
  for (int c = 0; c < 3; ++c) {
    const float v00 = base[((long long)by.i0 * Wd + bx.i0) * ld_d + c], v01 = base[((long long)by.i0 * Wd + bx.i1) * ld_d + c];
    const float v10 = base[((long long)by.i1 * Wd + bx.i0) * ld_d + c], v11 = base[((long long)by.i1 * Wd + bx.i1) * ld_d + c];
    const float u = (1.f - by.l) * ((1.f - bx.l) * v00 + bx.l * v01) + by.l * ((1.f - bx.l) * v10 + bx.l * v11);
    lap[idx * ld_l + c] = cur[idx * ld + c] - u;
  }
  for (int c = 3; c < ld_l; ++c) lap[idx * ld_l + c] = 0.f;
}

// Edge level merge (edge_enhancement.py:240-250): out[:, off:off+C] = weight * bilinear_up_f(feat * att) for one pyramid level.
// feat: bf16 [B][Hl][Wl][ld_f] (C channels), att: fp32 [B][Hl][Wl] (sigmoid map); out: bf16 [B][H][W][ld_o].
__global__ void __launch_bounds__(256) edge_merge_kernel(const bf16* __restrict__ feat, int ld_f, const float* __restrict__ att, int B,
                                                        int Hl, int Wl, int C, float weight, bf16* __restrict__ out, int H, int W,
                                                        int ld_o, int off) {
  const int groups = C >> 3;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * H * W * groups) return;
  const int g = (int)(idx % groups);
  const long long pix = idx / groups;
  const int xo = (int)(pix % W), yo = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
  const Bilin by = bilin(yo, (float)Hl / H, Hl), bx = bilin(xo, (float)Wl / W, Wl);
  const int ys[2] = {by.i0, by.i1}, xs[2] = {bx.i0, bx.i1};
  const float wy[2] = {1.f - by.l, by.l}, wx[2] = {1.f - bx.l, bx.l};
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const long long sp = (long long)(b * Hl + ys[a]) * Wl + xs[c];
      const uint4 q = *reinterpret_cast<const uint4*>(feat + sp * ld_f + g * 8);
      const uint32_t wq[4] = {q.x, q.y, q.z, q.w};
      const float ww = wy[a] * wx[c] * att[sp];
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[2 * i] += ww * __uint_as_float(wq[i] << 16); acc[2 * i + 1] += ww * __uint_as_float(wq[i] & 0xffff0000u); }
    }
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { __nv_bfloat162 hh = __floats2bfloat162_rn(acc[2 * i] * weight, acc[2 * i + 1] * weight); o[i] = *reinterpret_cast<uint32_t*>(&hh); }
  *reinterpret_cast<uint4*>(out + pix * ld_o + off + g * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}

// Final edge gating (edge_enhancement.py:256-260): out_nchw = clamp(sr + gate * strength * edge, 0, 1)
// se: fp32 [P][8] = (sr 0..2, edge 3..5); gate fp32 [P][gate_ld]
__global__ void __launch_bounds__(256) edge_final_kernel(const float* __restrict__ se, const float* __restrict__ gate, int gate_ld, int B, int H,
                                                        int W, float strength, float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long hw = (long long)H * W;
  if (idx >= B * hw) return;
  const int b = (int)(idx / hw);
  const long long p = idx - b * hw;
  const float g = gate[idx * gate_ld] * strength;
  for (int c = 0; c < 3; ++c) {
    const float v = se[idx * 8 + c] + g * se[idx * 8 + 3 + c];
    out[((long long)b * 3 + c) * hw + p] = fminf(fmaxf(v, 0.f), 1.f);
  }
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int ff_cb_embed_ln(const float* bands, long long tokens, const float* proj_w, const float* proj_b, const float* ln_w,
                              const float* ln_b, void* stacked, void* normed, void* stream) {
  FF_CHECK_ARG(bands && proj_w && proj_b && ln_w && ln_b && stacked && normed, "ff_cb_embed_ln: null buffer");
  if ((reinterpret_cast<uintptr_t>(stacked) & 15) == 0 && (reinterpret_cast<uintptr_t>(normed) & 15) == 0) {
    const long long want = ff_cdiv(tokens * 8, 256);
    const long long cap = (long long)ff_num_sms() * 16;
    cb_embed_ln8_kernel<<<(int)(want < cap ? want : cap), 256, 0, ST(stream)>>>(bands, tokens, proj_w, proj_b, ln_w, ln_b, reinterpret_cast<bf16*>(stacked), reinterpret_cast<bf16*>(normed));
  } else
    cb_embed_ln_kernel<<<ff_cdiv(tokens, 8), 256, 0, ST(stream)>>>(bands, tokens, proj_w, proj_b, ln_w, ln_b, reinterpret_cast<bf16*>(stacked), reinterpret_cast<bf16*>(normed));
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_cb_embed_ln"); return FF_OK;
}
extern "C" int ff_cb_attention(const void* qkv, long long tokens, int num_bands, void* out, void* stream) {
  FF_CHECK_ARG(qkv && out && num_bands > 0 && num_bands <= 9 && tokens % num_bands == 0, "ff_cb_attention: bad args");
  cb_attn_kernel<64><<<ff_cdiv(tokens * 4, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(qkv), tokens, num_bands, reinterpret_cast<bf16*>(out));
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_cb_attention"); return FF_OK;
}
extern "C" int ff_token_attention(const void* qkv, long long tokens, int group, int dim, void* out, void* stream) {
  FF_CHECK_ARG(qkv && out && group > 0 && group <= 9 && tokens % group == 0 && (dim == 64 || dim == 128), "ff_token_attention: bad args");
  if (dim == 64) cb_attn_kernel<64><<<ff_cdiv(tokens * 4, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(qkv), tokens, group, reinterpret_cast<bf16*>(out));
  else cb_attn_kernel<128><<<ff_cdiv(tokens * 8, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(qkv), tokens, group, reinterpret_cast<bf16*>(out));
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_token_attention"); return FF_OK;
}

// ---- collaborative branch helpers (large_kernel_attention.py:327-419) ----
namespace {
// NCHW fp32 features -> NHWC bf16 rows [B*H*W][ld], columns >= C zero (GEMM operand of the align convs, :341-356)
__global__ void __launch_bounds__(256) nchw_to_nhwc_bf16_kernel(const float* __restrict__ x, int B, int C, long long hw, bf16* __restrict__ out, int ld) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const int groups = ld >> 3;
  if (idx >= (long long)B * hw * groups) return;
  const int g = (int)(idx % groups);
  const long long pix = idx / groups;
  const int b = (int)(pix / hw);
  const long long p = pix - (long long)b * hw;
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = g * 8 + 2 * i;
    const float v0 = c < C ? x[((long long)b * C + c) * hw + p] : 0.f, v1 = c + 1 < C ? x[((long long)b * C + c + 1) * hw + p] : 0.f;
    w[i] = pack_bf16(v0, v1);
  }
  *reinterpret_cast<uint4*>(out + pix * ld + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
}
// partial[b][blk][c] = sum over the HR pixels of block blk of GELU(bilinear_up_s(g)[pixel][c]):  the AdaptiveAvgPool2d(1) of
// GELU(conv1x1(upsample(feat))) in a modulation head (:407-411) -- the 1x1 conv commutes with the bilinear up-sampling, so g is
// the conv output at LR resolution and the up-sampled tensor never exists.  lane = channel (C <= 32), warps stride over pixels.
__global__ void __launch_bounds__(256) up_gelu_pool_kernel(const float* __restrict__ g, int ld, int h, int w, int C, int s, int nblk, float* __restrict__ partial) {
  __shared__ float red[8][33];
  const int b = blockIdx.y, blk = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int H = h * s, W = w * s;
  const long long total = (long long)H * W, per = (total + nblk - 1) / nblk;
  const long long p0 = blk * per, p1 = min(total, p0 + per);
  const float ratio = 1.0f / s;
  const float* gb = g + (long long)b * h * w * ld;
  float acc = 0.f;
  if (lane < C)
    for (long long p = p0 + warp; p < p1; p += 8) {
      const int X = (int)(p % W), Y = (int)(p / W);
      const Bilin by = bilin(Y, ratio, h), bx = bilin(X, ratio, w);
      const float v00 = gb[((long long)by.i0 * w + bx.i0) * ld + lane], v01 = gb[((long long)by.i0 * w + bx.i1) * ld + lane];
      const float v10 = gb[((long long)by.i1 * w + bx.i0) * ld + lane], v11 = gb[((long long)by.i1 * w + bx.i1) * ld + lane];
      acc += gelu_erf((1.f - by.l) * ((1.f - bx.l) * v00 + bx.l * v01) + by.l * ((1.f - bx.l) * v10 + bx.l * v11));
    }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && lane < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][lane];
    partial[((long long)b * nblk + blk) * C + lane] = t;
  }
}
// x[p][c_off + c] = clamp(x * f[b][c], 0, 1): the soft modulation of an expert's SR output (:414-415) on the fp32 expert stack
__global__ void __launch_bounds__(256) scale_clamp_channels_kernel(float* __restrict__ x, int ld, long long per_b, int B, int c_off, int C, const float* __restrict__ m, int m_ld,
                                                                  float f0, float f1) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= per_b * B) return;
  const int b = (int)(idx / per_b);
  for (int c = 0; c < C; ++c) {
    float* q = x + idx * ld + c_off + c;
    *q = fminf(fmaxf(*q * (f0 + f1 * m[(long long)b * m_ld + c]), 0.f), 1.f);
  }
}
}  // namespace
extern "C" int ff_nchw_to_nhwc_bf16(const float* x, int B, int C, int H, int W, void* out, int ld, void* stream) {
  FF_CHECK_ARG(x && out && ld >= C && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "ff_nchw_to_nhwc_bf16: bad args");
  nchw_to_nhwc_bf16_kernel<<<ff_cdiv((long long)B * H * W * (ld / 8), 256), 256, 0, ST(stream)>>>(x, B, C, (long long)H * W, reinterpret_cast<bf16*>(out), ld);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_nchw_to_nhwc_bf16"); return FF_OK;
}
extern "C" int ff_up_gelu_pool(const float* g, int ld, int B, int h, int w, int C, int scale, int nblk, float* partial, void* stream) {
  FF_CHECK_ARG(g && partial && C > 0 && C <= 32 && ld >= C && scale >= 1 && nblk > 0 && B > 0 && B <= 65535, "ff_up_gelu_pool: bad args");
  up_gelu_pool_kernel<<<dim3(nblk, B), 256, 0, ST(stream)>>>(g, ld, h, w, C, scale, nblk, partial);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_up_gelu_pool"); return FF_OK;
}
extern "C" int ff_scale_clamp_channels(float* x, int ld, int B, long long pixels_per_sample, int c_off, int C, const float* m, int m_ld, float f0, float f1,
                                       void* stream) {
  FF_CHECK_ARG(x && m && C > 0 && c_off >= 0 && ld >= c_off + C && m_ld >= C, "ff_scale_clamp_channels: bad args");
  scale_clamp_channels_kernel<<<ff_cdiv(pixels_per_sample * B, 256), 256, 0, ST(stream)>>>(x, ld, pixels_per_sample, B, c_off, C, m, m_ld, f0, f1);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_scale_clamp_channels"); return FF_OK;
}
extern "C" int ff_affine_rows(const void* x, long long rows, int C, const float* a, const float* b, void* y, void* stream) {
  FF_CHECK_ARG(x && a && b && y && C % 8 == 0, "ff_affine_rows: bad args");
  affine_rows_kernel<<<ff_cdiv(rows * (C / 8), 256), 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(x), rows, C, a, b, reinterpret_cast<bf16*>(y));
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_affine_rows"); return FF_OK;
}
extern "C" int ff_band_fuse(const float* bands, const float* att, int att_ld, long long P, const float* blob, int blob_len, float* feats,
                            float* guidance, void* stream) {
  FF_CHECK_ARG(bands && att && blob && feats && guidance, "ff_band_fuse: null buffer");
  FF_CHECK_ARG(blob_len == BF_BLOB, "ff_band_fuse: weight blob has %d floats, expected %d", blob_len, BF_BLOB);
  band_fuse_kernel<<<ff_cdiv(P, 128), 128, 0, ST(stream)>>>(bands, att, att_ld, P, blob, feats, guidance);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_band_fuse"); return FF_OK;
}
extern "C" int ff_bilinear_f32(const float* in, int B, int Hi, int Wi, int ld_in, int C, float* out, int Ho, int Wo, int ld_out,
                               int accumulate, const float* bias, void* stream) {
  FF_CHECK_ARG(in && out && C <= ld_in && C <= ld_out, "ff_bilinear_f32: bad args");
  bilinear_f32_kernel<<<ff_cdiv((long long)B * Ho * Wo * ((C + 3) / 4), 256), 256, 0, ST(stream)>>>(in, B, Hi, Wi, ld_in, C, out, Ho, Wo, ld_out, accumulate, bias,
                                                                                                    (float)Hi / Ho, (float)Wi / Wo);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_bilinear_f32"); return FF_OK;
}
extern "C" int ff_bilinear_f32_scaled(const float* in, int B, int Hi, int Wi, int ld_in, int C, float* out, int Ho, int Wo, int ld_out, float ratio_y,
                                      float ratio_x, void* stream) {
  FF_CHECK_ARG(in && out && C <= ld_in && C <= ld_out && ratio_y > 0.f && ratio_x > 0.f, "ff_bilinear_f32_scaled: bad args");
  bilinear_f32_kernel<<<ff_cdiv((long long)B * Ho * Wo * ((C + 3) / 4), 256), 256, 0, ST(stream)>>>(in, B, Hi, Wi, ld_in, C, out, Ho, Wo, ld_out, 0, nullptr, ratio_y, ratio_x);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_bilinear_f32_scaled"); return FF_OK;
}
extern "C" int ff_bilinear_up2_bf16(const void* in, int B, int Hi, int Wi, int ld_in, int C, void* out, int ld_out, void* stream) {
  FF_CHECK_ARG(in && out && C % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0, "ff_bilinear_up2_bf16: bad args");
  bilinear_up2_bf16_kernel<<<ff_cdiv((long long)B * 4 * Hi * Wi * (C / 8), 256), 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(in), B, Hi, Wi, ld_in, C, reinterpret_cast<bf16*>(out), ld_out);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_bilinear_up2_bf16"); return FF_OK;
}
extern "C" int ff_selector_tail(float* gates_difficulty, long long P, void* stream) {
  FF_CHECK_ARG(gates_difficulty != nullptr, "ff_selector_tail: null buffer");
  selector_tail_kernel<<<ff_cdiv(P, 256), 256, 0, ST(stream)>>>(gates_difficulty, P);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_selector_tail"); return FF_OK;
}
extern "C" int ff_experts_resize(const float* stack, int ld, int B, int H, int W, int factor, void* out, int out_ld, int out_off, void* stream) {
  FF_CHECK_ARG(stack && out && (factor == 1 || factor == 2 || factor == 4) && H % factor == 0 && W % factor == 0 && ld >= 9, "ff_experts_resize: bad args");
  const bool vec = ld % 4 == 0 && ld >= 12 && out_off % 8 == 0 && out_ld % 8 == 0 && out_ld >= out_off + 16 && (reinterpret_cast<uintptr_t>(stack) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (vec) experts_resize_vec_kernel<<<ff_cdiv((long long)B * (H / factor) * (W / factor), 256), 256, 0, ST(stream)>>>(stack, ld, B, H, W, factor, reinterpret_cast<bf16*>(out), out_ld, out_off);
  else experts_resize_kernel<<<ff_cdiv((long long)B * (H / factor) * (W / factor), 256), 256, 0, ST(stream)>>>(stack, ld, B, H, W, factor, reinterpret_cast<bf16*>(out), out_ld, out_off);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_experts_resize"); return FF_OK;
}
extern "C" int ff_pixel_gate(void* x, int ld, long long P, int C, const float* w1, const float* b1, const float* w2, float b2, void* stream) {
  FF_CHECK_ARG(x && w1 && b1 && w2 && (C == 32 || C == 64), "ff_pixel_gate: C must be 32 or 64");
  FF_CHECK_ARG(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "ff_pixel_gate: rows must be 16-byte aligned");
  if (C == 64) pixel_gate_kernel<64><<<ff_cdiv(P, 256), 256, 0, ST(stream)>>>(reinterpret_cast<bf16*>(x), ld, P, w1, b1, w2, b2);
  else pixel_gate_kernel<32><<<ff_cdiv(P, 256), 256, 0, ST(stream)>>>(reinterpret_cast<bf16*>(x), ld, P, w1, b1, w2, b2);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_pixel_gate"); return FF_OK;
}
extern "C" int ff_blend(const float* stack, int ld_s, const float* hier, const float* guidance, const float* gates, const float* lr, int B, int h,
                        int w, float residual_scale, float* fused, float* base, void* stream) {
  FF_CHECK_ARG(stack && hier && guidance && gates && lr && fused && base && ld_s >= 9, "ff_blend: bad args");
  blend_kernel<<<ff_cdiv((long long)B * 16 * h * w, 256), 256, 0, ST(stream)>>>(stack, ld_s, hier, guidance, gates, lr, B, h, w, residual_scale, fused, base);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_blend"); return FF_OK;
}
extern "C" int ff_gauss_down(const float* cur, int ld, int B, int H, int W, const float* k1d, float* down, int ld_o, void* stream) {
  FF_CHECK_ARG(cur && k1d && down && H % 2 == 0 && W % 2 == 0 && ld >= 3 && ld_o >= 3, "ff_gauss_down: bad args");
  gauss_down_kernel<<<ff_cdiv((long long)B * (H / 2) * (W / 2), 256), 256, 0, ST(stream)>>>(cur, ld, B, H, W, k1d, down, ld_o);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_gauss_down"); return FF_OK;
}
extern "C" int ff_lap_sub(const float* cur, int ld, const float* down, int ld_d, int B, int H, int W, float* lap, int ld_l, void* stream) {
  FF_CHECK_ARG(cur && down && lap, "ff_lap_sub: null buffer");
  lap_sub_kernel<<<ff_cdiv((long long)B * H * W, 256), 256, 0, ST(stream)>>>(cur, ld, down, ld_d, B, H, W, lap, ld_l);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_lap_sub"); return FF_OK;
}
extern "C" int ff_edge_merge(const void* feat, int ld_f, const float* att, int B, int Hl, int Wl, int C, float weight, void* out, int H, int W,
                             int ld_o, int off, void* stream) {
  FF_CHECK_ARG(feat && att && out && C % 8 == 0 && off % 8 == 0, "ff_edge_merge: bad args");
  edge_merge_kernel<<<ff_cdiv((long long)B * H * W * (C / 8), 256), 256, 0, ST(stream)>>>(reinterpret_cast<const bf16*>(feat), ld_f, att, B, Hl, Wl, C, weight, reinterpret_cast<bf16*>(out), H, W, ld_o, off);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_edge_merge"); return FF_OK;
}
extern "C" int ff_edge_final(const float* se, const float* gate, int gate_ld, int B, int H, int W, float strength, float* out, void* stream) {
  FF_CHECK_ARG(se && gate && out, "ff_edge_final: null buffer");
  edge_final_kernel<<<ff_cdiv((long long)B * H * W, 256), 256, 0, ST(stream)>>>(se, gate, gate_ld, B, H, W, strength, out);
  ++g_ff_launches; FF_CHECK_LAUNCH("ff_edge_final"); return FF_OK;
}
