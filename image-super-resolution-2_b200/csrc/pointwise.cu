// Memory-bound building blocks shared by the three experts and the fusion head (NHWC layouts):
// layer norm, global average pool, tiny per-sample linear layers (SE / SCA / AIM heads), depthwise
// convolutions with fused gates, direct small-channel convolutions (fp32), layout conversion.
#include "ff_common.cuh"
#include <type_traits>
#include "../../include/ffb200.h"

extern long long g_ff_launches;

namespace {

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case FF_ACT_GELU: return gelu_erf(v);
    case FF_ACT_RELU: return fmaxf(v, 0.f);
    case FF_ACT_LRELU: return v > 0.f ? v : 0.01f * v;
    case FF_ACT_SIGMOID: return sigmoidf_(v);
    case FF_ACT_CLAMP01: return fminf(fmaxf(v, 0.f), 1.f);
    default: return v;
  }
}

// GELU on the hardware tanh unit for bf16-stored activations (see csrc/conv_gemm.cu: gelu_tanh_hw, |err| << bf16 rounding)
__device__ __forceinline__ float gelu_tanh_hw(float x) {
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// ------------------------------------------------------------------------------------------
// LayerNorm over the channel axis of [rows][ld]; one warp per row, values kept in registers.
// Lanes own channel PAIRS (c = 2*lane + 64*i): 8-byte fp32 / 4-byte bf16 loads and 4-byte bf16x2 stores halve the
// load/store instruction count of this issue-bound kernel.  C and out_cols must be even.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ld_pair(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ld_pair(const bf16* p) {
  const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(p);
  return make_float2(__low2float(h), __high2float(h));
}

template <typename TIn, int NP>  // NP = pairs per lane = ceil(out_cols / 64)
__global__ void __launch_bounds__(256) layernorm_kernel(const TIn* __restrict__ x, int in_ld, long long rows, int C,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, bf16* __restrict__ out_bf16, int out_ld, int out_cols,
                                                       float* __restrict__ out_f32, int out_f32_ld) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const TIn* xr = x + row * in_ld;
  float2 v[NP];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int c = 2 * lane + 64 * i;
    v[i] = (c < C) ? ld_pair(xr + c) : make_float2(0.f, 0.f);
    s += v[i].x + v[i].y;
  }
  const float mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int c = 2 * lane + 64 * i;
    if (c < C) {
      const float d0 = v[i].x - mean, d1 = v[i].y - mean;
      q += d0 * d0 + d1 * d1;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int c = 2 * lane + 64 * i;
    if (c < out_cols) {
      float2 y = make_float2(0.f, 0.f);
      if (c < C) {
        const float2 g = __ldg(reinterpret_cast<const float2*>(gamma + c)), bb = __ldg(reinterpret_cast<const float2*>(beta + c));
        y.x = (v[i].x - mean) * rstd * g.x + bb.x;
        y.y = (v[i].y - mean) * rstd * g.y + bb.y;
      }
      if (out_bf16) *reinterpret_cast<__nv_bfloat162*>(out_bf16 + row * out_ld + c) = __floats2bfloat162_rn(y.x, y.y);
      if (out_f32) *reinterpret_cast<float2*>(out_f32 + row * out_f32_ld + c) = y;
    }
  }
}

__device__ __forceinline__ uint4 ln_pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Narrow rows (C = 64 or 128, fp32 in, bf16 out, no padding): LPR lanes per row, each lane owns 8 contiguous channels
// (two 16-byte loads, one 16-byte bf16 store), 32/LPR rows per warp, log2(LPR) shuffle steps.  NAFNet's high-resolution levels.
template <int LPR>
__global__ void __launch_bounds__(256) layernorm_narrow_kernel(const float* __restrict__ x, int in_ld, long long rows,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                              bf16* __restrict__ out, int out_ld) {
  constexpr int C = LPR * 8;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const long long row = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool ok = row < rows;
  const float* xr = x + (ok ? row : 0) * in_ld + sub * 8;
  const float4 a = *reinterpret_cast<const float4*>(xr), b = *reinterpret_cast<const float4*>(xr + 4);
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] -= mean; q += v[i] * v[i]; }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.f / C) + eps);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + sub * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + sub * 8) + 1);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + sub * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + sub * 8) + 1);
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  float o8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o8[i] = v[i] * rstd * gg[i] + bb[i];
  if (ok) *reinterpret_cast<uint4*>(out + row * out_ld + sub * 8) = ln_pack8(o8);
}

// The residual-stream LayerNorm of HAT / DAT (fp32 [rows][>=192] -> bf16 / fp32, out_cols = 192, C <= 192, C % 4 == 0):
// 16 lanes per row, two rows per warp, lanes own channel QUADS (c = 4*sub + 64*i): three 16-byte loads and three 8-byte bf16
// stores per lane -- twice the bytes in flight per warp and half the memory instructions of the pair kernel above.
__global__ void __launch_bounds__(256) layernorm_w192_kernel(const float* __restrict__ x, int in_ld, long long rows, int C,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                            bf16* __restrict__ out_bf16, int out_ld, float* __restrict__ out_f32, int out_f32_ld) {
  const int lane = threadIdx.x & 31, sub = lane & 15;
  const long long row = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * 2 + (lane >> 4);
  const bool ok = row < rows;
  const float* xr = x + (ok ? row : 0) * in_ld;
  float4 v[3];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = 4 * sub + 64 * i;
    v[i] = (c < C) ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (4 * sub + 64 * i < C) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  if (!ok) return;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = 4 * sub + 64 * i;
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c)), bb = __ldg(reinterpret_cast<const float4*>(beta + c));
      y.x = v[i].x * rstd * g.x + bb.x; y.y = v[i].y * rstd * g.y + bb.y; y.z = v[i].z * rstd * g.z + bb.z; y.w = v[i].w * rstd * g.w + bb.w;
    }
    if (out_bf16) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
      *reinterpret_cast<uint2*>(out_bf16 + row * out_ld + c) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * out_f32_ld + c) = y;
  }
}

// ------------------------------------------------------------------------------------------
// Global average pool over pixels: x [B][P][ld] -> partial sums [B][nsplit][C] -> mean [B][out_ld]
// (two phases, fixed summation order => deterministic).
// ------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256) gap_partial_kernel(const TIn* __restrict__ x, int ld, int P, int C, int nsplit,
                                                         float* __restrict__ partial) {
  __shared__ float red[8][33];
  const int b = blockIdx.z, split = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;
  const int per = (P + nsplit - 1) / nsplit;
  const int p0 = split * per, p1 = min(P, p0 + per);
  float s = 0.f;
  if (c < C) {
    const TIn* xb = x + ((long long)b * P) * ld + c;
    for (int p = p0 + w; p < p1; p += 8) s += (float)xb[(long long)p * ld];
  }
  red[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    if (c < C) partial[((long long)b * nsplit + split) * C + c] = t;
  }
}
// one block per (32 channels, sample): lane = channel, the 32 warps split the partial rows (4 independent loads in flight each)
constexpr int GAPF_WARPS = 32;
__global__ void __launch_bounds__(GAPF_WARPS * 32) gap_final_kernel(const float* __restrict__ partial, int C, int nsplit, float inv,
                                                                   float* __restrict__ out, int out_ld) {
  __shared__ float red[GAPF_WARPS][33];
  const int b = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5, c = blockIdx.x * 32 + lane;
  // 16 independent loads in flight per warp: a whole image (one sample, thousands of partial rows) gives this kernel only
  // C / 32 blocks, so its time is the number of dependent memory round trips per warp
  float t[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) t[i] = 0.f;
  if (c < C) {
    const float* pb = partial + (long long)b * nsplit * C + c;
    int s = w;
    for (; s + 15 * GAPF_WARPS < nsplit; s += 16 * GAPF_WARPS) {
#pragma unroll
      for (int i = 0; i < 16; ++i) t[i] += pb[(long long)(s + i * GAPF_WARPS) * C];
    }
    for (; s < nsplit; s += GAPF_WARPS) t[0] += pb[(long long)s * C];
  }
#pragma unroll
  for (int st = 8; st > 0; st >>= 1)
#pragma unroll
    for (int i = 0; i < st; ++i) t[i] += t[i + st];
  red[w][lane] = t[0];
  __syncthreads();
  if (w == 0 && c < out_ld) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < GAPF_WARPS; ++i) t += red[i][lane];
    out[(long long)b * out_ld + c] = c < C ? t * inv : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// y[r][n] = act(sum_k x[r][k] * W[n][k] + bias[n])  for a handful of rows (per-sample vectors).
// One warp per output element.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vec_linear_kernel(const float* __restrict__ x, int x_ld, int R, int K,
                                                        const float* __restrict__ W, const float* __restrict__ bias, int N,
                                                        int act, float* __restrict__ y, int y_ld, int y_cols) {
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gw >= R * y_cols) return;
  const int r = gw / y_cols, n = gw - r * y_cols;
  float s = 0.f;
  if (n < N) {
    const float* xr = x + (long long)r * x_ld;
    const float* wr = W + (long long)n * K;
    for (int k = lane; k < K; k += 32) s += xr[k] * __ldg(wr + k);
    s = warp_sum(s);
    if (bias) s += __ldg(bias + n);
    s = act_apply(s, act);
  }
  if (lane == 0) y[(long long)r * y_ld + n] = (n < N) ? s : 0.f;
}

// Second phase of a global average pool FUSED with the one- or two-layer per-sample MLP that consumes it (squeeze-excite of HAT's
// CAB, DAT's channel interaction, NAFNet's simplified channel attention): the blocks of a sample write their 32 channel means, take a
// ticket, and the LAST one runs  hidden = act1(W1 . mean + b1),  out = act2(W2 . hidden + b2)  (h1 == 0: out = act1(W1 . mean + b1)).
// Replaces three dependent micro-launches (pool finalise, two vec_linear: ~5 us of kernel + launch gap each) per block of the model.
// Deterministic: fixed summation order, no floating-point atomics.
__global__ void __launch_bounds__(GAPF_WARPS * 32) gap_final_mlp_kernel(const float* __restrict__ partial, int C, int nsplit, float inv, float* __restrict__ mean,
                                                                       int mean_ld, unsigned int* __restrict__ counters, const float* __restrict__ w1,
                                                                       const float* __restrict__ b1, int k1, int h1, int act1, const float* __restrict__ w2,
                                                                       const float* __restrict__ b2, int n_out, int act2, float* __restrict__ out, int out_ld,
                                                                       int out_cols) {
  __shared__ float red[GAPF_WARPS][33];
  __shared__ float s_in[1024];
  __shared__ float s_hid[128];
  __shared__ int s_last;
  const int b = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5, c = blockIdx.x * 32 + lane;
  float t[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) t[i] = 0.f;
  if (c < C) {
    const float* pb = partial + (long long)b * nsplit * C + c;
    int s = w;
    for (; s + 15 * GAPF_WARPS < nsplit; s += 16 * GAPF_WARPS) {
#pragma unroll
      for (int i = 0; i < 16; ++i) t[i] += pb[(long long)(s + i * GAPF_WARPS) * C];
    }
    for (; s < nsplit; s += GAPF_WARPS) t[0] += pb[(long long)s * C];
  }
#pragma unroll
  for (int st = 8; st > 0; st >>= 1)
#pragma unroll
    for (int i = 0; i < st; ++i) t[i] += t[i + st];
  red[w][lane] = t[0];
  __syncthreads();
  if (w == 0 && c < mean_ld) {
    float tt = 0.f;
#pragma unroll
    for (int i = 0; i < GAPF_WARPS; ++i) tt += red[i][lane];
    mean[(long long)b * mean_ld + c] = c < C ? tt * inv : 0.f;
  }
  // ticket: the last block of this sample sees every block's means
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int tk = atomicAdd(&counters[b], 1u);
    s_last = tk == gridDim.x - 1;
    if (s_last) counters[b] = 0u;      // every block of the sample has taken its ticket: re-arm for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int k = threadIdx.x; k < k1; k += GAPF_WARPS * 32) s_in[k] = __ldcg(mean + (long long)b * mean_ld + k);
  __syncthreads();
  if (h1 > 0) {
    for (int h = w; h < h1; h += GAPF_WARPS) {
      float s = 0.f;
      for (int k = lane; k < k1; k += 32) s += s_in[k] * __ldg(w1 + (long long)h * k1 + k);
      s = warp_sum(s);
      if (lane == 0) s_hid[h] = act_apply(s + (b1 ? __ldg(b1 + h) : 0.f), act1);
    }
    __syncthreads();
    for (int n = threadIdx.x; n < out_cols; n += GAPF_WARPS * 32) {
      float s = 0.f;
      if (n < n_out) {
        s = b2 ? __ldg(b2 + n) : 0.f;
        for (int h = 0; h < h1; ++h) s += s_hid[h] * __ldg(w2 + (long long)n * h1 + h);
        s = act_apply(s, act2);
      }
      out[(long long)b * out_ld + n] = s;
    }
  } else {
    for (int n = w; n < out_cols; n += GAPF_WARPS) {
      float s = 0.f;
      if (n < n_out) {
        for (int k = lane; k < k1; k += 32) s += s_in[k] * __ldg(w1 + (long long)n * k1 + k);
        s = warp_sum(s);
        s = act_apply(s + (b1 ? __ldg(b1 + n) : 0.f), act1);
      }
      if (lane == 0) out[(long long)b * out_ld + n] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Depthwise convolution, NHWC, 8 channels per thread (16-byte bf16 vectors).
//   mode 0: out[c] = act(dw(x)[c] + bias[c]) (* mul[c])
//   mode 1 (SimpleGate): out[c] = (dw(x)[c]+bias[c]) * (dw(x)[c+Cout]+bias[c+Cout]),  c < Cout = C/2
// Weights are fp32 [kh*kw][C] (tap-major) so channel vectors are contiguous.
// ------------------------------------------------------------------------------------------
struct DwArgs {
  const bf16* x; int x_ld;
  int B, H, W, C;
  int kh, kw;
  const float* w; const float* bias;
  int act; int mode;
  const bf16* mul; int mul_ld;
  bf16* out; int out_ld;
  float* col_sums;      // optional (TMA-staged 3x3 path): per-tile column sums of the stored values, [B][tiles_y*tiles_x][cout]
};

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// bf16 rows wider than the residual stream (DAT's SGFN norm over the gated half, 360 of 384 columns): 16 lanes per row, two rows
// per warp, lanes own channel OCTETS (c = 8*sub + 128*i): NV 16-byte loads and NV 16-byte stores per lane instead of the
// 4-byte accesses of the generic pair kernel.  C % 8 == 0, out_cols <= 128*NV, padding columns are written as zero.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_bf16_wide_kernel(const bf16* __restrict__ x, int in_ld, long long rows, int C,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                                 bf16* __restrict__ out, int out_ld, int out_cols) {
  const int lane = threadIdx.x & 31, sub = lane & 15;
  const long long row = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * 2 + (lane >> 4);
  const bool ok = row < rows;
  const bf16* xr = x + (ok ? row : 0) * in_ld;
  float v[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = 8 * sub + 128 * i;
    if (c < C) {
      unpack8(*reinterpret_cast<const uint4*>(xr + c), v[i]);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[i][e];
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (8 * sub + 128 * i < C) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { v[i][e] -= mean; q += v[i][e] * v[i][e]; }
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  if (!ok) return;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = 8 * sub + 128 * i;
    if (c >= out_cols) continue;
    float y[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c < C) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c) + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = v[i][e] * rstd * gg[e] + bb[e];
    }
    *reinterpret_cast<uint4*>(out + row * out_ld + c) = pack8(y);
  }
}

__device__ __forceinline__ void dw_accum(const DwArgs& a, int b, int y, int x, int c0, float (&acc)[8]) {
  const int ph = a.kh >> 1, pw = a.kw >> 1;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = a.bias ? __ldg(a.bias + c0 + i) : 0.f;
  for (int dy = 0; dy < a.kh; ++dy) {
    const int yy = y + dy - ph;
    if (yy < 0 || yy >= a.H) continue;
    for (int dx = 0; dx < a.kw; ++dx) {
      const int xx = x + dx - pw;
      if (xx < 0 || xx >= a.W) continue;
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(a.x + ((long long)(b * a.H + yy) * a.W + xx) * a.x_ld + c0));
      float f[8];
      unpack8(q, f);
      const float* wp = a.w + (long long)(dy * a.kw + dx) * a.C + c0;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp) + 1);
      acc[0] += f[0] * w0.x; acc[1] += f[1] * w0.y; acc[2] += f[2] * w0.z; acc[3] += f[3] * w0.w;
      acc[4] += f[4] * w1.x; acc[5] += f[5] * w1.y; acc[6] += f[6] * w1.z; acc[7] += f[7] * w1.w;
    }
  }
}

__global__ void __launch_bounds__(256) dwconv_kernel(const __grid_constant__ DwArgs a) {
  const int cout = a.mode == 1 ? a.C / 2 : a.C;
  const int groups = cout >> 3;
  const long long total = (long long)a.B * a.H * a.W * groups;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const long long pix = idx / groups;
  const int x = (int)(pix % a.W);
  const int y = (int)((pix / a.W) % a.H);
  const int b = (int)(pix / ((long long)a.W * a.H));
  const int c0 = g * 8;
  float acc[8];
  dw_accum(a, b, y, x, c0, acc);
  if (a.mode == 1) {
    float acc2[8];
    dw_accum(a, b, y, x, c0 + cout, acc2);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] *= acc2[i];
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = act_apply(acc[i], a.act);
    if (a.mul) {
      float m[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(a.mul + pix * a.mul_ld + c0)), m);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] *= m[i];
    }
  }
  *reinterpret_cast<uint4*>(a.out + pix * a.out_ld + c0) = pack8(acc);
}


// 3x3 specialisation: one thread = 8 channels x 4 consecutive output pixels of one row; the 3x6 input patch and the 9 weight
// vectors are loaded once (18 + 18 vector loads for 4 outputs instead of 36 + 72) and the MACs run as packed fp32x2 FMAs
// (FFMA2, sm_100): 144 FFMA2 per thread instead of 288 FFMA -- the kernel is instruction-issue bound, not HBM bound.
__device__ __forceinline__ void unpack8_f2(const uint4& q, float2 (&f)[4]) {
  f[0] = make_float2(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u));
  f[1] = make_float2(__uint_as_float(q.y << 16), __uint_as_float(q.y & 0xffff0000u));
  f[2] = make_float2(__uint_as_float(q.z << 16), __uint_as_float(q.z & 0xffff0000u));
  f[3] = make_float2(__uint_as_float(q.w << 16), __uint_as_float(q.w & 0xffff0000u));
}

// ACT: 0 none, 1 GELU (hardware tanh form), 2 runtime a.act
template <int MODE, int ACT>
__global__ void __launch_bounds__(128, MODE == 1 ? 2 : 3) dwconv3x3_kernel(const __grid_constant__ DwArgs a) {
  const int cout = MODE == 1 ? a.C / 2 : a.C;
  const int groups = cout >> 3;
  const int wq = a.W >> 2;
  const long long total = (long long)a.B * a.H * wq * groups;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  long long r = idx / groups;
  const int xq = (int)(r % wq);
  r /= wq;
  const int y = (int)(r % a.H), b = (int)(r / a.H);
  const int x0 = xq * 4;
  const int c0 = g * 8;
  constexpr int NH = MODE == 1 ? 2 : 1;
  float2 acc[NH][4][4];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const int cc = c0 + h * cout;
    float2 bv[4];
    if (a.bias) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + cc)), b1 = __ldg(reinterpret_cast<const float4*>(a.bias + cc) + 1);
      bv[0] = make_float2(b0.x, b0.y); bv[1] = make_float2(b0.z, b0.w); bv[2] = make_float2(b1.x, b1.y); bv[3] = make_float2(b1.z, b1.w);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) bv[i] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[h][px][i] = bv[i];
    // all 18 input vectors are requested before any math (branch-free clamped addresses; out-of-image taps are zeroed
    // afterwards) so each warp keeps 18 x 512 B in flight -- this kernel lives or dies by memory-level parallelism
    uint4 in[3][6];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = min(max(y + dy - 1, 0), a.H - 1);
      const bf16* rowp = a.x + ((long long)(b * a.H + yy) * a.W) * a.x_ld + cc;
#pragma unroll
      for (int col = 0; col < 6; ++col) {
        const int xx = min(max(x0 + col - 1, 0), a.W - 1);
        in[dy][col] = __ldg(reinterpret_cast<const uint4*>(rowp + (long long)xx * a.x_ld));
      }
    }
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      const bool yok = yy >= 0 && yy < a.H;
      float2 wv[3][4];
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float* wp = a.w + (long long)(dy * 3 + dx) * a.C + cc;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp)), w1 = __ldg(reinterpret_cast<const float4*>(wp) + 1);
        wv[dx][0] = make_float2(w0.x, w0.y); wv[dx][1] = make_float2(w0.z, w0.w); wv[dx][2] = make_float2(w1.x, w1.y); wv[dx][3] = make_float2(w1.z, w1.w);
      }
#pragma unroll
      for (int col = 0; col < 6; ++col) {
        const int xx = x0 + col - 1;
        const bool ok = yok && xx >= 0 && xx < a.W;
        const uint4 q = ok ? in[dy][col] : make_uint4(0, 0, 0, 0);
        float2 f[4];
        unpack8_f2(q, f);
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          const int dx = col - px;          // input column col contributes to output px with tap dx
          if (dx >= 0 && dx < 3) {
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[h][px][i] = __ffma2_rn(f[i], wv[dx][i], acc[h][px][i]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int px = 0; px < 4; ++px) {
    const long long pix = (long long)(b * a.H + y) * a.W + x0 + px;
    float o[8];
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { o[2 * i] = acc[0][px][i].x * acc[NH - 1][px][i].x; o[2 * i + 1] = acc[0][px][i].y * acc[NH - 1][px][i].y; }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v0 = acc[0][px][i].x, v1 = acc[0][px][i].y;
        if constexpr (ACT == 1) { v0 = gelu_tanh_hw(v0); v1 = gelu_tanh_hw(v1); }
        else if constexpr (ACT == 2) { v0 = act_apply(v0, a.act); v1 = act_apply(v1, a.act); }
        o[2 * i] = v0; o[2 * i + 1] = v1;
      }
      if (a.mul) {
        float m[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(a.mul + pix * a.mul_ld + c0)), m);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] *= m[i];
      }
    }
    *reinterpret_cast<uint4*>(a.out + pix * a.out_ld + c0) = pack8(o);
  }
}

// ------------------------------------------------------------------------------------------
// 3x3 depthwise conv, TMA-staged.  The register kernel above is latency bound (a warp issues its loads, waits a full
// memory round trip, computes, and only then asks for more: ~1.5-2.4 TB/s).  Here persistent CTAs stream halo tiles
// [10 rows][34 px][64 ch] (hardware zero fill = the conv padding), plus the [8][32][64] tile of the fused multiplier when
// there is one, through a shared-memory ring filled by TMA, so the next tiles are in flight while one is consumed.
// 512 threads: a thread owns one pixel column x 4 channels and slides down the 8 output rows with the unpacked 3x3
// window and its 36 weights in registers.
// MODE 1 = SimpleGate pair (NAFNet): the tile holds 32 channels of each half; partner lanes meet through a shuffle.
// ------------------------------------------------------------------------------------------
constexpr int DWT_TX = 32, DWT_TY = 8;
constexpr int DWT_THREADS = 512;
constexpr int DWT_TILE_BYTES = (DWT_TY + 2) * (DWT_TX + 2) * 64 * 2;      // 43520
constexpr int DWT_HALF_BYTES = DWT_TILE_BYTES / 2;                        // MODE 1: one 32-channel box per half (21760 = 170 * 128)
constexpr int DWT_MUL_BYTES = DWT_TY * DWT_TX * 64 * 2;                   // 32768
__device__ __forceinline__ void unpack4_f2(const uint2& q, float2 (&f)[2]) {
  f[0] = make_float2(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u));
  f[1] = make_float2(__uint_as_float(q.y << 16), __uint_as_float(q.y & 0xffff0000u));
}
template <bool MUL> struct DwtCfg {
  static constexpr int STAGES = MUL ? 2 : 3;
  static constexpr int STAGE_BYTES = DWT_TILE_BYTES + (MUL ? DWT_MUL_BYTES : 0);
  static constexpr int SMEM = STAGES * STAGE_BYTES + 128;
};

template <int MODE, int ACT, bool MUL>
__global__ void __launch_bounds__(DWT_THREADS, 1) dwconv3x3_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmM,
                                                                       const __grid_constant__ DwArgs a, int tiles_x, int tiles_y, int ctiles) {
  using Cf = DwtCfg<MUL>;
  extern __shared__ uint8_t dw_smem_raw[];
  __shared__ __align__(8) uint64_t full[Cf::STAGES];
  __shared__ float4 pool_red[2][DWT_THREADS / 32][16];      // per-warp column sums of a tile (global-average-pool partials), by tile parity
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dw_smem_raw) + 127) & ~(uintptr_t)127);
  const int tid = threadIdx.x;
  const int cout = MODE == 1 ? a.C / 2 : a.C;
  const int num_tiles = a.B * tiles_y * tiles_x * ctiles;
  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    if (MUL) tma_prefetch_desc(&tmM);
    for (int s = 0; s < Cf::STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int tile, int s) {     // one thread
    const int ct = tile % ctiles;
    int r = tile / ctiles;
    const int tx = r % tiles_x;
    r /= tiles_x;
    const int ty = r % tiles_y, b = r / tiles_y;
    uint8_t* dst = smem + s * Cf::STAGE_BYTES;
    mbar_arrive_expect_tx(&full[s], Cf::STAGE_BYTES);
    if constexpr (MODE == 1) {
      tma_load_4d(dst, &tmX, &full[s], ct * 32, tx * DWT_TX - 1, ty * DWT_TY - 1, b);
      tma_load_4d(dst + DWT_HALF_BYTES, &tmX, &full[s], cout + ct * 32, tx * DWT_TX - 1, ty * DWT_TY - 1, b);
    } else {
      tma_load_4d(dst, &tmX, &full[s], ct * 64, tx * DWT_TX - 1, ty * DWT_TY - 1, b);
    }
    if constexpr (MUL) tma_load_4d(dst + DWT_TILE_BYTES, &tmM, &full[s], ct * 64, tx * DWT_TX, ty * DWT_TY, b);
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < Cf::STAGES - 1; ++s) {
      const int t = blockIdx.x + s * gridDim.x;
      if (t < num_tiles) issue(t, s);
    }
  }
  // thread -> (pixel column, 4-channel group).  MODE 0: 16 groups per pixel (128-B pixel pitch), a half warp reads one pixel.
  // MODE 1: lane = sub(3 bits) | px&1 << 3 | half << 4, so that a half warp reads two adjacent 64-B pixels of ONE half
  // (no bank conflicts) and the SimpleGate partner is lane ^ 16.
  int px, sub, half;
  if constexpr (MODE == 1) {
    sub = tid & 7; half = (tid >> 4) & 1; px = ((tid >> 3) & 1) | ((tid >> 5) << 1);
  } else {
    sub = tid & 15; half = 0; px = tid >> 4;
  }
  constexpr int PP = MODE == 1 ? 64 : 128;
  const int toff = half * DWT_HALF_BYTES + px * PP + sub * 8;
  int stage = 0;
  uint32_t phase = 0;
  float2 w[9][2], bv[2];
  int w_ct = -1;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int ct = tile % ctiles;
    int r = tile / ctiles;
    const int tx = r % tiles_x;
    r /= tiles_x;
    const int ty = r % tiles_y, b = r / tiles_y;
    const int cin0 = MODE == 1 ? half * cout + ct * 32 + sub * 4 : ct * 64 + sub * 4;     // input channel of this thread
    const int co0 = MODE == 1 ? ct * 32 + sub * 4 : cin0;                                   // output channel
    // the grid is a multiple of the channel-tile count whenever possible, so a CTA keeps one channel tile and its weights
    if (ct != w_ct) {
      w_ct = ct;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(a.w + (long long)t * a.C + cin0));
        w[t][0] = make_float2(q.x, q.y); w[t][1] = make_float2(q.z, q.w);
      }
      if (a.bias) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(a.bias + cin0));
        bv[0] = make_float2(q.x, q.y); bv[1] = make_float2(q.z, q.w);
      } else {
        bv[0] = bv[1] = make_float2(0.f, 0.f);
      }
    }
    // refill the stage consumed in the previous iteration (every thread passed the barrier at its end)
    if (tid == 0) {
      const int t2 = tile + (Cf::STAGES - 1) * gridDim.x;
      if (t2 < num_tiles) issue(t2, (stage + Cf::STAGES - 1) % Cf::STAGES);
    }
    mbar_wait(&full[stage], phase);
    const uint8_t* base = smem + stage * Cf::STAGE_BYTES + toff;
    const uint8_t* mbase = smem + stage * Cf::STAGE_BYTES + DWT_TILE_BYTES + px * 128 + sub * 8;
    (void)mbase;
    float2 win[3][3][2];      // [row slot][column][channel pair], unpacked
    auto load_row = [&](int slot, int hr) {
#pragma unroll
      for (int c = 0; c < 3; ++c) unpack4_f2(*reinterpret_cast<const uint2*>(base + (hr * (DWT_TX + 2) + c) * PP), win[slot][c]);
    };
    load_row(0, 0);
    load_row(1, 1);
    bf16* op = a.out + ((long long)(b * a.H + ty * DWT_TY) * a.W + tx * DWT_TX + px) * a.out_ld + co0;
    float4 psum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int row = 0; row < DWT_TY; ++row) {
      load_row((row + 2) % 3, row + 2);
      float2 acc[2] = {bv[0], bv[1]};
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          acc[0] = __ffma2_rn(win[(row + dy) % 3][dx][0], w[dy * 3 + dx][0], acc[0]);
          acc[1] = __ffma2_rn(win[(row + dy) % 3][dx][1], w[dy * 3 + dx][1], acc[1]);
        }
      float o[4] = {acc[0].x, acc[0].y, acc[1].x, acc[1].y};
      bool writer = true;
      if constexpr (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] *= __shfl_xor_sync(0xffffffffu, o[i], 16);
        writer = half == 0;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if constexpr (ACT == 1) o[i] = gelu_tanh_hw(o[i]);
          else if constexpr (ACT == 2) o[i] = act_apply(o[i], a.act);
        }
        if constexpr (MUL) {
          float2 m[2];
          unpack4_f2(*reinterpret_cast<const uint2*>(mbase + row * (DWT_TX * 128)), m);
          o[0] *= m[0].x; o[1] *= m[0].y; o[2] *= m[1].x; o[3] *= m[1].y;
        }
      }
      psum.x += o[0]; psum.y += o[1]; psum.z += o[2]; psum.w += o[3];
      if (writer && tx * DWT_TX + px < a.W && ty * DWT_TY + row < a.H) {      // (partial edge tiles: the TMA load zero-filled the rest)
        const __nv_bfloat162 lo = __floats2bfloat162_rn(o[0], o[1]), hi = __floats2bfloat162_rn(o[2], o[3]);
        *reinterpret_cast<uint2*>(op + (long long)row * a.W * a.out_ld) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
      }
    }
    const int par = ((tile - (int)blockIdx.x) / (int)gridDim.x) & 1;
    if (a.col_sums) {
      // a warp holds two pixel columns of every channel group: fold them, then leave one float4 per (warp, channel group)
      constexpr int XOR = MODE == 1 ? 8 : 16;
      psum.x += __shfl_xor_sync(0xffffffffu, psum.x, XOR); psum.y += __shfl_xor_sync(0xffffffffu, psum.y, XOR);
      psum.z += __shfl_xor_sync(0xffffffffu, psum.z, XOR); psum.w += __shfl_xor_sync(0xffffffffu, psum.w, XOR);
      const int lane = tid & 31;
      if (lane < (MODE == 1 ? 8 : 16)) pool_red[par][tid >> 5][lane] = psum;
    }
    __syncthreads();      // everyone is done with this stage before it is refilled in the next iteration
    if (a.col_sums) {
      constexpr int GROUPS = MODE == 1 ? 8 : 16;             // 4-channel groups per tile
      if (tid < GROUPS) {
        float4 t = pool_red[par][0][tid];
#pragma unroll
        for (int w2 = 1; w2 < DWT_THREADS / 32; ++w2) {
          const float4 q = pool_red[par][w2][tid];
          t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w;
        }
        const long long prow = ((long long)b * tiles_y + ty) * tiles_x + tx;
        *reinterpret_cast<float4*>(a.col_sums + prow * cout + ct * (MODE == 1 ? 32 : 64) + tid * 4) = t;
      }
    }
    if (++stage == Cf::STAGES) { stage = 0; phase ^= 1; }
  }
}

constexpr int DWG_TY = 16;      // taller tiles than the MODE 0 kernels: the per-tile index arithmetic and the 2-row halo amortise over twice the rows
struct DwgCfg {
  static constexpr int STAGES = 2;
  static constexpr int HALF_BYTES = (DWG_TY + 2) * (DWT_TX + 2) * 32 * 2;      // one 32-channel box
  static constexpr int STAGE_BYTES = 2 * HALF_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 128;
};
// SimpleGate depthwise 3x3 (NAFBlock conv2 + x1 * x2, nafnet_arch.py:116-118), TMA-staged like the kernel above but with BOTH halves
// of a gate pair in one thread: thread = (pixel column, channel pair) reads the two bf16 of x1 and the two of x2 with one LDS.32
// each, slides down the 16 rows of the tile with two 3x3 windows in registers (18 FFMA2 per output row), multiplies the halves and
// stores one bf16 pair -- no cross-lane shuffles, no duplicated gate work.  (The MODE 1 path above spent 93 warp-instructions per
// 4-channel item against 18 FFMA2 of real work: issue bound at 31 % of DRAM.)
__global__ void __launch_bounds__(DWT_THREADS, 1) dwconv3x3_gate_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ DwArgs a, int tiles_x,
                                                                            int tiles_y, int ctiles) {
  using Cf = DwgCfg;
  extern __shared__ uint8_t dwg_smem_raw[];
  __shared__ __align__(8) uint64_t full[Cf::STAGES];
  __shared__ float2 pool_red[2][DWT_THREADS / 32][16];      // per-warp column sums of a tile (global-average-pool partials), by tile parity
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dwg_smem_raw) + 127) & ~(uintptr_t)127);
  const int tid = threadIdx.x;
  const int cout = a.C / 2;
  const int num_tiles = a.B * tiles_y * tiles_x * ctiles;
  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < Cf::STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int tile, int s) {     // one thread: the two 32-channel halves (x1 chunk, its x2 partner chunk) of a halo tile
    const int ct = tile % ctiles;
    int r = tile / ctiles;
    const int tx = r % tiles_x;
    r /= tiles_x;
    const int ty = r % tiles_y, b = r / tiles_y;
    uint8_t* dst = smem + s * Cf::STAGE_BYTES;
    mbar_arrive_expect_tx(&full[s], Cf::STAGE_BYTES);
    tma_load_4d(dst, &tmX, &full[s], ct * 32, tx * DWT_TX - 1, ty * DWG_TY - 1, b);
    tma_load_4d(dst + Cf::HALF_BYTES, &tmX, &full[s], cout + ct * 32, tx * DWT_TX - 1, ty * DWG_TY - 1, b);
  };
  if (tid == 0 && (int)blockIdx.x < num_tiles) issue(blockIdx.x, 0);
  const int g = tid & 15, px = tid >> 4;      // channel pair of the 32-channel chunk, pixel column of the tile: a warp reads 2 x 64 contiguous bytes
  const int toff = px * 64 + g * 4;
  const long long row_step = (long long)a.W * a.out_ld;
  int stage = 0;
  uint32_t phase = 0;
  float2 w1[9], w2[9], b1, b2;
  int w_ct = -1;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int ct = tile % ctiles;
    int r = tile / ctiles;
    const int tx = r % tiles_x;
    r /= tiles_x;
    const int ty = r % tiles_y, b = r / tiles_y;
    const int c1 = ct * 32 + g * 2;      // x1 input channel = output channel; the x2 partner is c1 + cout
    if (ct != w_ct) {
      w_ct = ct;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        w1[t] = __ldg(reinterpret_cast<const float2*>(a.w + (long long)t * a.C + c1));
        w2[t] = __ldg(reinterpret_cast<const float2*>(a.w + (long long)t * a.C + cout + c1));
      }
      b1 = a.bias ? __ldg(reinterpret_cast<const float2*>(a.bias + c1)) : make_float2(0.f, 0.f);
      b2 = a.bias ? __ldg(reinterpret_cast<const float2*>(a.bias + cout + c1)) : make_float2(0.f, 0.f);
    }
    if (tid == 0) {      // refill the other stage (every thread passed the barrier that ended its last use)
      const int t2 = tile + gridDim.x;
      if (t2 < num_tiles) issue(t2, stage ^ 1);
    }
    mbar_wait(&full[stage], phase);
    const uint32_t base = smem_u32(smem) + stage * Cf::STAGE_BYTES + toff;
    float2 wa[3][3], wb[3][3];      // [row slot][column] of x1 / x2, unpacked
    auto load_row = [&](int slot, int hr) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t q1, q2;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(q1) : "r"(base + (hr * (DWT_TX + 2) + c) * 64));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(q2) : "r"(base + Cf::HALF_BYTES + (hr * (DWT_TX + 2) + c) * 64));
        wa[slot][c] = make_float2(__uint_as_float(q1 << 16), __uint_as_float(q1 & 0xffff0000u));
        wb[slot][c] = make_float2(__uint_as_float(q2 << 16), __uint_as_float(q2 & 0xffff0000u));
      }
    };
    load_row(0, 0);
    load_row(1, 1);
    // rows of this tile inside the image (0 when the pixel column lies outside: partial edge tiles)
    const int rows_in = (tx * DWT_TX + px < a.W) ? min(DWG_TY, a.H - ty * DWG_TY) : 0;
    bf16* op = a.out + ((long long)(b * a.H + ty * DWG_TY) * a.W + tx * DWT_TX + px) * a.out_ld + c1;
    float2 psum = make_float2(0.f, 0.f);
#pragma unroll
    for (int row = 0; row < DWG_TY; ++row) {
      load_row((row + 2) % 3, row + 2);
      float2 acc1 = b1, acc2 = b2;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          acc1 = __ffma2_rn(wa[(row + dy) % 3][dx], w1[dy * 3 + dx], acc1);
          acc2 = __ffma2_rn(wb[(row + dy) % 3][dx], w2[dy * 3 + dx], acc2);
        }
      const float2 o = __fmul2_rn(acc1, acc2);
      if (row < rows_in) {      // rows / columns beyond the image hold bias products, not zeros: keep them out of the store and the pool
        psum = __fadd2_rn(psum, o);
        *reinterpret_cast<uint32_t*>(op) = pack_bf16(o.x, o.y);
      }
      op += row_step;
    }
    const int par = ((tile - (int)blockIdx.x) / (int)gridDim.x) & 1;
    if (a.col_sums) {
      // a warp holds two pixel columns of every channel pair: fold them, then one float2 per (warp, channel pair)
      psum.x += __shfl_xor_sync(0xffffffffu, psum.x, 16);
      psum.y += __shfl_xor_sync(0xffffffffu, psum.y, 16);
      if ((tid & 31) < 16) pool_red[par][tid >> 5][tid & 15] = psum;
    }
    __syncthreads();      // everyone is done with this stage before it is refilled in the next iteration
    if (a.col_sums && tid < 16) {
      float2 t = pool_red[par][0][tid];
#pragma unroll
      for (int w2_ = 1; w2_ < DWT_THREADS / 32; ++w2_) {
        const float2 q = pool_red[par][w2_][tid];
        t.x += q.x; t.y += q.y;
      }
      const long long prow = ((long long)b * tiles_y + ty) * tiles_x + tx;
      *reinterpret_cast<float2*>(a.col_sums + prow * cout + ct * 32 + tid * 2) = t;
    }
    stage ^= 1;
    if (stage == 0) phase ^= 1;
  }
}

static int launch_dw_gate_tma(const CUtensorMap& tx_, const DwArgs& a, int tiles_x, int tiles_y, int ctiles, int grid, cudaStream_t st) {
  static FFPerDeviceFlag configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv3x3_gate_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DwgCfg::SMEM);
    if (e != cudaSuccess) { ff_set_error("ff_dwconv: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FF_ERR_CUDA; }
    configured = true;
  }
  dwconv3x3_gate_tma_kernel<<<grid, DWT_THREADS, DwgCfg::SMEM, st>>>(tx_, a, tiles_x, tiles_y, ctiles);
  return FF_OK;
}

template <int MODE, int ACT, bool MUL>
static int launch_dw_tma(const CUtensorMap& tx_, const CUtensorMap& tm_, const DwArgs& a, int tiles_x, int tiles_y, int ctiles, int grid, cudaStream_t st) {
  static FFPerDeviceFlag configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv3x3_tma_kernel<MODE, ACT, MUL>, cudaFuncAttributeMaxDynamicSharedMemorySize, DwtCfg<MUL>::SMEM);
    if (e != cudaSuccess) { ff_set_error("ff_dwconv: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FF_ERR_CUDA; }
    configured = true;
  }
  dwconv3x3_tma_kernel<MODE, ACT, MUL><<<grid, DWT_THREADS, DwtCfg<MUL>::SMEM, st>>>(tx_, tm_, a, tiles_x, tiles_y, ctiles);
  return FF_OK;
}

// ------------------------------------------------------------------------------------------
// Large-kernel depthwise conv (the 5x5 / 1x21 / 21x1 chain of the fusion head's large-kernel attention), TMA-staged.
// Persistent CTAs stream zero-filled halo tiles [TY+KH-1][TX+KW-1][64 ch] through a two-stage shared-memory ring.
// A warp owns whole pixels: lane = channel pair, so one LDS.32 per lane reads a 128-byte pixel chunk conflict-free and
// the store of a pixel is one 128-byte line.  A thread slides along a run of 16 outputs on the long axis of the kernel
// with its K x Kc weights (float2 per tap) in registers: every staged input is read once per cross tap and feeds up to K
// packed FFMA2, so the kernel needs ~28 (21 taps) / ~44 (5x5) issue slots per output pair and stays under the HBM time.
// ------------------------------------------------------------------------------------------
constexpr int DWL_THREADS = 512;
constexpr int DWL_R = 16;
template <int KH, int KW, int TY, int TX>
struct DwlCfg {
  static constexpr bool ALONG_X = KW >= KH;
  static constexpr int HH = TY + KH - 1, HW = TX + KW - 1;
  static constexpr int TILE_BYTES = HH * HW * 128;
  static constexpr int STAGES = 2;
  static constexpr int SMEM = STAGES * TILE_BYTES + 128;
  static constexpr int K = ALONG_X ? KW : KH;     // taps along the run
  static constexpr int KC = ALONG_X ? KH : KW;    // taps across it
  static constexpr int RUNS = TY * TX / DWL_R;
  static_assert((ALONG_X ? TX : TY) % DWL_R == 0, "tile must hold whole runs");
  static_assert(SMEM <= 227 * 1024, "halo ring exceeds shared memory");
};

template <int KH, int KW, int TY, int TX>
__global__ void __launch_bounds__(DWL_THREADS, 1) dwconv_large_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ DwArgs a,
                                                                          int tiles_x, int tiles_y, int ctiles) {
  using Cf = DwlCfg<KH, KW, TY, TX>;
  constexpr int K = Cf::K, KC = Cf::KC, R = DWL_R;
  extern __shared__ uint8_t dwl_smem_raw[];
  __shared__ __align__(8) uint64_t full[Cf::STAGES];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dwl_smem_raw) + 127) & ~(uintptr_t)127);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int num_tiles = a.B * tiles_y * tiles_x * ctiles;
  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < Cf::STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int tile, int s) {     // one thread
    const int ct = tile % ctiles;
    int r = tile / ctiles;
    const int tx = r % tiles_x;
    r /= tiles_x;
    const int ty = r % tiles_y, b = r / tiles_y;
    mbar_arrive_expect_tx(&full[s], Cf::TILE_BYTES);
    tma_load_4d(smem + s * Cf::TILE_BYTES, &tmX, &full[s], ct * 64, tx * TX - (KW >> 1), ty * TY - (KH >> 1), b);
  };
  if (tid == 0 && (int)blockIdx.x < num_tiles) issue(blockIdx.x, 0);
  int stage = 0;
  uint32_t phase = 0;
  float2 w[KC][K], bv = make_float2(0.f, 0.f);
  int w_ct = -1;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int ct = tile % ctiles;
    int r = tile / ctiles;
    const int tx = r % tiles_x;
    r /= tiles_x;
    const int ty = r % tiles_y, b = r / tiles_y;
    const int c0 = ct * 64 + lane * 2;
    // the other stage was released by the barrier that ended the previous iteration
    if (tid == 0) {
      const int t2 = tile + gridDim.x;
      if (t2 < num_tiles) issue(t2, stage ^ 1);
    }
    if (ct != w_ct) {       // the grid is a multiple of the channel-tile count whenever possible: one load per CTA
      w_ct = ct;
#pragma unroll
      for (int c = 0; c < KC; ++c)
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int tap = Cf::ALONG_X ? c * KW + k : k * KW + c;
          w[c][k] = __ldg(reinterpret_cast<const float2*>(a.w + (long long)tap * a.C + c0));
        }
      bv = a.bias ? __ldg(reinterpret_cast<const float2*>(a.bias + c0)) : make_float2(0.f, 0.f);
    }
    mbar_wait(&full[stage], phase);
    const uint32_t base = smem_u32(smem) + stage * Cf::TILE_BYTES + lane * 4;
#pragma unroll 1
    for (int run = warp; run < Cf::RUNS; run += DWL_THREADS / 32) {
      int oy, ox;           // first output pixel of the run (tile coordinates)
      if constexpr (Cf::ALONG_X) { oy = run / (TX / R); ox = (run - oy * (TX / R)) * R; }
      else { oy = (run / TX) * R; ox = run - (run / TX) * TX; }
      float2 acc[R];
#pragma unroll
      for (int o = 0; o < R; ++o) acc[o] = bv;
      const uint32_t rbase = base + (oy * Cf::HW + ox) * 128;      // halo pixel of (first output, tap 0); the rest are immediates
#pragma unroll
      for (int c = 0; c < KC; ++c) {
#pragma unroll
        for (int i = 0; i < R + K - 1; ++i) {
          const int off = (Cf::ALONG_X ? c * Cf::HW + i : i * Cf::HW + c) * 128;
          uint32_t q;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(q) : "r"(rbase + off));
          const float2 v = make_float2(__uint_as_float(q << 16), __uint_as_float(q & 0xffff0000u));
#pragma unroll
          for (int o = 0; o < R; ++o) {
            if (i - o >= 0 && i - o < K) acc[o] = __ffma2_rn(v, w[c][i - o], acc[o]);
          }
        }
      }
      const int gy = ty * TY + oy, gx = tx * TX + ox;
      bf16* op = a.out + ((long long)(b * a.H + gy) * a.W + gx) * a.out_ld + c0;
      const long long step = Cf::ALONG_X ? (long long)a.out_ld : (long long)a.W * a.out_ld;
      const int room = Cf::ALONG_X ? (gy < a.H ? a.W - gx : 0) : (gx < a.W ? a.H - gy : 0);      // outputs of this run inside the image
#pragma unroll
      for (int o = 0; o < R; ++o) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(acc[o].x, acc[o].y);
        if (o < room) *reinterpret_cast<uint32_t*>(op + o * step) = *reinterpret_cast<const uint32_t*>(&h);
      }
    }
    __syncthreads();      // everyone is done with this stage before it is refilled in the next iteration
    stage ^= 1;
    if (stage == 0) phase ^= 1;
  }
}

typedef CUresult (*DwEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DwEncodeTiledFn dw_encode_fn() {
  static DwEncodeTiledFn enc = []() -> DwEncodeTiledFn {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<DwEncodeTiledFn>(ptr);
    return nullptr;
  }();
  return enc;
}

template <int KH, int KW, int TY, int TX>
static int launch_dw_large(const DwArgs& a, cudaStream_t st) {
  using Cf = DwlCfg<KH, KW, TY, TX>;
  DwEncodeTiledFn enc = dw_encode_fn();
  if (!enc) { ff_set_error("ff_dwconv: cuTensorMapEncodeTiled entry point unavailable"); return FF_ERR_DRIVER; }
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.B};
  cuuint64_t strides[3] = {(cuuint64_t)a.x_ld * 2, (cuuint64_t)a.x_ld * 2 * a.W, (cuuint64_t)a.x_ld * 2 * a.W * a.H};
  cuuint32_t box[4] = {64, (cuuint32_t)Cf::HW, (cuuint32_t)Cf::HH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(a.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ff_set_error("ff_dwconv: cuTensorMapEncodeTiled failed with %d", (int)r); return FF_ERR_DRIVER; }
  static FFPerDeviceFlag configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dwconv_large_tma_kernel<KH, KW, TY, TX>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::SMEM);
    if (e != cudaSuccess) { ff_set_error("ff_dwconv: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FF_ERR_CUDA; }
    configured = true;
  }
  const int tiles_x = ff_cdiv(a.W, TX), tiles_y = ff_cdiv(a.H, TY), ctiles = a.C / 64;      // partial edge tiles: zero-filled loads, masked stores
  const long long ntiles = (long long)a.B * tiles_x * tiles_y * ctiles;
  int grid = (int)(ntiles < ff_num_sms() ? ntiles : ff_num_sms());
  if (grid >= 8 * ctiles) grid -= grid % ctiles;      // constant channel tile per CTA: weights are loaded once
  dwconv_large_tma_kernel<KH, KW, TY, TX><<<grid, DWL_THREADS, Cf::SMEM, st>>>(tm, a, tiles_x, tiles_y, ctiles);
  return FF_OK;
}

// x[p][c] *= s[b][c]  (bf16 in place), 8 channels per thread
__global__ void __launch_bounds__(256) scale_channels_kernel(bf16* __restrict__ x, int ld, long long P_per_b, int B, int C,
                                                            const float* __restrict__ s, int s_ld) {
  const int groups = C >> 3;
  const long long total = (long long)B * P_per_b * groups;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const long long pix = idx / groups;
  const int b = (int)(pix / P_per_b);
  uint4* ptr = reinterpret_cast<uint4*>(x + pix * ld + g * 8);
  float f[8];
  unpack8(*ptr, f);
  const float* sp = s + (long long)b * s_ld + g * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] *= __ldg(sp + i);
  *ptr = pack8(f);
}

// ------------------------------------------------------------------------------------------
// Direct convolution (1x1 or 3x3, zero pad) in fp32 for small channel counts.
// Block = 16x8 output pixels, each thread computes 8 output channels of one pixel; grid.y walks
// groups of 8 output channels.  Input patch and the weight slice are staged in shared memory.
// ------------------------------------------------------------------------------------------
struct DirectArgs {
  const void* x; int x_is_bf16; int x_ld;
  int B, H, W, Cin, k;
  const float* w;     // [Cout_pad][k*k*Cin], Cout_pad multiple of 8 (zero rows as padding)
  const float* bias;  // [Cout_pad] or null
  int Cout_pad, n_store;
  int act;
  const float* mul_f32; int mul_ld;   // optional multiply by an fp32 NHWC tensor (same pixel, channel n) after act
  bf16* out_bf16; int out_ld;
  float* out_f32; int out_f32_ld;
  int vec_bf16, vec_f32;   // 16-byte aligned rows: vector stores allowed
};

// R = output rows per thread (tile = 16 x 8R pixels).  R = 4 for the small-Cin layers that run at HR resolution: the weight
// staging is amortised over four times the pixels and every weight vector read from smem feeds four pixels.
template <int R>
__global__ void __launch_bounds__(128) conv_direct_kernel(const __grid_constant__ DirectArgs a) {
  extern __shared__ float sm[];
  const int pad = a.k >> 1;
  const int PW = 16 + 2 * pad, PH = 8 * R + 2 * pad;
  const int cs = a.Cin | 1;  // odd pitch -> conflict-free
  const int KK = a.k * a.k * a.Cin;
  const bool all_w = gridDim.y == 1;          // small Cin: every output-channel group's weights are staged once
  float* sIn = sm;                       // [PH*PW][cs]
  float* sW = sm + PH * PW * cs;         // [groups][KK][8]
  const int tiles_x = (a.W + 15) / 16, tiles_y = (a.H + 8 * R - 1) / (8 * R);      // partial edge tiles: loads zero-fill, stores are masked
  int t = blockIdx.x;
  const int b = t / (tiles_x * tiles_y);
  t -= b * tiles_x * tiles_y;
  const int ty = t / tiles_x, tx = t - ty * tiles_x;
  const int y0 = ty * 8 * R - pad, x0 = tx * 16 - pad;
  for (int i = threadIdx.x; i < PH * PW * a.Cin; i += 128) {
    const int c = i % a.Cin, pp = i / a.Cin;
    const int py = pp / PW, px = pp - py * PW;
    const int y = y0 + py, x = x0 + px;
    float v = 0.f;
    if (y >= 0 && y < a.H && x >= 0 && x < a.W) {
      const long long off = ((long long)(b * a.H + y) * a.W + x) * a.x_ld + c;
      v = a.x_is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(a.x)[off]) : reinterpret_cast<const float*>(a.x)[off];
    }
    sIn[pp * cs + c] = v;
  }
  const int g_begin = all_w ? 0 : blockIdx.y, g_end = all_w ? a.Cout_pad / 8 : blockIdx.y + 1;
  for (int i = threadIdx.x; i < (g_end - g_begin) * KK * 8; i += 128) {
    const int o = i & 7, kk = (i >> 3) % KK, g = (i >> 3) / KK;
    sW[(g * KK + kk) * 8 + o] = __ldg(a.w + (long long)((g_begin + g) * 8 + o) * KK + kk);
  }
  __syncthreads();
  const int py = threadIdx.x >> 4, px = threadIdx.x & 15;
  for (int grp = g_begin; grp < g_end; ++grp) {
    const int n0 = grp * 8;
    if (n0 >= a.n_store) break;
    const float* wg = sW + (long long)(grp - g_begin) * KK * 8;
    float accr[R][8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const float bvv = a.bias ? __ldg(a.bias + n0 + o) : 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) accr[r][o] = bvv;
    }
    for (int dy = 0; dy < a.k; ++dy)
      for (int dx = 0; dx < a.k; ++dx) {
        const float* ip = sIn + ((py + dy) * PW + px + dx) * cs;
        const float* wp = wg + (dy * a.k + dx) * a.Cin * 8;
        for (int c = 0; c < a.Cin; ++c) {
          const float4 w0 = *reinterpret_cast<const float4*>(wp + c * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(wp + c * 8 + 4);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float xv = ip[r * 8 * PW * cs + c];
            accr[r][0] += xv * w0.x; accr[r][1] += xv * w0.y; accr[r][2] += xv * w0.z; accr[r][3] += xv * w0.w;
            accr[r][4] += xv * w1.x; accr[r][5] += xv * w1.y; accr[r][6] += xv * w1.z; accr[r][7] += xv * w1.w;
          }
        }
      }
#pragma unroll
    for (int r = 0; r < R; ++r) {
    float (&acc)[8] = accr[r];
    if (ty * 8 * R + r * 8 + py >= a.H || tx * 16 + px >= a.W) continue;
    const long long opix = (long long)(b * a.H + ty * 8 * R + r * 8 + py) * a.W + tx * 16 + px;
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = act_apply(acc[o], a.act);
    if (a.mul_f32) {
      for (int o = 0; o < 8; ++o)
        if (n0 + o < a.n_store) acc[o] *= a.mul_f32[opix * a.mul_ld + n0 + o];
    }
    const bool full = n0 + 8 <= a.n_store;
    if (a.out_bf16) {
      bf16* q = a.out_bf16 + opix * a.out_ld + n0;
      if (full && a.vec_bf16) *reinterpret_cast<uint4*>(q) = pack8(acc);
      else
        for (int o = 0; o < 8; ++o)
          if (n0 + o < a.n_store) q[o] = __float2bfloat16_rn(acc[o]);
    }
    if (a.out_f32) {
      float* q = a.out_f32 + opix * a.out_f32_ld + n0;
      if (full && a.vec_f32) {
        reinterpret_cast<float4*>(q)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        reinterpret_cast<float4*>(q)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      } else
        for (int o = 0; o < 8; ++o)
          if (n0 + o < a.n_store) q[o] = acc[o];
    }
    }
  }
}

// NCHW fp32 image -> NHWC fp32 [pixels][ld] with per-channel offset subtraction (x - mean)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int B, int C, int H, int W, const float* __restrict__ sub,
                                    float* __restrict__ out, int ld) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W;
  if (idx >= total) return;
  const long long hw = (long long)H * W;
  const int b = (int)(idx / hw);
  const long long p = idx - b * hw;
  for (int c = 0; c < ld; ++c) {
    float v = 0.f;
    if (c < C) v = x[((long long)b * C + c) * hw + p] - (sub ? sub[c] : 0.f);
    out[idx * ld + c] = v;
  }
}
// NCHW fp32 image [B][C][H][W] -> NHWC fp32 [B*Hp*Wp][ld] padded on the right / bottom: mode 1 = reflect (F.pad(mode='reflect') of
// pad_to_window_size, expert_loader.py:83-91), mode 0 = zeros; (x - sub) as above
__global__ void nchw_to_nhwc_pad_kernel(const float* __restrict__ x, int B, int C, int H, int W, const float* __restrict__ sub,
                                        float* __restrict__ out, int ld, int Hp, int Wp, int mode) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * Hp * Wp) return;
  const int xo = (int)(idx % Wp), yo = (int)((idx / Wp) % Hp), b = (int)(idx / ((long long)Wp * Hp));
  int ys = yo, xs = xo;
  bool inside = yo < H && xo < W;
  if (!inside && mode == 1) {
    if (ys >= H) ys = 2 * (H - 1) - ys;
    if (xs >= W) xs = 2 * (W - 1) - xs;
    inside = true;
  }
  for (int c = 0; c < ld; ++c) {
    float v = 0.f;
    if (c < C && inside) v = x[(((long long)b * C + c) * H + ys) * W + xs] - (sub ? sub[c] : 0.f);
    out[idx * ld + c] = v;
  }
}
// NHWC fp32 [pixels][ld] (first C channels) -> NCHW fp32
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, int ld, int coff, int B, int C, int H, int W,
                                    float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * H * W;
  if (idx >= total) return;
  const long long hw = (long long)H * W;
  const int b = (int)(idx / hw);
  const long long p = idx - b * hw;
  for (int c = 0; c < C; ++c) out[((long long)b * C + c) * hw + p] = x[idx * ld + coff + c];
}

}  // namespace

extern "C" int ff_layernorm(const void* x, int x_is_bf16, int in_ld, long long rows, int C, const float* gamma,
                            const float* beta, float eps, void* out_bf16, int out_ld, int out_cols, float* out_f32,
                            int out_f32_ld, void* stream) {
  FF_CHECK_ARG(x && gamma && beta && (out_bf16 || out_f32), "ff_layernorm: null buffer");
  FF_CHECK_ARG(C > 0 && C <= 1024 && out_cols >= C && out_cols <= 1024, "ff_layernorm: C=%d out_cols=%d unsupported", C, out_cols);
  if (rows <= 0) return FF_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FF_CHECK_ARG(C % 2 == 0 && out_cols % 2 == 0 && in_ld % 2 == 0 && (!out_bf16 || out_ld % 2 == 0) && (!out_f32 || out_f32_ld % 2 == 0),
               "ff_layernorm: C, out_cols and all pitches must be even");
  FF_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 7) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 7) == 0 && (reinterpret_cast<uintptr_t>(beta) & 7) == 0,
               "ff_layernorm: x / gamma / beta must be 8-byte aligned");
  if (!x_is_bf16 && out_bf16 && !out_f32 && out_cols == C && (C == 64 || C == 128) && in_ld % 4 == 0 && out_ld % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(beta) & 15) == 0) {
    if (C == 64) layernorm_narrow_kernel<8><<<ff_cdiv(rows, 32), 256, 0, st>>>(reinterpret_cast<const float*>(x), in_ld, rows, gamma, beta, eps, reinterpret_cast<bf16*>(out_bf16), out_ld);
    else layernorm_narrow_kernel<16><<<ff_cdiv(rows, 16), 256, 0, st>>>(reinterpret_cast<const float*>(x), in_ld, rows, gamma, beta, eps, reinterpret_cast<bf16*>(out_bf16), out_ld);
    ++g_ff_launches;
    FF_CHECK_LAUNCH("ff_layernorm");
    return FF_OK;
  }
  auto a16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!x_is_bf16 && out_cols == 192 && C % 4 == 0 && in_ld % 4 == 0 && a16(x) && a16(gamma) && a16(beta) && (!out_bf16 || (out_ld % 4 == 0 && a16(out_bf16))) &&
      (!out_f32 || (out_f32_ld % 4 == 0 && a16(out_f32)))) {
    layernorm_w192_kernel<<<ff_cdiv(rows, 16), 256, 0, st>>>(reinterpret_cast<const float*>(x), in_ld, rows, C, gamma, beta, eps,
                                                              reinterpret_cast<bf16*>(out_bf16), out_ld, out_f32, out_f32_ld);
    ++g_ff_launches;
    FF_CHECK_LAUNCH("ff_layernorm");
    return FF_OK;
  }
  if (x_is_bf16 && out_bf16 && !out_f32 && C % 8 == 0 && out_cols % 8 == 0 && out_cols > 128 && out_cols <= 512 && in_ld % 8 == 0 && out_ld % 8 == 0 &&
      a16(x) && a16(out_bf16) && a16(gamma) && a16(beta)) {
    const bf16* xb = reinterpret_cast<const bf16*>(x);
    bf16* ob = reinterpret_cast<bf16*>(out_bf16);
    const int nv = ff_cdiv(out_cols, 128), nb = ff_cdiv(rows, 16);
    if (nv == 2) layernorm_bf16_wide_kernel<2><<<nb, 256, 0, st>>>(xb, in_ld, rows, C, gamma, beta, eps, ob, out_ld, out_cols);
    else if (nv == 3) layernorm_bf16_wide_kernel<3><<<nb, 256, 0, st>>>(xb, in_ld, rows, C, gamma, beta, eps, ob, out_ld, out_cols);
    else layernorm_bf16_wide_kernel<4><<<nb, 256, 0, st>>>(xb, in_ld, rows, C, gamma, beta, eps, ob, out_ld, out_cols);
    ++g_ff_launches;
    FF_CHECK_LAUNCH("ff_layernorm");
    return FF_OK;
  }
  const int grid = ff_cdiv(rows, 8);
  const int maxv = ff_cdiv(out_cols, 64);
#define LN_LAUNCH(T, MV) layernorm_kernel<T, MV><<<grid, 256, 0, st>>>(reinterpret_cast<const T*>(x), in_ld, rows, C, gamma, beta, eps, reinterpret_cast<bf16*>(out_bf16), out_ld, out_cols, out_f32, out_f32_ld)
  if (x_is_bf16) {
    if (maxv <= 1) LN_LAUNCH(bf16, 1); else if (maxv <= 2) LN_LAUNCH(bf16, 2); else if (maxv <= 3) LN_LAUNCH(bf16, 3); else if (maxv <= 4) LN_LAUNCH(bf16, 4); else if (maxv <= 6) LN_LAUNCH(bf16, 6); else if (maxv <= 8) LN_LAUNCH(bf16, 8); else LN_LAUNCH(bf16, 16);
  } else {
    if (maxv <= 1) LN_LAUNCH(float, 1); else if (maxv <= 2) LN_LAUNCH(float, 2); else if (maxv <= 3) LN_LAUNCH(float, 3); else if (maxv <= 4) LN_LAUNCH(float, 4); else if (maxv <= 6) LN_LAUNCH(float, 6); else if (maxv <= 8) LN_LAUNCH(float, 8); else LN_LAUNCH(float, 16);
  }
#undef LN_LAUNCH
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_layernorm");
  return FF_OK;
}

extern "C" int ff_gap(const void* x, int x_is_bf16, int ld, int B, int P, int C, float* out, int out_ld, float* scratch,
                      size_t scratch_bytes, void* stream) {
  FF_CHECK_ARG(x && out && scratch, "ff_gap: null buffer");
  FF_CHECK_ARG(out_ld >= C, "ff_gap: out_ld < C");
  int nsplit = P / 512;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > 64) nsplit = 64;
  FF_CHECK_ARG(scratch_bytes >= (size_t)B * nsplit * C * sizeof(float), "ff_gap: scratch too small (%zu)", scratch_bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(ff_cdiv(C, 32), nsplit, B);
  if (x_is_bf16) gap_partial_kernel<bf16><<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), ld, P, C, nsplit, scratch);
  else gap_partial_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), ld, P, C, nsplit, scratch);
  gap_final_kernel<<<dim3(ff_cdiv(out_ld, 32), B), GAPF_WARPS * 32, 0, st>>>(scratch, C, nsplit, 1.0f / P, out, out_ld);
  g_ff_launches += 2;
  FF_CHECK_LAUNCH("ff_gap");
  return FF_OK;
}

extern "C" int ff_gap_finalize(const float* partial, int B, int nsplit, int C, float inv, float* out, int out_ld, void* stream) {
  FF_CHECK_ARG(partial && out && B > 0 && nsplit > 0 && out_ld >= C, "ff_gap_finalize: bad args");
  gap_final_kernel<<<dim3(ff_cdiv(out_ld, 32), B), GAPF_WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(partial, C, nsplit, inv, out, out_ld);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_gap_finalize");
  return FF_OK;
}

extern "C" int ff_vec_linear(const float* x, int x_ld, int R, int K, const float* W, const float* bias, int N, int act,
                             float* y, int y_ld, int y_cols, void* stream) {
  FF_CHECK_ARG(x && W && y, "ff_vec_linear: null buffer");
  FF_CHECK_ARG(y_cols >= N && y_ld >= y_cols, "ff_vec_linear: bad y_cols/y_ld");
  vec_linear_kernel<<<ff_cdiv((long long)R * y_cols, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, x_ld, R, K, W, bias, N, act, y, y_ld, y_cols);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_vec_linear");
  return FF_OK;
}

extern "C" int ff_gap_finalize_mlp(const float* partial, int B, int nsplit, int C, float inv, float* mean, int mean_ld, unsigned int* counters,
                                   const float* w1, const float* b1, int k1, int h1, int act1, const float* w2, const float* b2, int n_out, int act2,
                                   float* out, int out_ld, int out_cols, void* stream) {
  FF_CHECK_ARG(partial && mean && counters && w1 && out && B > 0 && nsplit > 0 && mean_ld >= C, "ff_gap_finalize_mlp: bad args");
  FF_CHECK_ARG(k1 > 0 && k1 <= 1024 && k1 <= mean_ld && h1 >= 0 && h1 <= 128 && (h1 == 0 || w2) && n_out > 0 && out_cols >= n_out && out_ld >= out_cols,
               "ff_gap_finalize_mlp: k1=%d (<= 1024, <= mean_ld), h1=%d (<= 128), n_out=%d, out_cols=%d, out_ld=%d", k1, h1, n_out, out_cols, out_ld);
  gap_final_mlp_kernel<<<dim3(ff_cdiv(mean_ld, 32), B), GAPF_WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      partial, C, nsplit, inv, mean, mean_ld, counters, w1, b1, k1, h1, act1, w2, b2, n_out, act2, out, out_ld, out_cols);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_gap_finalize_mlp");
  return FF_OK;
}

static int dwconv_impl(const void* x, int x_ld, int B, int H, int W, int C, int kh, int kw, const float* w,
                       const float* bias, int act, int mode, const void* mul, int mul_ld, void* out, int out_ld,
                       float* col_sums, void* stream);

extern "C" int ff_dwconv(const void* x, int x_ld, int B, int H, int W, int C, int kh, int kw, const float* w,
                         const float* bias, int act, int mode, const void* mul, int mul_ld, void* out, int out_ld,
                         void* stream) {
  return dwconv_impl(x, x_ld, B, H, W, C, kh, kw, w, bias, act, mode, mul, mul_ld, out, out_ld, nullptr, stream);
}

extern "C" int ff_dwconv_pool(const void* x, int x_ld, int B, int H, int W, int C, const float* w, const float* bias, int act, int mode,
                              const void* mul, int mul_ld, void* out, int out_ld, float* col_sums, void* stream) {
  FF_CHECK_ARG(col_sums != nullptr, "ff_dwconv_pool: null col_sums");
  FF_CHECK_ARG(ff_dwconv_pool_rows(H, W, mode == 1 ? C / 2 : C, mode) > 0 && (reinterpret_cast<uintptr_t>(col_sums) & 15) == 0,
               "ff_dwconv_pool: %dx%d x %d channels does not tile (need H %% 8 == 0, W %% 32 == 0, 64-channel (gate: 32) chunks)", H, W, C);
  return dwconv_impl(x, x_ld, B, H, W, C, 3, 3, w, bias, act, mode, mul, mul_ld, out, out_ld, col_sums, stream);
}

static bool dw_gate2_enabled() {
  static const bool on = []() { const char* e = getenv("FFB200_DW_GATE2"); return !(e && e[0] == '0'); }();
  return on;
}
static int dw_tile_rows(int mode) { return (mode == 1 && dw_gate2_enabled()) ? DWG_TY : DWT_TY; }

extern "C" int ff_dwconv_pool_rows(int H, int W, int cout, int mode) {
  const int ty = dw_tile_rows(mode);
  const bool partial_rows_ok = mode == 1 && dw_gate2_enabled();      // the SimpleGate kernel masks its pool sums by row
  if (H <= 0 || W <= 0 || W % DWT_TX || (H % ty && !partial_rows_ok) || cout % (mode == 1 ? 32 : 64)) return 0;
  return ff_cdiv(H, ty) * (W / DWT_TX);
}

static int dwconv_impl(const void* x, int x_ld, int B, int H, int W, int C, int kh, int kw, const float* w,
                       const float* bias, int act, int mode, const void* mul, int mul_ld, void* out, int out_ld,
                       float* col_sums, void* stream) {
  FF_CHECK_ARG(x && w && out, "ff_dwconv: null buffer");
  FF_CHECK_ARG(C % 8 == 0 && x_ld % 8 == 0 && out_ld % 8 == 0 && (mode != 1 || C % 16 == 0), "ff_dwconv: channels/pitches must be multiples of 8");
  FF_CHECK_ARG((kh & 1) && (kw & 1), "ff_dwconv: odd kernel sizes only");
  DwArgs a{reinterpret_cast<const bf16*>(x), x_ld, B, H, W, C, kh, kw, w, bias, act, mode, reinterpret_cast<const bf16*>(mul), mul_ld, reinterpret_cast<bf16*>(out), out_ld, col_sums};
  const long long total = (long long)B * H * W * ((mode == 1 ? C / 2 : C) / 8);
  static const bool tma_enabled_ = []() { const char* e = getenv("FFB200_DW_TMA"); return !(e && e[0] == '0'); }();
  const bool tma3_ok = kh == 3 && kw == 3 && (tma_enabled_ || col_sums) && (mode == 1 ? C / 2 : C) % (mode == 1 ? 32 : 64) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (!mul || (mode != 1 && (reinterpret_cast<uintptr_t>(mul) & 15) == 0 && mul_ld % 8 == 0));
  if (kh == 3 && kw == 3 && (W % 4 == 0 || (tma3_ok && !col_sums))) {      // (the register-tiled fallback below needs W % 4 == 0; the TMA kernel takes any size)
    cudaStream_t st_ = reinterpret_cast<cudaStream_t>(stream);
    static const bool tma_enabled = []() { const char* e = getenv("FFB200_DW_TMA"); return !(e && e[0] == '0'); }();
    const int cout_ = mode == 1 ? C / 2 : C;
    const int TYv = dw_tile_rows(mode);
    // the fused pool needs whole tiles (the SimpleGate kernel: whole tile columns); the plain kernels mask edge tiles
    const bool tiles_ok = W % DWT_TX == 0 && (H % TYv == 0 || (mode == 1 && dw_gate2_enabled()));
    if ((tma_enabled || col_sums) && (tiles_ok || !col_sums) && cout_ % (mode == 1 ? 32 : 64) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (!mul || (mode != 1 && (reinterpret_cast<uintptr_t>(mul) & 15) == 0 && mul_ld % 8 == 0))) {
      typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
      static EncodeTiledFn enc = []() -> EncodeTiledFn {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
          return reinterpret_cast<EncodeTiledFn>(ptr);
        return nullptr;
      }();
      if (!enc) { ff_set_error("ff_dwconv: cuTensorMapEncodeTiled entry point unavailable"); return FF_ERR_DRIVER; }
      CUtensorMap tm, tmm;
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
      cuuint64_t strides[3] = {(cuuint64_t)x_ld * 2, (cuuint64_t)x_ld * 2 * W, (cuuint64_t)x_ld * 2 * W * H};
      cuuint32_t box[4] = {(cuuint32_t)(mode == 1 ? 32 : 64), DWT_TX + 2, (cuuint32_t)(TYv + 2), 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { ff_set_error("ff_dwconv: cuTensorMapEncodeTiled failed with %d", (int)r); return FF_ERR_DRIVER; }
      tmm = tm;
      if (mul) {
        cuuint64_t mstr[3] = {(cuuint64_t)mul_ld * 2, (cuuint64_t)mul_ld * 2 * W, (cuuint64_t)mul_ld * 2 * W * H};
        cuuint32_t mbox[4] = {64, DWT_TX, DWT_TY, 1};
        r = enc(&tmm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(mul), dims, mstr, mbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { ff_set_error("ff_dwconv: cuTensorMapEncodeTiled(mul) failed with %d", (int)r); return FF_ERR_DRIVER; }
      }
      const int tiles_x = ff_cdiv(W, DWT_TX), tiles_y = ff_cdiv(H, TYv), ctiles = cout_ / (mode == 1 ? 32 : 64);
      const long long ntiles = (long long)B * tiles_x * tiles_y * ctiles;
      int grid = (int)(ntiles < ff_num_sms() ? ntiles : ff_num_sms());
      if (grid >= 8 * ctiles) grid -= grid % ctiles;      // constant channel tile per CTA: weights are loaded once
      int rc;
      const bool gate2 = dw_gate2_enabled();
      if (mode == 1) rc = gate2 ? launch_dw_gate_tma(tm, a, tiles_x, tiles_y, ctiles, grid, st_) : launch_dw_tma<1, 0, false>(tm, tmm, a, tiles_x, tiles_y, ctiles, grid, st_);
      else if (mul) rc = act == FF_ACT_NONE ? launch_dw_tma<0, 0, true>(tm, tmm, a, tiles_x, tiles_y, ctiles, grid, st_)
                                            : launch_dw_tma<0, 2, true>(tm, tmm, a, tiles_x, tiles_y, ctiles, grid, st_);
      else if (act == FF_ACT_NONE) rc = launch_dw_tma<0, 0, false>(tm, tmm, a, tiles_x, tiles_y, ctiles, grid, st_);
      else if (act == FF_ACT_GELU) rc = launch_dw_tma<0, 1, false>(tm, tmm, a, tiles_x, tiles_y, ctiles, grid, st_);
      else rc = launch_dw_tma<0, 2, false>(tm, tmm, a, tiles_x, tiles_y, ctiles, grid, st_);
      if (rc != FF_OK) return rc;
      ++g_ff_launches;
      FF_CHECK_LAUNCH("ff_dwconv");
      return FF_OK;
    }
    FF_CHECK_ARG(!col_sums, "ff_dwconv_pool: operands are not 16-byte aligned for the TMA-staged path");
    const int nb = ff_cdiv(total / 4, 128);
    if (mode == 1) dwconv3x3_kernel<1, 0><<<nb, 128, 0, st_>>>(a);
    else if (act == FF_ACT_NONE) dwconv3x3_kernel<0, 0><<<nb, 128, 0, st_>>>(a);
    else if (act == FF_ACT_GELU) dwconv3x3_kernel<0, 1><<<nb, 128, 0, st_>>>(a);
    else dwconv3x3_kernel<0, 2><<<nb, 128, 0, st_>>>(a);
  } else {
    cudaStream_t st_ = reinterpret_cast<cudaStream_t>(stream);
    // large-kernel chain of the fusion head: plain depthwise conv (optional bias), 64-channel chunks, tile-aligned images
    const bool plain = mode == 0 && act == FF_ACT_NONE && !mul && C % 64 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(out) & 3) == 0 && out_ld % 2 == 0;
    int rc = 1;            // > 0: no fast path for this shape
    if (plain && kh == 5 && kw == 5) rc = launch_dw_large<5, 5, 16, 32>(a, st_);
    else if (plain && kh == 1 && kw == 21) rc = launch_dw_large<1, 21, 8, 64>(a, st_);
    else if (plain && kh == 21 && kw == 1) rc = launch_dw_large<21, 1, 64, 8>(a, st_);
    if (rc < 0) return rc;
    if (rc > 0) dwconv_kernel<<<ff_cdiv(total, 256), 256, 0, st_>>>(a);
  }
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_dwconv");
  return FF_OK;
}

// out[b][n][k] = bf16(w[n][k] * s[b][k]): per-sample copies of a small 1x1-conv weight with the channel-attention scale of
// sample b folded into its input columns, so  conv(x * s_b) = conv_b(x)  and the activation tensor is not rewritten.
__global__ void __launch_bounds__(256) scale_weight_cols_kernel(const float* __restrict__ w, int N, int K, const float* __restrict__ s, int s_ld,
                                                               int B, bf16* __restrict__ out, int n_pad, int k_pad) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)B * n_pad * k_pad;
  if (idx >= total) return;
  const int k = (int)(idx % k_pad);
  const int n = (int)((idx / k_pad) % n_pad);
  const int b = (int)(idx / ((long long)k_pad * n_pad));
  const float v = (n < N && k < K) ? __ldg(w + (long long)n * K + k) * __ldg(s + (long long)b * s_ld + k) : 0.f;
  out[idx] = __float2bfloat16_rn(v);
}

// out[b][n][k] = k < k1 ? w[n][k] : (k - k1 == n ? bf16(alpha * s[b][n]) : 0)   (FFConvGemm.x2: the aux term as a diagonal weight block)
__global__ void __launch_bounds__(256) concat_diag_weights_kernel(const bf16* __restrict__ w, int n_pad, int k1, const float* __restrict__ s, int s_ld, float alpha,
                                                                 int B, bf16* __restrict__ out) {
  const int K = k1 + n_pad;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * n_pad * K) return;
  const int k = (int)(idx % K);
  const int n = (int)((idx / K) % n_pad);
  const int b = (int)(idx / ((long long)K * n_pad));
  out[idx] = k < k1 ? w[(long long)n * k1 + k] : (k - k1 == n ? __float2bfloat16_rn(alpha * __ldg(s + (long long)b * s_ld + n)) : __float2bfloat16_rn(0.f));
}
extern "C" int ff_build_concat_diag_weights(const void* w, int n_pad, int k1, const float* s, int s_ld, float alpha, int B, void* out, void* stream) {
  FF_CHECK_ARG(w && s && out && n_pad > 0 && k1 > 0 && B > 0 && s_ld >= n_pad, "ff_build_concat_diag_weights: bad args");
  const long long total = (long long)B * n_pad * (k1 + n_pad);
  concat_diag_weights_kernel<<<ff_cdiv(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const bf16*>(w), n_pad, k1, s, s_ld, alpha, B,
                                                                                                  reinterpret_cast<bf16*>(out));
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_build_concat_diag_weights");
  return FF_OK;
}

extern "C" int ff_scale_weight_cols(const float* w, int N, int K, const float* s, int s_ld, int B, void* out, int n_pad, int k_pad, void* stream) {
  FF_CHECK_ARG(w && s && out && N > 0 && K > 0 && B > 0 && n_pad >= N && k_pad >= K, "ff_scale_weight_cols: bad args");
  const long long total = (long long)B * n_pad * k_pad;
  scale_weight_cols_kernel<<<ff_cdiv(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(w, N, K, s, s_ld, B, reinterpret_cast<bf16*>(out), n_pad, k_pad);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_scale_weight_cols");
  return FF_OK;
}

// ------------------------------------------------------------------------------------------
// fp32 -> split-bf16 operand packing for the tensor-core path of fp32 layers.
//   out[p][(t*K*K + tap)*CIN + c] = term_t(x[p + tap offset][c])  (zero outside the image, zero in the padding columns)
//   term 0 = hi = bf16(x), term 1 = lo = bf16(x - hi), term 2 = hi again.
// With the weight rows laid out [w_hi ; w_hi ; w_lo] (TERMS = 3) the bf16 GEMM accumulates a_hi*w_hi + a_lo*w_hi + a_hi*w_lo in
// fp32, i.e. ~16 mantissa bits on both operands; TERMS = 2 keeps the activation at 16 bits against bf16 weights.  K = 3 also
// gathers the 3x3 neighbourhood (im2col of a <= 7-channel image into one 64-wide k-block), so the conv runs as a 1x1 GEMM.
// 8 lanes per pixel, one 16-byte store per lane.
// ------------------------------------------------------------------------------------------
template <int CIN, int K, int TERMS>
__global__ void __launch_bounds__(256) pack_taps_kernel(const float* __restrict__ x, int x_ld, int B, int H, int W, bf16* __restrict__ out,
                                                        int out_ld, int out_cols) {
  constexpr int KK = K * K, TOT = KK * CIN * TERMS;
  const int groups = out_cols >> 3;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)B * H * W * groups;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const long long pix = idx / groups;
  const int px = (int)(pix % W), py = (int)((pix / W) % H);
  uint32_t wout[4];
#pragma unroll
  for (int e2 = 0; e2 < 4; ++e2) {
    float v[2];
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int kidx = g * 8 + e2 * 2 + h2;
      float val = 0.f;
      if (kidx < TOT) {
        const int t = kidx / (KK * CIN), r = kidx - t * (KK * CIN);
        const int tap = r / CIN, c = r - tap * CIN;
        const int dy = tap / K - K / 2, dx = tap - (tap / K) * K - K / 2;
        const int yy = py + dy, xx = px + dx;
        float xv = 0.f;
        if (K == 1 || (yy >= 0 && yy < H && xx >= 0 && xx < W)) xv = __ldg(x + (pix + (long long)dy * W + dx) * x_ld + c);
        const float hi = __bfloat162float(__float2bfloat16_rn(xv));
        val = (t == 1) ? (xv - hi) : hi;
      }
      v[h2] = val;
    }
    __nv_bfloat162 hh = __floats2bfloat162_rn(v[0], v[1]);
    wout[e2] = *reinterpret_cast<uint32_t*>(&hh);
  }
  *reinterpret_cast<uint4*>(out + pix * out_ld + g * 8) = make_uint4(wout[0], wout[1], wout[2], wout[3]);
}

// The 3-channel 3x3 im2col with two terms (CIN = 3, K = 3, TERMS = 2: 54 of 64 columns) through a shared-memory tile: a CTA covers
// 32 x 8 pixels, splits every input value of the tile (+ 1-pixel halo, zero outside the image) into (hi, lo) ONCE -- the gather form
// above does it nine times per value, each behind its own 4-byte global load -- and a lane assembles its eight bf16 from eight
// 16-bit shared-memory reads at offsets it works out once (the lane's column group is fixed).  A warp stores 512 contiguous bytes.
__global__ void __launch_bounds__(256) pack_taps3_tile_kernel(const float* __restrict__ x, int x_ld, int B, int H, int W, bf16* __restrict__ out, int out_ld) {
  constexpr int TX = 32, TY = 8, PW = TX + 2, PH = TY + 2;
  __shared__ uint16_t sT[2][PH][PW][3];
  const int tiles_x = (W + TX - 1) / TX, tiles_y = (H + TY - 1) / TY;
  int t = blockIdx.x;
  const int b = t / (tiles_x * tiles_y);
  t -= b * tiles_x * tiles_y;
  const int ty = t / tiles_x, tx = t - ty * tiles_x;
  const int y0 = ty * TY - 1, x0 = tx * TX - 1;
  for (int i = threadIdx.x; i < PH * PW; i += 256) {
    const int py = i / PW, px = i - py * PW;
    const int y = y0 + py, xx = x0 + px;
    float v[3] = {0.f, 0.f, 0.f};
    if (y >= 0 && y < H && xx >= 0 && xx < W) {
      const float* p = x + ((long long)(b * H + y) * W + xx) * x_ld;
      v[0] = __ldg(p); v[1] = __ldg(p + 1); v[2] = __ldg(p + 2);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const bf16 hi = __float2bfloat16_rn(v[c]);
      const bf16 lo = __float2bfloat16_rn(v[c] - __bfloat162float(hi));
      sT[0][py][px][c] = __bfloat16_as_ushort(hi);
      sT[1][py][px][c] = __bfloat16_as_ushort(lo);
    }
  }
  __syncthreads();
  const int g = threadIdx.x & 7, lx = threadIdx.x >> 3;      // column group (8 of the 64 columns), x inside the tile
  int off[8];                                                // element offsets into sT relative to the pixel's (row, x); -1 = zero column
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int kidx = g * 8 + e;
    const int term = kidx / 27, r = kidx - term * 27;
    const int tap = r / 3, c = r - tap * 3;
    const int dy = tap / 3, dx = tap - dy * 3;
    off[e] = kidx < 54 ? ((term * PH + dy) * PW + dx) * 3 + c : -1;
  }
  const uint16_t* base = &sT[0][0][0][0];
  const int ox = tx * TX + lx;
  if (ox >= W) return;
#pragma unroll
  for (int j = 0; j < TY; ++j) {
    const int oy = ty * TY + j;
    if (oy >= H) break;
    const uint16_t* pp = base + (j * PW + lx) * 3;
    uint32_t w[4];
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      const uint32_t lo16 = off[2 * e2] >= 0 ? pp[off[2 * e2]] : 0u;
      const uint32_t hi16 = off[2 * e2 + 1] >= 0 ? pp[off[2 * e2 + 1]] : 0u;
      w[e2] = lo16 | (hi16 << 16);
    }
    *reinterpret_cast<uint4*>(out + ((long long)(b * H + oy) * W + ox) * out_ld + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

extern "C" int ff_pack_taps(const float* x, int x_ld, int B, int H, int W, int Cin, int k, int terms, void* out, int out_ld, void* stream) {
  FF_CHECK_ARG(x && out && B > 0 && H > 0 && W > 0, "ff_pack_taps: bad args");
  const int tot = k * k * Cin * terms;
  const int out_cols = (tot + 63) / 64 * 64;
  FF_CHECK_ARG(out_ld >= out_cols && out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && x_ld >= Cin, "ff_pack_taps: out_ld=%d < %d or unaligned", out_ld, out_cols);
  const long long total = (long long)B * H * W * (out_cols / 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nb = ff_cdiv(total, 256);
  bf16* o = reinterpret_cast<bf16*>(out);
  if (Cin == 3 && k == 3 && terms == 2) pack_taps3_tile_kernel<<<B * ff_cdiv(H, 8) * ff_cdiv(W, 32), 256, 0, st>>>(x, x_ld, B, H, W, o, out_ld);
  else if (Cin == 64 && k == 1 && terms == 3) pack_taps_kernel<64, 1, 3><<<nb, 256, 0, st>>>(x, x_ld, B, H, W, o, out_ld, out_cols);
  else if (Cin == 32 && k == 1 && terms == 3) pack_taps_kernel<32, 1, 3><<<nb, 256, 0, st>>>(x, x_ld, B, H, W, o, out_ld, out_cols);
  else { ff_set_error("ff_pack_taps: (Cin=%d, k=%d, terms=%d) is not an instantiated combination", Cin, k, terms); return FF_ERR_ARG; }
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_pack_taps");
  return FF_OK;
}

extern "C" int ff_scale_channels(void* x, int ld, int B, long long pixels_per_sample, int C, const float* s, int s_ld,
                                 void* stream) {
  FF_CHECK_ARG(x && s && C % 8 == 0 && ld % 8 == 0, "ff_scale_channels: bad args");
  const long long total = (long long)B * pixels_per_sample * (C / 8);
  scale_channels_kernel<<<ff_cdiv(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<bf16*>(x), ld, pixels_per_sample, B, C, s, s_ld);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_scale_channels");
  return FF_OK;
}

// ------------------------------------------------------------------------------------------
// Specialised forms of the direct convolution for the small fp32 layers the fusion head runs at OUTPUT resolution (the edge
// refiner's attention / gate heads, edge_enhancement.py:100-106, 168-180): the generic kernel above pads every layer to groups of
// eight output channels and walks run-time loops, which costs 8x the FMAs for the 1-channel heads.
//   conv3x3_small_kernel<CIN, NOUT>  fp32 NHWC -> fp32, 3x3, zero padding.  CTA = 32 x 32 pixels, thread = a column strip of four
//       rows: per (channel, dx) six vertically adjacent inputs are read once from the shared-memory tile ([row][channel][x], lanes
//       along x: conflict free) and feed the 4 x 3 x NOUT FMAs that use them; weights are broadcast reads.
//   conv1x1_rows_kernel<CIN, NOUT>   bf16 NHWC (first CIN channels) -> fp32, one pixel per thread, weights broadcast from smem.
//   conv1x1_wide_bf16_kernel         3 fp32 channels -> 64 bf16 channels: eight lanes per pixel (16-byte stores, a warp writes
//       512 contiguous bytes), the lane's 24 weights live in registers across a grid-stride loop.
// ------------------------------------------------------------------------------------------
template <int CIN, int NOUT>
__global__ void __launch_bounds__(256) conv3x3_small_kernel(const __grid_constant__ DirectArgs a) {
  constexpr int TS = 32, PY = 4, PW = TS + 2, PH = TS + 2, XP = PW + 1;      // tile, rows per thread, tile + halo, padded x pitch
  extern __shared__ __align__(16) float sm3[];
  float* sIn = sm3;                          // [PH][CIN][XP]
  float* sW = sm3 + ((PH * CIN * XP + 3) & ~3);      // [9][CIN][NOUT], 16-byte aligned
  const int tiles_x = (a.W + TS - 1) / TS, tiles_y = (a.H + TS - 1) / TS;
  int t = blockIdx.x;
  const int b = t / (tiles_x * tiles_y);
  t -= b * tiles_x * tiles_y;
  const int ty = t / tiles_x, tx = t - ty * tiles_x;
  const int y0 = ty * TS - 1, x0 = tx * TS - 1;
  const float* xin = reinterpret_cast<const float*>(a.x);
  constexpr int CV = (CIN + 3) / 4;
  if (a.x_ld % 4 == 0 && a.x_ld >= 4 * CV && (reinterpret_cast<uintptr_t>(xin) & 15) == 0) {      // 16-byte loads (may read pitch padding past CIN)
    for (int i = threadIdx.x; i < PH * PW * CV; i += 256) {
      const int cv = i % CV, pp = i / CV;
      const int py = pp / PW, px = pp - py * PW;
      const int y = y0 + py, x = x0 + px;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < a.H && x >= 0 && x < a.W) v = __ldg(reinterpret_cast<const float4*>(xin + ((long long)(b * a.H + y) * a.W + x) * a.x_ld) + cv);
      float* d = sIn + (py * CIN + cv * 4) * XP + px;
      d[0] = v.x;
      if (cv * 4 + 1 < CIN) d[XP] = v.y;
      if (cv * 4 + 2 < CIN) d[2 * XP] = v.z;
      if (cv * 4 + 3 < CIN) d[3 * XP] = v.w;
    }
  } else {
    for (int i = threadIdx.x; i < PH * PW * CIN; i += 256) {
      const int c = i % CIN, pp = i / CIN;
      const int py = pp / PW, px = pp - py * PW;
      const int y = y0 + py, x = x0 + px;
      float v = 0.f;
      if (y >= 0 && y < a.H && x >= 0 && x < a.W) v = __ldg(xin + ((long long)(b * a.H + y) * a.W + x) * a.x_ld + c);
      sIn[(py * CIN + c) * XP + px] = v;
    }
  }
  for (int i = threadIdx.x; i < 9 * CIN * NOUT; i += 256) {
    const int o = i % NOUT, kk = i / NOUT;
    sW[i] = __ldg(a.w + (long long)o * (9 * CIN) + kk);
  }
  __syncthreads();
  const int px = threadIdx.x & 31, pg = threadIdx.x >> 5;      // a warp = one strip row group: 32 columns x rows 4 pg .. 4 pg + 3
  float acc[PY][NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    const float bv = a.bias ? __ldg(a.bias + o) : 0.f;
#pragma unroll
    for (int r = 0; r < PY; ++r) acc[r][o] = bv;
  }
  // (NOUT = 16: the fully unrolled body is ~80 KB of straight-line code, more than the instruction cache holds with eight warps at
  //  different addresses; one channel per iteration keeps it at ~13 KB)
#pragma unroll(NOUT >= 8 ? 1 : CIN)
  for (int c = 0; c < CIN; ++c) {
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      float v[PY + 2];
#pragma unroll
      for (int j = 0; j < PY + 2; ++j) v[j] = sIn[((pg * PY + j) * CIN + c) * XP + px + dx];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const float* wp = sW + ((dy * 3 + dx) * CIN + c) * NOUT;
        if constexpr (NOUT % 4 == 0) {
#pragma unroll
          for (int o = 0; o < NOUT; o += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wp + o);
#pragma unroll
            for (int r = 0; r < PY; ++r) {
              acc[r][o] = fmaf(v[r + dy], w4.x, acc[r][o]); acc[r][o + 1] = fmaf(v[r + dy], w4.y, acc[r][o + 1]);
              acc[r][o + 2] = fmaf(v[r + dy], w4.z, acc[r][o + 2]); acc[r][o + 3] = fmaf(v[r + dy], w4.w, acc[r][o + 3]);
            }
          }
        } else {
#pragma unroll
          for (int o = 0; o < NOUT; ++o) {
            const float wv = wp[o];
#pragma unroll
            for (int r = 0; r < PY; ++r) acc[r][o] = fmaf(v[r + dy], wv, acc[r][o]);
          }
        }
      }
    }
  }
  const int ox = tx * TS + px;
  if (ox >= a.W) return;
#pragma unroll
  for (int r = 0; r < PY; ++r) {
    const int oy = ty * TS + pg * PY + r;
    if (oy >= a.H) break;
    float* q = a.out_f32 + ((long long)(b * a.H + oy) * a.W + ox) * a.out_f32_ld;
    if constexpr (NOUT % 4 == 0) {
      if (a.vec_f32) {
#pragma unroll
        for (int o = 0; o < NOUT; o += 4)
          *reinterpret_cast<float4*>(q + o) = make_float4(act_apply(acc[r][o], a.act), act_apply(acc[r][o + 1], a.act), act_apply(acc[r][o + 2], a.act), act_apply(acc[r][o + 3], a.act));
        continue;
      }
    }
#pragma unroll
    for (int o = 0; o < NOUT; ++o) q[o] = act_apply(acc[r][o], a.act);
  }
}

template <int CIN, int NOUT>
__global__ void __launch_bounds__(256) conv1x1_rows_kernel(const __grid_constant__ DirectArgs a, long long pixels) {
  __shared__ __align__(16) float sW[CIN * NOUT];      // [c][o]
  __shared__ float sB[NOUT];
  for (int i = threadIdx.x; i < CIN * NOUT; i += 256) sW[i] = __ldg(a.w + (long long)(i % NOUT) * CIN + i / NOUT);
  if (threadIdx.x < NOUT) sB[threadIdx.x] = a.bias ? __ldg(a.bias + threadIdx.x) : 0.f;
  __syncthreads();
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= pixels) return;
  const uint4* xin = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.x) + p * a.x_ld);
  float acc[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; ++o) acc[o] = sB[o];
#pragma unroll
  for (int c8 = 0; c8 < CIN / 8; ++c8) {
    const uint4 u = __ldg(xin + c8);
    const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xv = __uint_as_float((j & 1) ? (uw[j >> 1] & 0xffff0000u) : (uw[j >> 1] << 16));
      const float4* wp = reinterpret_cast<const float4*>(sW + (c8 * 8 + j) * NOUT);
#pragma unroll
      for (int o = 0; o < NOUT; o += 4) {
        const float4 w4 = wp[o >> 2];
        acc[o] = fmaf(xv, w4.x, acc[o]); acc[o + 1] = fmaf(xv, w4.y, acc[o + 1]); acc[o + 2] = fmaf(xv, w4.z, acc[o + 2]); acc[o + 3] = fmaf(xv, w4.w, acc[o + 3]);
      }
    }
  }
  float* q = a.out_f32 + p * a.out_f32_ld;
#pragma unroll
  for (int o = 0; o < NOUT; o += 4)
    *reinterpret_cast<float4*>(q + o) = make_float4(act_apply(acc[o], a.act), act_apply(acc[o + 1], a.act), act_apply(acc[o + 2], a.act), act_apply(acc[o + 3], a.act));
}

// fp32 1x1 layer with C inputs and C outputs (the multi-scale mixers of the routing path, fusion_network.py:137-160, which stay fp32 in
// front of the selector's hard threshold): one pixel per thread, the pixel's C inputs and C accumulators in registers, weights as
// broadcast 16-byte shared-memory reads ([c][o] layout).  The generic kernel re-stages the input patch once per group of eight
// output channels.
template <int C>
__global__ void __launch_bounds__(128) conv1x1_f32_square_kernel(const __grid_constant__ DirectArgs a, long long pixels) {
  constexpr int WP = C + 4;                           // row pitch of the transposed weights (16-byte aligned rows, staggered banks)
  extern __shared__ __align__(16) float sq_w[];      // [C][WP] as [c][o]
  for (int i = threadIdx.x; i < C * C; i += 128) sq_w[(i % C) * WP + i / C] = __ldg(a.w + i);      // coalesced read of w[o][c], staged once per CTA
  __syncthreads();
  for (long long p = (long long)blockIdx.x * 128 + threadIdx.x; p < pixels; p += (long long)gridDim.x * 128) {
  float xin[C], acc[C];
  const float4* xp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.x) + p * a.x_ld);
#pragma unroll
  for (int i = 0; i < C / 4; ++i) {
    const float4 v = __ldg(xp + i);
    xin[4 * i] = v.x; xin[4 * i + 1] = v.y; xin[4 * i + 2] = v.z; xin[4 * i + 3] = v.w;
  }
#pragma unroll
  for (int o = 0; o < C; ++o) acc[o] = a.bias ? __ldg(a.bias + o) : 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float4* wp = reinterpret_cast<const float4*>(sq_w + c * WP);
#pragma unroll
    for (int o = 0; o < C; o += 4) {
      const float4 w4 = wp[o >> 2];
      acc[o] = fmaf(xin[c], w4.x, acc[o]); acc[o + 1] = fmaf(xin[c], w4.y, acc[o + 1]);
      acc[o + 2] = fmaf(xin[c], w4.z, acc[o + 2]); acc[o + 3] = fmaf(xin[c], w4.w, acc[o + 3]);
    }
  }
  float4* q = reinterpret_cast<float4*>(a.out_f32 + p * a.out_f32_ld);
#pragma unroll
  for (int o = 0; o < C; o += 4)
    q[o >> 2] = make_float4(act_apply(acc[o], a.act), act_apply(acc[o + 1], a.act), act_apply(acc[o + 2], a.act), act_apply(acc[o + 3], a.act));
  }
}

__global__ void __launch_bounds__(256) conv1x1_wide_bf16_kernel(const __grid_constant__ DirectArgs a, long long pixels) {
  const int oct = threadIdx.x & 7;      // output channels 8 oct .. 8 oct + 7
  float w[3][8], bv[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    bv[o] = a.bias ? __ldg(a.bias + oct * 8 + o) : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) w[c][o] = __ldg(a.w + (long long)(oct * 8 + o) * 3 + c);
  }
  const float* xin = reinterpret_cast<const float*>(a.x);
  for (long long p = ((long long)blockIdx.x * 256 + threadIdx.x) >> 3; p < pixels; p += ((long long)gridDim.x * 256) >> 3) {
    const float x0 = __ldg(xin + p * a.x_ld), x1 = __ldg(xin + p * a.x_ld + 1), x2 = __ldg(xin + p * a.x_ld + 2);
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = act_apply(fmaf(x2, w[2][o], fmaf(x1, w[1][o], fmaf(x0, w[0][o], bv[o]))), a.act);
    *reinterpret_cast<uint4*>(a.out_bf16 + p * a.out_ld + oct * 8) = pack8(acc);
  }
}

extern "C" int ff_conv_direct(const void* x, int x_is_bf16, int x_ld, int B, int H, int W, int Cin, int k, const float* w,
                              const float* bias, int Cout_pad, int n_store, int act, const float* mul_f32, int mul_ld,
                              void* out_bf16, int out_ld, float* out_f32, int out_f32_ld, void* stream) {
  FF_CHECK_ARG(x && w && (out_bf16 || out_f32), "ff_conv_direct: null buffer");
  FF_CHECK_ARG((k == 1 || k == 3) && Cin > 0 && Cin <= 192, "ff_conv_direct: k=%d Cin=%d unsupported", k, Cin);
  FF_CHECK_ARG(B > 0 && H > 0 && W > 0, "ff_conv_direct: bad size %dx%dx%d", B, H, W);
  FF_CHECK_ARG(Cout_pad % 8 == 0 && n_store <= Cout_pad, "ff_conv_direct: Cout_pad must be a multiple of 8");
  const int vb = (out_bf16 && out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0) ? 1 : 0;
  const int vf = (out_f32 && out_f32_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0) ? 1 : 0;
  DirectArgs a{x, x_is_bf16, x_ld, B, H, W, Cin, k, w, bias, Cout_pad, n_store, act, mul_f32, mul_ld, reinterpret_cast<bf16*>(out_bf16), out_ld, out_f32, out_f32_ld, vb, vf};
  cudaStream_t st_ = reinterpret_cast<cudaStream_t>(stream);
  // specialised kernels for the small fp32 layers that run at output resolution (see above)
  if (k == 3 && !x_is_bf16 && !mul_f32 && !out_bf16 && out_f32 && ((n_store == 1 && (Cin == 8 || Cin == 16)) || (n_store == 16 && Cin == 6))) {
    const int grid = B * ff_cdiv(H, 32) * ff_cdiv(W, 32);
    auto launch = [&](auto kern, int cin, int nout) -> int {
      const int smem = (((34 * cin * 35 + 3) & ~3) + 9 * cin * nout) * (int)sizeof(float);
      if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);      // (idempotent, cheap)
        if (e != cudaSuccess) { ff_set_error("ff_conv_direct: smem %d: %s", smem, cudaGetErrorString(e)); return FF_ERR_CUDA; }
      }
      kern<<<grid, 256, smem, st_>>>(a);
      return FF_OK;
    };
    int rc = FF_OK;
    if (Cin == 8) rc = launch(conv3x3_small_kernel<8, 1>, 8, 1);
    else if (Cin == 16) rc = launch(conv3x3_small_kernel<16, 1>, 16, 1);
    else rc = launch(conv3x3_small_kernel<6, 16>, 6, 16);
    if (rc != FF_OK) return rc;
    ++g_ff_launches;
    FF_CHECK_LAUNCH("ff_conv_direct");
    return FF_OK;
  }
  if (k == 1 && x_is_bf16 && Cin == 32 && n_store == 8 && !mul_f32 && !out_bf16 && vf && x_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const long long pixels = (long long)B * H * W;
    conv1x1_rows_kernel<32, 8><<<ff_cdiv(pixels, 256), 256, 0, st_>>>(a, pixels);
    ++g_ff_launches;
    FF_CHECK_LAUNCH("ff_conv_direct");
    return FF_OK;
  }
  if (k == 1 && !x_is_bf16 && Cin == 64 && n_store == 64 && !mul_f32 && !out_bf16 && vf && x_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const long long pixels = (long long)B * H * W;
    static FFPerDeviceFlag sq_configured;
    bool& conf = sq_configured.get();
    if (!conf) {
      cudaError_t e = cudaFuncSetAttribute(conv1x1_f32_square_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 68 * 4);
      if (e != cudaSuccess) { ff_set_error("ff_conv_direct: smem: %s", cudaGetErrorString(e)); return FF_ERR_CUDA; }
      conf = true;
    }
    const long long want = ff_cdiv(pixels, 128), cap = (long long)ff_num_sms() * 4;
    conv1x1_f32_square_kernel<64><<<(int)(want < cap ? want : cap), 128, 64 * 68 * 4, st_>>>(a, pixels);
    ++g_ff_launches;
    FF_CHECK_LAUNCH("ff_conv_direct");
    return FF_OK;
  }
  if (k == 1 && !x_is_bf16 && Cin == 3 && n_store == 64 && !mul_f32 && !out_f32 && vb) {
    const long long pixels = (long long)B * H * W;
    const long long want = (pixels * 8 + 255) / 256;
    const int grid = (int)(want < (long long)ff_num_sms() * 16 ? want : (long long)ff_num_sms() * 16);
    conv1x1_wide_bf16_kernel<<<grid, 256, 0, st_>>>(a, pixels);
    ++g_ff_launches;
    FF_CHECK_LAUNCH("ff_conv_direct");
    return FF_OK;
  }
  const int pad = k / 2;
  // small input-channel counts: one block computes every output-channel group (input patch and all weights staged once)
  const bool all_w = Cin <= 16 && (size_t)k * k * Cin * Cout_pad * sizeof(float) <= 96 * 1024;
  const int wgroups = all_w ? Cout_pad / 8 : 1;
  // four rows per thread when the whole problem is large and the weights are staged once per block
  const int R = (all_w && H % 32 == 0 && (long long)B * H * W >= (1 << 18)) ? 4 : 1;
  const size_t smem = ((size_t)(16 + 2 * pad) * (8 * R + 2 * pad) * (Cin | 1) + (size_t)k * k * Cin * 8 * wgroups) * sizeof(float);
  static size_t configured_dev[64][2];
  int dev_ = 0;
  cudaGetDevice(&dev_);
  size_t* configured = configured_dev[dev_ & 63];
  if (configured[0] == 0) configured[0] = configured[1] = 48 * 1024;
  if (smem > configured[R == 4]) {
    cudaError_t e = R == 4 ? cudaFuncSetAttribute(conv_direct_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                           : cudaFuncSetAttribute(conv_direct_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ff_set_error("ff_conv_direct: smem %zu: %s", smem, cudaGetErrorString(e)); return FF_ERR_CUDA; }
    configured[R == 4] = smem;
  }
  dim3 grid(B * ff_cdiv(H, 8 * R) * ff_cdiv(W, 16), all_w ? 1 : Cout_pad / 8);
  if (R == 4) conv_direct_kernel<4><<<grid, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  else conv_direct_kernel<1><<<grid, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_conv_direct");
  return FF_OK;
}

extern "C" int ff_nchw_to_nhwc(const float* x, int B, int C, int H, int W, const float* sub, float* out, int ld, void* stream) {
  FF_CHECK_ARG(x && out && ld >= C, "ff_nchw_to_nhwc: bad args");
  nchw_to_nhwc_kernel<<<ff_cdiv((long long)B * H * W, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, B, C, H, W, sub, out, ld);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_nchw_to_nhwc");
  return FF_OK;
}
extern "C" int ff_nchw_to_nhwc_pad(const float* x, int B, int C, int H, int W, const float* sub, float* out, int ld, int Hp, int Wp, int mode,
                                   void* stream) {
  FF_CHECK_ARG(x && out && ld >= C && Hp >= H && Wp >= W && (mode == 0 || mode == 1), "ff_nchw_to_nhwc_pad: bad args");
  FF_CHECK_ARG(mode == 0 || (Hp - H < H && Wp - W < W), "ff_nchw_to_nhwc_pad: reflect padding %d/%d must be smaller than the image %dx%d", Hp - H, Wp - W, H, W);
  nchw_to_nhwc_pad_kernel<<<ff_cdiv((long long)B * Hp * Wp, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, B, C, H, W, sub, out, ld, Hp, Wp, mode);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_nchw_to_nhwc_pad");
  return FF_OK;
}
extern "C" int ff_nhwc_to_nchw(const float* x, int ld, int coff, int B, int C, int H, int W, float* out, void* stream) {
  FF_CHECK_ARG(x && out && ld >= coff + C, "ff_nhwc_to_nchw: bad args");
  nhwc_to_nchw_kernel<<<ff_cdiv((long long)B * H * W, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, ld, coff, B, C, H, W, out);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_nhwc_to_nchw");
  return FF_OK;
}
