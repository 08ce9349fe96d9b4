// DAT-specific kernels:
//  * ff_dat_aim         -- Adaptive Interaction Module tail (dat_arch.py:544-560 / :650-664): second layer of the
//                          spatial-interaction MLP (the first, C -> C/16 + BN + GELU, is an ff_conv_gemm) + both gates + sum.
//  * ff_dat_chan_gram   -- channel attention statistics (dat_arch.py:636-646): per (sample, head) Gram matrix
//                          q^T k over all tokens plus the squared L2 norms of the q / k channels, split over
//                          token chunks (deterministic two-phase reduction).
//  * ff_dat_chan_softmax -- finalise: cosine-normalise, * temperature, softmax over 30 keys; emits a per-sample
//                          block-diagonal bf16 [192 x 192] matrix so `attn @ v` runs as a tcgen05 GEMM with
//                          per-sample weights (ff_conv_gemm, w_batch_rows = 192).
#include "ff_common.cuh"
#include "../../include/ffb200.h"

extern long long g_ff_launches;

namespace {

constexpr int CP = 192;
constexpr int HID_MAX = 32;

struct AimArgs {
  const bf16* att; int att_ld;
  const bf16* conv; int conv_ld;
  const float* cgate; int cgate_ld;   // [B][192] sigmoid(channel_interaction)
  const bf16* hid; int hid_ld;        // [M][>= 32] gelu(W1 x + b1), columns >= nhid ignored
  const float* w2; float b2; int nhid;
  int mode;                           // 0: spatial block, 1: channel block
  long long M; int pixels_per_sample;
  bf16* out; int out_ld;
};

// Memory-bound gate: 8 lanes per pixel (4 pixels per warp), each lane owns 24 channels as 3 x 16-byte vectors
// (c = 64*i + 8*sub) and 4 of the 32 hidden units of the spatial-interaction MLP (whose first layer ran on the tensor cores).
__global__ void __launch_bounds__(256) dat_aim_kernel(const __grid_constant__ AimArgs a) {
  const int lane = threadIdx.x & 31, sub = lane & 7;
  float w2r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) w2r[i] = (sub * 4 + i < a.nhid) ? __ldg(a.w2 + sub * 4 + i) : 0.f;
  const long long base0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * 4;     // warp-uniform; M % 4 == 0 (checked on the host)
  const long long stride = (long long)gridDim.x * 32;
  for (long long base = base0; base < a.M; base += stride) {
    const long long p = base + (lane >> 3);
    uint4 x[3], y[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = i * 64 + sub * 8;
      x[i] = __ldg(reinterpret_cast<const uint4*>(a.att + p * a.att_ld + c));
      y[i] = __ldg(reinterpret_cast<const uint4*>(a.conv + p * a.conv_ld + c));
    }
    const uint2 hq = __ldg(reinterpret_cast<const uint2*>(a.hid + p * a.hid_ld + sub * 4));
    float s = __uint_as_float(hq.x << 16) * w2r[0] + __uint_as_float(hq.x & 0xffff0000u) * w2r[1] + __uint_as_float(hq.y << 16) * w2r[2] +
              __uint_as_float(hq.y & 0xffff0000u) * w2r[3];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    const float sg = sigmoidf_(s + a.b2);
    const int b = (int)(p / a.pixels_per_sample);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = i * 64 + sub * 8;
      const float4 m0 = __ldg(reinterpret_cast<const float4*>(a.cgate + (long long)b * a.cgate_ld + c));
      const float4 m1 = __ldg(reinterpret_cast<const float4*>(a.cgate + (long long)b * a.cgate_ld + c + 4));
      const float cg[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
      const uint32_t xw[4] = {x[i].x, x[i].y, x[i].z, x[i].w}, yw[4] = {y[i].x, y[i].y, y[i].z, y[i].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xa0 = __uint_as_float(xw[j] << 16), xa1 = __uint_as_float(xw[j] & 0xffff0000u);
        const float xc0 = __uint_as_float(yw[j] << 16), xc1 = __uint_as_float(yw[j] & 0xffff0000u);
        float r0, r1;
        if (a.mode == 0) { r0 = fmaf(xa0, cg[2 * j], sg * xc0); r1 = fmaf(xa1, cg[2 * j + 1], sg * xc1); }
        else { r0 = fmaf(xc0, cg[2 * j], sg * xa0); r1 = fmaf(xc1, cg[2 * j + 1], sg * xa1); }
        __nv_bfloat162 hh = __floats2bfloat162_rn(r0, r1);
        o[j] = *reinterpret_cast<uint32_t*>(&hh);
      }
      *reinterpret_cast<uint4*>(a.out + p * a.out_ld + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// partial[b*heads+h][chunk][0..1023] = G (i*32+j), [1024..1055] = |q_i|^2, [1056..1087] = |k_j|^2
constexpr int GRAM_STRIDE = 1088;
constexpr int GRAM_TOK = 64;

__global__ void __launch_bounds__(256) dat_chan_gram_kernel(const bf16* __restrict__ qkv, int ld, int q_off, int k_off, int N,
                                                           int heads, int chunk_tokens, float* __restrict__ partial, int nchunks) {
  __shared__ float sq[GRAM_TOK][33], sk[GRAM_TOK][33];
  const int bh = blockIdx.x, chunk = blockIdx.y;
  const int b = bh / heads, h = bh - b * heads;
  const int t0 = chunk * chunk_tokens, t1 = min(N, t0 + chunk_tokens);
  const int tid = threadIdx.x;
  const int i0 = (tid >> 4) * 2, j0 = (tid & 15) * 2;   // 2x2 sub-block of the 32x32 Gram matrix
  float g00 = 0.f, g01 = 0.f, g10 = 0.f, g11 = 0.f, nq = 0.f, nk = 0.f;
  const bf16* base = qkv + ((long long)b * N) * ld + h * 32;
  for (int t = t0; t < t1; t += GRAM_TOK) {
    // load 64 tokens x 32 dims of q and k (each thread: 8 q + 8 k values)
    {
      const int tok = tid >> 2, part = (tid & 3) * 8;
      float fq[8] = {0, 0, 0, 0, 0, 0, 0, 0}, fk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (t + tok < t1) {
        const bf16* row = base + (long long)(t + tok) * ld + part;
        const uint4 uq = *reinterpret_cast<const uint4*>(row + q_off);
        const uint4 uk = *reinterpret_cast<const uint4*>(row + k_off);
        const uint32_t wq[4] = {uq.x, uq.y, uq.z, uq.w}, wk[4] = {uk.x, uk.y, uk.z, uk.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          fq[2 * e] = __uint_as_float(wq[e] << 16); fq[2 * e + 1] = __uint_as_float(wq[e] & 0xffff0000u);
          fk[2 * e] = __uint_as_float(wk[e] << 16); fk[2 * e + 1] = __uint_as_float(wk[e] & 0xffff0000u);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { sq[tok][part + e] = fq[e]; sk[tok][part + e] = fk[e]; }
    }
    __syncthreads();
#pragma unroll 8
    for (int n = 0; n < GRAM_TOK; ++n) {
      const float a0 = sq[n][i0], a1 = sq[n][i0 + 1], b0 = sk[n][j0], b1 = sk[n][j0 + 1];
      g00 += a0 * b0; g01 += a0 * b1; g10 += a1 * b0; g11 += a1 * b1;
    }
    if (tid < 32) {
      for (int n = 0; n < GRAM_TOK; ++n) nq += sq[n][tid] * sq[n][tid];
    } else if (tid < 64) {
      for (int n = 0; n < GRAM_TOK; ++n) nk += sk[n][tid - 32] * sk[n][tid - 32];
    }
    __syncthreads();
  }
  float* out = partial + ((long long)bh * nchunks + chunk) * GRAM_STRIDE;
  out[i0 * 32 + j0] = g00; out[i0 * 32 + j0 + 1] = g01; out[(i0 + 1) * 32 + j0] = g10; out[(i0 + 1) * 32 + j0 + 1] = g11;
  if (tid < 32) out[1024 + tid] = nq;
  else if (tid < 64) out[1056 + tid - 32] = nk;
}

__global__ void __launch_bounds__(1024) dat_chan_softmax_kernel(const float* __restrict__ partial, int nchunks, int heads, int hd,
                                                               const float* __restrict__ temperature, bf16* __restrict__ wout) {
  __shared__ float G[32][33], nq[32], nk[32];
  const int bh = blockIdx.x;
  const int b = bh / heads, h = bh - b * heads;
  const int tid = threadIdx.x, i = tid >> 5, j = tid & 31;
  const float* p = partial + (long long)bh * nchunks * GRAM_STRIDE;
  float g = 0.f;
  for (int c = 0; c < nchunks; ++c) g += p[(long long)c * GRAM_STRIDE + tid];
  G[i][j] = g;
  if (tid < 64) {
    float s = 0.f;
    for (int c = 0; c < nchunks; ++c) s += p[(long long)c * GRAM_STRIDE + 1024 + tid];
    if (tid < 32) nq[tid] = fmaxf(sqrtf(s), 1e-12f); else nk[tid - 32] = fmaxf(sqrtf(s), 1e-12f);
  }
  __syncthreads();
  // row i: softmax over j < hd of G/(|q_i||k_j|) * temperature[h]
  float v = -1e30f;
  if (i < hd && j < hd) v = G[i][j] / (nq[i] * nk[j]) * temperature[h];
  const float m = warp_max(v);
  const float e = (i < hd && j < hd) ? __expf(v - m) : 0.f;
  const float s = warp_sum(e);
  const float a = (i < hd && j < hd) ? e / s : 0.f;
  // block-diagonal weight row (out channel h*32+i), column (h*32+j)
  wout[((long long)b * CP + h * 32 + i) * CP + h * 32 + j] = __float2bfloat16_rn(a);
}

}  // namespace

extern "C" int ff_dat_aim(const void* att, int att_ld, const void* conv, int conv_ld, const float* cgate, int cgate_ld,
                          const void* hid, int hid_ld, const float* w2, float b2, int nhid, int mode, long long M,
                          int pixels_per_sample, void* out, int out_ld, void* stream) {
  FF_CHECK_ARG(att && conv && cgate && hid && w2 && out, "ff_dat_aim: null buffer");
  FF_CHECK_ARG(nhid > 0 && nhid <= HID_MAX && hid_ld >= HID_MAX && hid_ld % 4 == 0, "ff_dat_aim: nhid=%d (max %d), hid_ld=%d", nhid, HID_MAX, hid_ld);
  FF_CHECK_ARG(M % 4 == 0 && att_ld % 8 == 0 && conv_ld % 8 == 0 && out_ld % 8 == 0 && cgate_ld % 4 == 0, "ff_dat_aim: M %% 4 and 16-byte rows required");
  AimArgs a{reinterpret_cast<const bf16*>(att), att_ld, reinterpret_cast<const bf16*>(conv), conv_ld, cgate, cgate_ld, reinterpret_cast<const bf16*>(hid), hid_ld,
            w2, b2, nhid, mode, M, pixels_per_sample, reinterpret_cast<bf16*>(out), out_ld};
  int grid = ff_cdiv(M, 32);
  const int cap = ff_num_sms() * 16;
  if (grid > cap) grid = cap;
  dat_aim_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_dat_aim");
  return FF_OK;
}

extern "C" int ff_dat_channel_attention_weights(const void* qkv, int ld, int q_off, int k_off, int B, int N, int heads, int hd,
                                                const float* temperature, void* wout, float* scratch, size_t scratch_bytes,
                                                void* stream) {
  FF_CHECK_ARG(qkv && temperature && wout && scratch, "ff_dat_channel_attention_weights: null buffer");
  FF_CHECK_ARG(heads * 32 == CP && hd <= 32, "ff_dat_channel_attention_weights: expects 6 heads padded to 32 dims");
  int chunk = 512;
  int nchunks = ff_cdiv(N, chunk);
  FF_CHECK_ARG(scratch_bytes >= (size_t)B * heads * nchunks * GRAM_STRIDE * sizeof(float), "ff_dat_channel_attention_weights: scratch too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dat_chan_gram_kernel<<<dim3(B * heads, nchunks), 256, 0, st>>>(reinterpret_cast<const bf16*>(qkv), ld, q_off, k_off, N, heads, chunk, scratch, nchunks);
  dat_chan_softmax_kernel<<<B * heads, 1024, 0, st>>>(scratch, nchunks, heads, hd, temperature, reinterpret_cast<bf16*>(wout));
  g_ff_launches += 2;
  FF_CHECK_LAUNCH("ff_dat_channel_attention_weights");
  return FF_OK;
}
