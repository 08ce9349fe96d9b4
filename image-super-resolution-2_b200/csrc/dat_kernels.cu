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

// Tensor-core Gram matrix: G = Q^T K over a chunk of 512 tokens.  Each of the 8 warps copies its own 64 tokens (q and k rows
// of this head, 64 B each) into a private XOR-swizzled staging area with cp.async, runs 4 k-steps of mma.sync m16n8k16
// (A = Q^T and B = K both come out of [token][dim] rows through ldmatrix.trans, fp32 accumulation of exact bf16 products)
// and the squared channel norms on the side; the warps' 32x32 partials are then summed through shared memory.
constexpr int GRAM_WARP_TOK = 64;
constexpr int GRAM_WARP_BYTES = 2 * GRAM_WARP_TOK * 64;                 // q + k rows of one warp
constexpr int GRAM_SMEM = 8 * GRAM_WARP_BYTES;
__device__ __forceinline__ int gram_swz(int row, int chunk) { return row * 32 + ((chunk ^ ((row >> 1) & 3)) << 3); }   // bf16 element offset

__global__ void __launch_bounds__(256) dat_chan_gram_kernel(const bf16* __restrict__ qkv, int ld, int q_off, int k_off, int N,
                                                           int heads, int chunk_tokens, float* __restrict__ partial, int nchunks) {
  extern __shared__ __align__(16) uint8_t gram_smem[];
  const int bh = blockIdx.x, chunk = blockIdx.y;
  const int b = bh / heads, h = bh - b * heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t1 = min(N, (chunk + 1) * chunk_tokens);
  const int tw = chunk * chunk_tokens + warp * GRAM_WARP_TOK;          // first token of this warp
  bf16* sQ = reinterpret_cast<bf16*>(gram_smem + warp * GRAM_WARP_BYTES);
  bf16* sK = sQ + GRAM_WARP_TOK * 32;
  const bf16* base = qkv + ((long long)b * N) * ld + h * 32;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int idx = it * 32 + lane, r = idx >> 2, part = idx & 3;
    const bool ok = tw + r < t1;
    const bf16* row = ok ? base + (long long)(tw + r) * ld + part * 8 : base;
    const uint32_t dq = smem_u32(sQ + gram_swz(r, part)), dk = smem_u32(sK + gram_swz(r, part));
    const int nb = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dq), "l"(row + (ok ? q_off : 0)), "r"(nb) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dk), "l"(row + (ok ? k_off : 0)), "r"(nb) : "memory");
  }
  asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
  float acc[2][4][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mi][nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < GRAM_WARP_TOK / 16; ++ks) {
    uint32_t a[2][4], bq[2][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      // matrices: (dims 16mi..+7, tok 0-7), (dims +8..+15, tok 0-7), (dims ..+7, tok 8-15), (dims +8.., tok 8-15)
      const int tok = ks * 16 + (lane & 7) + ((lane >> 4) << 3);
      const uint32_t addr = smem_u32(sQ + gram_swz(tok, mi * 2 + ((lane >> 3) & 1)));
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(a[mi][0]), "=r"(a[mi][1]), "=r"(a[mi][2]), "=r"(a[mi][3]) : "r"(addr));
    }
#pragma unroll
    for (int dp = 0; dp < 2; ++dp) {
      const int tok = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const uint32_t addr = smem_u32(sK + gram_swz(tok, dp * 2 + (lane >> 4)));
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(bq[dp][0]), "=r"(bq[dp][1]), "=r"(bq[dp][2]), "=r"(bq[dp][3]) : "r"(addr));
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float (&d)[4] = acc[mi][nt];
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[mi][0]), "r"(a[mi][1]), "r"(a[mi][2]), "r"(a[mi][3]), "r"(bq[nt >> 1][(nt & 1) * 2]), "r"(bq[nt >> 1][(nt & 1) * 2 + 1]));
      }
  }
  // squared norms of dim `lane` over this warp's tokens (a 64-byte row is one wavefront whatever the swizzle)
  float nq = 0.f, nk = 0.f;
#pragma unroll 8
  for (int r = 0; r < GRAM_WARP_TOK; ++r) {
    const int off = gram_swz(r, lane >> 3) + (lane & 7);
    const float vq = __bfloat162float(sQ[off]), vk = __bfloat162float(sK[off]);
    nq = fmaf(vq, vq, nq); nk = fmaf(vk, vk, nk);
  }
  __syncwarp();
  // this warp's partial over its own staging area: [1024 G][32 |q|^2][32 |k|^2] floats (4352 B <= 8 KB)
  float* red = reinterpret_cast<float*>(gram_smem + warp * GRAM_WARP_BYTES);
  {
    const int g = lane >> 2, q2 = (lane & 3) * 2;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int i = mi * 16 + g, j = nt * 8 + q2;
        red[i * 32 + j] = acc[mi][nt][0]; red[i * 32 + j + 1] = acc[mi][nt][1];
        red[(i + 8) * 32 + j] = acc[mi][nt][2]; red[(i + 8) * 32 + j + 1] = acc[mi][nt][3];
      }
    red[1024 + lane] = nq;
    red[1056 + lane] = nk;
  }
  __syncthreads();
  float* out = partial + ((long long)bh * nchunks + chunk) * GRAM_STRIDE;
  for (int e = tid; e < GRAM_STRIDE; e += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += reinterpret_cast<const float*>(gram_smem + w * GRAM_WARP_BYTES)[e];
    out[e] = t;
  }
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
  static FFPerDeviceFlag configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(dat_chan_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM);
    if (e != cudaSuccess) { ff_set_error("ff_dat_channel_attention_weights: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return FF_ERR_CUDA; }
    configured = true;
  }
  dat_chan_gram_kernel<<<dim3(B * heads, nchunks), 256, GRAM_SMEM, st>>>(reinterpret_cast<const bf16*>(qkv), ld, q_off, k_off, N, heads, chunk, scratch, nchunks);
  dat_chan_softmax_kernel<<<B * heads, 1024, 0, st>>>(scratch, nchunks, heads, hd, temperature, reinterpret_cast<bf16*>(wout));
  g_ff_launches += 2;
  FF_CHECK_LAUNCH("ff_dat_channel_attention_weights");
  return FF_OK;
}
