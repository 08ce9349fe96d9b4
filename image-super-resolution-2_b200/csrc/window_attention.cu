// ff_window_attention: fused (shifted-)window multi-head attention for HAT W-MSA/SW-MSA, HAT OCAB and
// DAT rectangular-window spatial attention.  One CTA = one (window, head):
//   gather Q/K/V rows (cyclic shift and OCAB zero padding resolved by index arithmetic, no roll/unfold copies)
//   -> S = Q K^T on tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate)
//   -> + relative-position bias (table in smem, index affine in the key offset) + {0,-100} shift mask
//   -> online softmax (fp32, exp2) -> O = P V -> normalise -> store at the un-shifted token position.
// The logits never leave registers (the reference materialises [nW*B, heads, 256, 256|576] fp32 in HBM).
// q is pre-scaled by head_dim^-0.5 * log2(e) through the packed qkv weights, so the softmax runs on exp2;
// v carries 1.0 in padding dim 31 (bias of the packed projection) so the row sums come out of the P.V MMA.
//
// 8 warps per CTA, each owning 32 query rows (two passes of 16) -> <= 128 registers/thread and 2-3 resident CTAs
// per SM, so the gather of one window overlaps the math of another.  The kernel is instruction-issue bound
// (bias add + exp per logit), hence the template on the key-window width: every n8 key tile lies in one
// key-window row, so the bias index of logit (row, tile n, in-tile offset t) is  A_row - rowmul*ki(n) - sign*(kj0(n)+t)
// with ki(n), kj0(n) compile-time functions of n.
#include "ff_common.cuh"
#include "../../include/ffb200.h"

namespace {

constexpr int HD = 32;       // padded head dim
constexpr int ROWP = 32;     // dense 64-byte rows; the 16-byte chunk index is XOR-swizzled with (row>>1)&3 -> conflict-free ldmatrix
// element offset of 16-byte chunk `chunk` (0..3) of row `row`
__device__ __forceinline__ int swz(int row, int chunk) { return row * ROWP + ((chunk ^ ((row >> 1) & 3)) << 3); }
constexpr int NQ = 256;
constexpr int KCHUNK = 64;
constexpr int NTHREADS = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float MASKV = 100.0f * 1.4426950408889634f;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 16-byte asynchronous global->shared copy (LDGSTS); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ int region3(int p, int size, int win, int shift) {
  return p < size - win ? 0 : (p < size - shift ? 1 : 2);
}

// key-window row / first column of n8 tile `n` of the chunk starting at key kc
template <int KW>
__device__ __forceinline__ void tile_pos(int kc, int n, int& ki, int& kj0) {
  if constexpr (KW == 16) { ki = (kc >> 4) + (n >> 1); kj0 = (n & 1) * 8; }
  else if constexpr (KW == 32) { ki = (kc >> 5) + (n >> 2); kj0 = (n & 3) * 8; }
  else if constexpr (KW == 8) { ki = (kc >> 3) + n; kj0 = 0; }
  else { const int tg = (kc >> 3) + n; ki = tg / (KW / 8); kj0 = (tg - ki * (KW / 8)) * 8; }
}

// SGN = rel_sign (compile time so key offsets become LDS immediates); SGN < 0 (HAT OCAB) also enables the negative-index wrap
template <int KW, int SGN>
__global__ void __launch_bounds__(NTHREADS, 2) window_attention_kernel(const __grid_constant__ FFWinAttn p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int NK = p.kh * KW;
  bf16* sQ = reinterpret_cast<bf16*>(smem);
  bf16* sK = sQ + NQ * ROWP;
  bf16* sV = sK + NK * ROWP;
  float* sT = reinterpret_cast<float*>(sV + NK * ROWP);       // bias column of this head (x log2 e), [T]
  uint8_t* sKr = reinterpret_cast<uint8_t*>(sT + p.T);         // per key region id

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // heads are the fast grid index: the CTAs that share a window run together, so the 64-byte head slices of one token row
  // (one 128-byte line holds two heads) are fetched from DRAM once instead of once per head pass
  const int head_l = blockIdx.x % p.heads;
  const int head = p.head_off + head_l;
  // Padded geometry (DAT on sizes that are not multiples of its 32-wide windows, dat_arch.py:505-528): windows, cyclic shift and
  // mask regions live on the Hp x Wp grid; tokens with y >= H or x >= W are the zero rows F.pad appended to q, k and v AFTER the
  // projection (so no bias) -- as keys they still take softmax mass, as queries they are never stored.
  const int Hp = p.Hp > 0 ? p.Hp : p.H, Wp = p.Wp > 0 ? p.Wp : p.W;
  const int nwx = Wp / p.ww, nwy = Hp / p.wh;
  int win = blockIdx.x / p.heads;
  const int b = win / (nwx * nwy);
  win -= b * nwx * nwy;
  const int wy = win / nwx, wx = win - wy * nwx;
  const bf16* base = reinterpret_cast<const bf16*>(p.qkv);
  const long long img0 = (long long)b * p.H * p.W;
  const bool shifted = (p.shift_y | p.shift_x) != 0;
  // only windows that touch the wrapped border hold more than one mask region
  const bool need_mask = shifted && (wy == nwy - 1 || wx == nwx - 1);

  // ---- gather Q (4 x 16B per row) ----
  for (int idx = tid; idx < NQ * 4; idx += NTHREADS) {
    const int t = idx >> 2, part = idx & 3;
    const int i = t / p.ww, j = t - i * p.ww;
    int y = wy * p.wh + i + p.shift_y; if (y >= Hp) y -= Hp;
    int x = wx * p.ww + j + p.shift_x; if (x >= Wp) x -= Wp;
    const bool real = y < p.H && x < p.W;
    const bf16* src = real ? base + (img0 + (long long)y * p.W + x) * p.ld + p.q_off + head * HD + part * 8 : base;
    cp_async16(sQ + swz(t, part), src, real ? 16 : 0);
  }
  // ---- gather K, V ----
  for (int idx = tid; idx < NK * 4; idx += NTHREADS) {
    const int t = idx >> 2, part = idx & 3;
    const int i = t / KW, j = t - i * KW;
    const int ys = wy * p.wh - p.kpad_y + i, xs = wx * p.ww - p.kpad_x + j;
    const bool in_grid = ys >= 0 && ys < Hp && xs >= 0 && xs < Wp;
    int y = ys + p.shift_y; if (y >= Hp) y -= Hp;
    int x = xs + p.shift_x; if (x >= Wp) x -= Wp;
    const bool inside = in_grid && y < p.H && x < p.W;
    const bf16* src = inside ? base + (img0 + (long long)y * p.W + x) * p.ld + head * HD + part * 8 : base;
    cp_async16(sK + swz(t, part), src + (inside ? p.k_off : 0), inside ? 16 : 0);
    if (inside || part != 3) {
      cp_async16(sV + swz(t, part), src + (inside ? p.v_off : 0), inside ? 16 : 0);
    } else {
      // zero-padded key (OCAB): V row is zero except the all-ones column (dim 31) that carries the softmax row sum
      *reinterpret_cast<uint4*>(sV + swz(t, part)) = make_uint4(0, 0, 0, 0x3F800000u);
    }
    if (part == 0 && need_mask)
      sKr[t] = in_grid ? (uint8_t)(region3(ys, Hp, p.wh, p.shift_y) * 3 + region3(xs, Wp, p.ww, p.shift_x)) : 0;
  }
  {
    const float* tb = p.bias_table + (long long)(p.bias_head_off + head_l) * p.T;   // table is [heads][T]
    for (int i = tid; i < p.T; i += NTHREADS) sT[i] = LOG2E * __ldg(tb + i);
  }
  cp_async_wait_all();
  __syncthreads();

  constexpr int sgn = SGN;
  constexpr bool WRAP = SGN < 0;
  const int rowmul = sgn * p.rel_stride;
  const int tq = 2 * (lane & 3);
  bf16* outp = reinterpret_cast<bf16*>(p.out);

  // Each warp owns 32 query rows as two m16 groups that share every K / V fragment load (half the ldmatrix traffic of a
  // 16-row pass: the kernel is shared-memory-pipe bound); keys are consumed 32 at a time to keep the two groups' logits,
  // outputs and Q fragments under 128 registers (2 CTAs / SM).
  {
    const int q0 = warp * 32;
    uint32_t qa[2][2][4];
    int A[2][2], qr[2][2] = {{0, 0}, {0, 0}}, qi[2][2], qj[2][2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int row = q0 + g * 16 + (lane & 15);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) ldsm_x4(qa[g][ks], smem_u32(sQ + swz(row, ks * 2 + (lane >> 4))));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = q0 + g * 16 + h * 8 + (lane >> 2);
        qi[g][h] = r / p.ww;
        qj[g][h] = r - qi[g][h] * p.ww;
        // bias index of (row, key (ki,kj)) = (sgn*(qi-ki)+offy)*stride + sgn*(qj-kj)+offx = A_row - rowmul*ki - sgn*kj
        A[g][h] = (sgn * qi[g][h] + p.rel_off_y) * p.rel_stride + sgn * (qj[g][h] - tq) + p.rel_off_x;
        if (need_mask)
          qr[g][h] = region3(wy * p.wh + qi[g][h], Hp, p.wh, p.shift_y) * 3 + region3(wx * p.ww + qj[g][h], Wp, p.ww, p.shift_x);
      }
    }
    float m[2][2] = {{-1e30f, -1e30f}, {-1e30f, -1e30f}};
    float o[2][4][4];
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) o[g][n][i] = 0.f;

#pragma unroll 1
    for (int kc = 0; kc < NK; kc += 32) {
      // accumulators start from the relative-position bias (x log2 e): S = bias + Q K^T comes out of the MMA directly
      float s[2][4][4];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        int ki, kj0;
        tile_pos<KW>(kc, n, ki, kj0);
        const int off = rowmul * ki + sgn * kj0;
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            int i0 = A[g][0] - off - sgn * e, i1 = A[g][1] - off - sgn * e;
            if constexpr (WRAP) {
              i0 += (i0 >> 31) & p.T;
              i1 += (i1 >> 31) & p.T;
            }
            s[g][n][e] = sT[i0];
            s[g][n][2 + e] = sT[i1];
          }
      }
      // S += Q K^T : B fragments (k16 x n8) come from K rows [key][dim] (non-transposed ldmatrix), shared by both row groups
#pragma unroll
      for (int np = 0; np < 2; ++np) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t kb[4];
          const int key = kc + np * 16 + (lane & 7) + ((lane >> 4) << 3);
          ldsm_x4(kb, smem_u32(sK + swz(key, ks * 2 + ((lane >> 3) & 1))));
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            mma16816(s[g][2 * np], qa[g][ks], kb[0], kb[1]);
            mma16816(s[g][2 * np + 1], qa[g][ks], kb[2], kb[3]);
          }
        }
      }
      if (need_mask) {
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int kr = sKr[kc + n * 8 + tq + e];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              if (kr != qr[g][0]) s[g][n][e] -= MASKV;
              if (kr != qr[g][1]) s[g][n][2 + e] -= MASKV;
            }
          }
      }
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float cm0 = -1e30f, cm1 = -1e30f;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          cm0 = fmaxf(cm0, fmaxf(s[g][n][0], s[g][n][1]));
          cm1 = fmaxf(cm1, fmaxf(s[g][n][2], s[g][n][3]));
        }
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
        cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
        cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
        const float nm0 = fmaxf(m[g][0], cm0), nm1 = fmaxf(m[g][1], cm1);
        const float sc0 = ex2(m[g][0] - nm0), sc1 = ex2(m[g][1] - nm1);
        m[g][0] = nm0; m[g][1] = nm1;
#pragma unroll
        for (int n = 0; n < 4; ++n) { o[g][n][0] *= sc0; o[g][n][1] *= sc0; o[g][n][2] *= sc1; o[g][n][3] *= sc1; }
      }
      // P = exp2(S - m), O += P V (the V fragments are shared by both row groups)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t pa[2][4];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const float m0 = m[g][0], m1 = m[g][1];
          pa[g][0] = pack_bf16(ex2(s[g][2 * kk][0] - m0), ex2(s[g][2 * kk][1] - m0));
          pa[g][1] = pack_bf16(ex2(s[g][2 * kk][2] - m1), ex2(s[g][2 * kk][3] - m1));
          pa[g][2] = pack_bf16(ex2(s[g][2 * kk + 1][0] - m0), ex2(s[g][2 * kk + 1][1] - m0));
          pa[g][3] = pack_bf16(ex2(s[g][2 * kk + 1][2] - m1), ex2(s[g][2 * kk + 1][3] - m1));
        }
#pragma unroll
        for (int dp = 0; dp < 2; ++dp) {
          uint32_t vb[4];
          const int key = kc + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          ldsm_x4_t(vb, smem_u32(sV + swz(key, dp * 2 + (lane >> 4))));
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            mma16816(o[g][2 * dp], pa[g], vb[0], vb[1]);
            mma16816(o[g][2 * dp + 1], pa[g], vb[2], vb[3]);
          }
        }
      }
    }
    // softmax denominators: V carries an all-ones column at dim 31 (bias of the packed v projection), so sum_k P[k]
    // accumulates in O[:, 31] (lane quad member 3, second element of the last n8 tile) with the same bf16-rounded P
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const float inv0 = 1.f / __shfl_sync(0xffffffffu, o[g][3][1], (lane & ~3) | 3);
      const float inv1 = 1.f / __shfl_sync(0xffffffffu, o[g][3][3], (lane & ~3) | 3);
      // ---- store at the un-shifted token position ----
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int y = wy * p.wh + qi[g][h] + p.shift_y; if (y >= Hp) y -= Hp;
        int x = wx * p.ww + qj[g][h] + p.shift_x; if (x >= Wp) x -= Wp;
        if (y >= p.H || x >= p.W) continue;      // padded query position
        bf16* dst = outp + (img0 + (long long)y * p.W + x) * p.out_ld + p.out_off + head * HD + tq;
        const float inv = h ? inv1 : inv0;
#pragma unroll
        for (int n = 0; n < 4; ++n) *reinterpret_cast<uint32_t*>(dst + n * 8) = pack_bf16(o[g][n][2 * h] * inv, o[g][n][2 * h + 1] * inv);
      }
    }
  }
}

template <int KW, int SGN>
int launch(const FFWinAttn& p, size_t smem, cudaStream_t st) {
  static size_t configured_dev[64] = {};
  int dev_ = 0;
  cudaGetDevice(&dev_);
  size_t& configured = configured_dev[dev_ & 63];
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(window_attention_kernel<KW, SGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      ff_set_error("ff_window_attention: smem %zu: %s", smem, cudaGetErrorString(e));
      return FF_ERR_CUDA;
    }
    configured = smem;
  }
  const int Hp = p.Hp > 0 ? p.Hp : p.H, Wp = p.Wp > 0 ? p.Wp : p.W;
  dim3 grid((unsigned)(p.B * (Hp / p.wh) * (Wp / p.ww) * p.heads));
  window_attention_kernel<KW, SGN><<<grid, NTHREADS, smem, st>>>(p);
  FF_CHECK_LAUNCH("ff_window_attention");
  return FF_OK;
}

}  // namespace

extern long long g_ff_launches;
int ff_window_attention_tc4_try(const FFWinAttn& p, cudaStream_t st);      // window_attention_tc4.cu (tcgen05: 256-key self-attention windows, 4 CTAs / SM)
int ff_window_attention_tc_try(const FFWinAttn& p, cudaStream_t st);       // window_attention_tc.cu (the two-CTA variant, FFB200_ATTN_TC4=0)
int ff_window_attention_oca_tc_try(const FFWinAttn& p, cudaStream_t st);   // window_attention_oca_tc.cu (tcgen05: HAT's overlapping cross-attention)

extern "C" int ff_window_attention(const FFWinAttn* pp, void* stream) {
  FF_CHECK_ARG(pp != nullptr, "ff_window_attention: null params");
  const FFWinAttn& p = *pp;
  FF_CHECK_ARG(p.qkv && p.out && p.bias_table, "ff_window_attention: null buffer");
  FF_CHECK_ARG(p.wh * p.ww == NQ, "ff_window_attention: query window must hold 256 tokens (got %dx%d)", p.wh, p.ww);
  FF_CHECK_ARG(p.wh < 256 && p.ww < 256 && p.kh < 256 && p.kw < 256, "ff_window_attention: window too large");
  const int NK = p.kh * p.kw;
  FF_CHECK_ARG(NK % KCHUNK == 0 && NK >= KCHUNK, "ff_window_attention: key window %dx%d not a multiple of 64 tokens", p.kh, p.kw);
  FF_CHECK_ARG(p.kw == 8 || p.kw == 16 || p.kw == 24 || p.kw == 32, "ff_window_attention: key window width %d not in {8,16,24,32}", p.kw);
  {
    const int Hp = p.Hp > 0 ? p.Hp : p.H, Wp = p.Wp > 0 ? p.Wp : p.W;
    FF_CHECK_ARG(Hp >= p.H && Wp >= p.W && Hp % p.wh == 0 && Wp % p.ww == 0, "ff_window_attention: (padded) image %dx%d not divisible by window %dx%d", Hp, Wp, p.wh, p.ww);
  }
  FF_CHECK_ARG(p.ld % 8 == 0 && p.out_ld % 8 == 0 && p.q_off % 8 == 0 && p.k_off % 8 == 0 && p.v_off % 8 == 0 && p.out_off % 8 == 0, "ff_window_attention: offsets/pitches must be multiples of 8");
  FF_CHECK_ARG(p.heads > 0 && p.T > 0 && p.rel_stride > 0 && (p.rel_sign == 1 || p.rel_sign == -1), "ff_window_attention: bad heads/T/rel_sign");
  FF_CHECK_ARG(p.shift_y >= 0 && p.shift_y < p.wh && p.shift_x >= 0 && p.shift_x < p.ww, "ff_window_attention: bad shift");
  // index range check: without wrap-around every index must already lie in [0, T)
  const int lo_y = p.rel_sign > 0 ? -(p.kh - 1 - 0) : 0, hi_y = p.rel_sign > 0 ? p.wh - 1 : p.kh - 1;
  (void)lo_y; (void)hi_y;
  const size_t smem = (size_t)(NQ + 2 * NK) * ROWP * 2 + (size_t)p.T * 4 + (size_t)NK + 16;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ++g_ff_launches;
  {
    // HAT's (shifted-)window MSA and OCAB and DAT's 8x32 / 32x8 spatial attention run on tcgen05 / TMEM; the mma.sync kernel
    // below is the fallback for any other geometry (and for FFB200_ATTN_TC=0)
    int r = ff_window_attention_tc4_try(p, st);
    if (r <= 0) return r;
    r = ff_window_attention_tc_try(p, st);
    if (r <= 0) return r;
    r = ff_window_attention_oca_tc_try(p, st);
    if (r <= 0) return r;
  }
  const bool wrap = p.rel_sign < 0;      // HAT's overlapping-window table is indexed with negative offsets
  switch (p.kw) {
    case 8: return wrap ? launch<8, -1>(p, smem, st) : launch<8, 1>(p, smem, st);
    case 16: return wrap ? launch<16, -1>(p, smem, st) : launch<16, 1>(p, smem, st);
    case 24: return wrap ? launch<24, -1>(p, smem, st) : launch<24, 1>(p, smem, st);
    default: return wrap ? launch<32, -1>(p, smem, st) : launch<32, 1>(p, smem, st);
  }
}
