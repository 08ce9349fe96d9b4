// Window attention on the 5th-generation tensor cores (tcgen05 + TMEM) for the 16x16 (shifted-)window MSA of HAT
// (hat_arch.py:120-176 WindowAttention, :281-306 HAB.forward: roll -> window_partition -> attn(+mask) -> reverse -> roll).
//
// One CTA = one (window, head PAIR); 256 threads; two CTAs per SM (each owns 256 of the 512 TMEM columns) so that the
// gather / MMA-latency / read-out phases of one CTA hide under the softmax of the other.
//
//   gather   Q, K, V rows of the window (256 tokens x 2 heads x 32 dims = 128 B per token and operand) with 16-byte
//            cp.async into the canonical 128-byte-swizzled K-major layout (1024-byte atoms of 8 rows) -- the cyclic shift
//            is index arithmetic.  The 128-byte row holds both heads; a head is a 64-byte K offset of the descriptor.
//            Q and K form one cp.async group, V a second one that is only waited for before the first P V.
//   per unit (head h, query half r):
//     S  = Q[r] K^T        tcgen05.mma M=128 N=256 K=16 x2  (A, B from smem)          -> TMEM cols [0,256)
//     pass 1 (thread = query row = TMEM lane, warps 0-3 keys 0-127, warps 4-7 keys 128-255): tcgen05.ld of the raw logits,
//            row max, exchanged between the two key halves through smem
//     pass 2 p = exp2(s + bias[qi-ki, qj-kj] (+ {0,-100} shift mask, border windows only: separate instantiation) - shift)
//            with the bias table in smem (row stride 48 -> conflict-free LDS with immediate offsets), packed FADD2, and
//            shift = max_k(q.k) + max(table) >= the true row max (softmax is shift invariant; exp2 has 126 binades of
//            headroom) -> bf16 pairs -> tcgen05.st over the first half of the thread's own S columns
//     O  = P V             tcgen05.mma M=128 N=64 K=16 x16 (A = P from TMEM, B = V from smem, MN-major) -> cols [64,128)
//            N = 64 covers both heads' dims; the 32 columns of the other head are ignored (the tensor pipe is idle anyway)
//     out    O[:, h*32 .. +32] / O[:, h*32+31]   (v carries 1.0 in padding dim 31 -> softmax row sums), bf16 store at the
//            un-shifted token position
// q is pre-scaled by head_dim^-0.5 * log2(e) in the packed qkv weights, so the softmax is exp2.
// Measured alternatives (persistent CTAs, operand prefetch, two softmax groups, 16-warp single-pass team) and the phase
// timers / ablations behind this shape: tools/micro/attn_variants/README.md, DESIGN.md section 6.
#include "ff_common.cuh"
#include "../../include/ffb200.h"
#include <stdlib.h>

namespace {

constexpr int NT = 256;            // tokens per window
constexpr int ROWB = 128;          // bytes per token row in smem (2 heads x 32 dims bf16)
constexpr int TSTRIDE = 48;        // bias table row stride in smem (31 used): lanes 0-15 / 16-31 of a warp hit disjoint banks
constexpr int TROWS = 31;
constexpr int NTHREADS = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float MASKV = 100.0f * 1.4426950408889634f;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t O_COL = 64;     // O accumulator columns [64,128): S columns that are dead once P is written

constexpr size_t SMEM_Q = 0, SMEM_K = NT * ROWB, SMEM_V = 2 * NT * ROWB;
constexpr size_t SMEM_TAB = 3 * NT * ROWB;                       // 2 heads x 31 x 48 floats
constexpr size_t SMEM_MAX = SMEM_TAB + 2 * TROWS * TSTRIDE * 4;  // [2][128] floats
constexpr size_t SMEM_TMAX = SMEM_MAX + 2 * 128 * 4;             // [2] floats: max of each head's table (x log2 e)
constexpr size_t SMEM_END = SMEM_TMAX + 16;
constexpr size_t SMEM_BYTES = SMEM_END + 1024;                   // + slack for the 1024-byte alignment of the operand tiles

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
  tmem_ld16(taddr + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
}
// shared-window load with an explicit state space (the carved smem pointer is generic to the compiler); ptxas folds the
// compile-time key offset into the instruction's immediate
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ int region3(int p, int size, int win, int shift) {
  return p < size - win ? 0 : (p < size - shift ? 1 : 2);
}

// Second softmax pass of one thread (= one query row, 128 of its keys): the first 32 raw logits are already in flight into
// raw[0]; exponentials are packed to bf16 pairs and stored over the first half of the thread's own S columns.
template <bool MASK>
__device__ __forceinline__ void softmax_pass2(uint32_t t_s, uint32_t (&raw)[2][32], float mshift, uint32_t tabp, uint32_t bad_y,
                                              uint32_t bad_x) {
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
    uint32_t pk[16];
    tc_wait_ld();
    if (cb < 3) tmem_ld32(t_s + (cb + 1) * 32, raw[(cb + 1) & 1]);
    const float2 nshift = make_float2(-mshift, -mshift);
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      const int kil = cb * 2 + (c >> 4), kj = c & 15;     // key row within this half, key column (pairs: packed FADD2)
      const float2 r2 = make_float2(__uint_as_float(raw[cb & 1][c]), __uint_as_float(raw[cb & 1][c + 1]));
      const float2 b2 = make_float2(lds_f32(tabp - 4u * (uint32_t)(kil * TSTRIDE + kj)), lds_f32(tabp - 4u * (uint32_t)(kil * TSTRIDE + kj + 1)));
      float2 s2 = __fadd2_rn(__fadd2_rn(r2, nshift), b2);
      if (MASK) {
        const uint32_t eff = ((bad_y >> kil) & 1u) ? 0xFFFFu : bad_x;
        if ((eff >> kj) & 1u) s2.x -= MASKV;
        if ((eff >> (kj + 1)) & 1u) s2.y -= MASKV;
      }
      raw[cb & 1][c] = __float_as_uint(ex2(s2.x));
      raw[cb & 1][c + 1] = __float_as_uint(ex2(s2.y));
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) pk[c] = pack_bf16(__uint_as_float(raw[cb & 1][2 * c]), __uint_as_float(raw[cb & 1][2 * c + 1]));
    tmem_st16(t_s + cb * 16, pk);
  }
}

#ifdef FF_ATTN_PROF
// development build: per-phase cycle counters of thread 0 of every CTA (gather, S wait, pass 1, pass 2, PV wait, read-out)
__device__ unsigned long long g_attn_prof[8];
#define PROF_DECL long long prof_t = clock64(); unsigned long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PROF(i) { const long long t_ = clock64(); prof_acc[i] += (unsigned long long)(t_ - prof_t); prof_t = t_; }
#define PROF_FLUSH if (tid == 0) { for (int i_ = 0; i_ < 8; ++i_) atomicAdd(&g_attn_prof[i_], prof_acc[i_]); atomicAdd(&g_attn_prof[7], 1ull); }
#else
#define PROF_DECL
#define PROF(i)
#define PROF_FLUSH
#endif

__global__ void __launch_bounds__(NTHREADS, 2) window_attention_tc_kernel(const __grid_constant__ FFWinAttn p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  float* sTab = reinterpret_cast<float*>(smem + SMEM_TAB);
  float* sMax = reinterpret_cast<float*>(smem + SMEM_MAX);
  float* sTabMax = reinterpret_cast<float*>(smem + SMEM_TMAX);
  __shared__ float sRed[2][NTHREADS / 32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int npairs = p.heads >> 1;
  const int pair = blockIdx.x % npairs;          // head pairs are the fast index: the CTAs sharing a window run together
  const int head0_l = pair * 2;                  // local head index (bias table row = bias_head_off + local)
  const int head0 = p.head_off + head0_l;        // absolute head (channel block)
  int win = blockIdx.x / npairs;
  const int nwx = p.W >> 4, nwy = p.H >> 4;
  const int b = win / (nwx * nwy);
  win -= b * nwx * nwy;
  const int wy = win / nwx, wx = win - wy * nwx;
  const long long img0 = (long long)b * p.H * p.W;
  const bool shifted = (p.shift_y | p.shift_x) != 0;
  const bool need_mask = shifted && (wy == nwy - 1 || wx == nwx - 1);

  PROF_DECL
  if (tid == 0) {
    mbar_init(&mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }

  // ---- gather: 3 operands x 256 tokens x 8 chunks of 16 B ----
  {
    // two cp.async groups: Q and K (needed by the first S = Q K^T) and V (first needed by the first P V, so it travels under
    // the first tile's logits and softmax)
    const bf16* base = reinterpret_cast<const bf16*>(p.qkv);
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      for (int idx = tid; idx < NT * 8; idx += NTHREADS) {
        const int t = idx >> 3, c = idx & 7;
        int y = wy * 16 + (t >> 4) + p.shift_y; if (y >= p.H) y -= p.H;
        int x = wx * 16 + (t & 15) + p.shift_x; if (x >= p.W) x -= p.W;
        const bf16* src = base + (img0 + (long long)y * p.W + x) * p.ld + head0 * 32 + c * 8;
        const uint32_t dst = sbase + t * ROWB + ((c ^ (t & 7)) << 4);
        if (part == 0) {
          cp_async16(dst + SMEM_Q, src + p.q_off);
          cp_async16(dst + SMEM_K, src + p.k_off);
        } else {
          cp_async16(dst + SMEM_V, src + p.v_off);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // bias tables of the two heads, x log2(e), re-laid with row stride 48
    const float* tb = p.bias_table + (long long)(p.bias_head_off + head0_l) * p.T;
    float tm[2] = {-1e30f, -1e30f};
#pragma unroll
    for (int h = 0; h < 2; ++h)
      for (int r = tid; r < TROWS * TROWS; r += NTHREADS) {
        const int di = r / TROWS, dj = r - di * TROWS;
        const float v = LOG2E * __ldg(tb + h * TROWS * TROWS + r);
        sTab[h * TROWS * TSTRIDE + di * TSTRIDE + dj] = v;
        tm[h] = fmaxf(tm[h], v);
      }
    tm[0] = warp_max(tm[0]);
    tm[1] = warp_max(tm[1]);
    if (lane == 0) { sRed[0][warp] = tm[0]; sRed[1][warp] = tm[1]; }
    asm volatile("cp.async.wait_group 1;" ::: "memory");      // Q and K have landed
    fence_proxy_async_smem();      // generic/cp.async writes -> visible to the tensor core's async-proxy reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  PROF(0)
  if (tid < 2) {
    float m = sRed[tid][0];
#pragma unroll
    for (int w = 1; w < NTHREADS / 32; ++w) m = fmaxf(m, sRed[tid][w]);
    sTabMax[tid] = m;      // read after the first in-loop __syncthreads
  }

  const int quad = warp & 3, ch = warp >> 2;     // TMEM lane quadrant; key half
  const int rih = quad * 32 + lane;              // row within the query half
  const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
  const uint32_t t_s = t_lane + ch * 128;        // this thread's S columns
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 256);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64) | (1u << 16);    // B (= V) is MN-major
  const uint64_t desc_q = umma_desc_k_sw128(sbase + SMEM_Q);
  const uint64_t desc_k = umma_desc_k_sw128(sbase + SMEM_K);
  const uint64_t desc_v = umma_desc_k_sw128(sbase + SMEM_V);   // same fields: SBO = 1024 B between 8-key groups, one 64-wide MN atom
  bf16* outp = reinterpret_cast<bf16*>(p.out);
  uint32_t phase = 0;

#pragma unroll 1
  for (int unit = 0; unit < 4; ++unit) {
    const int h = unit >> 1, r = unit & 1;
    // ---- S = Q[r] K^T ----
    if (warp == 0) {
      if (elect_one()) {
        const uint64_t da = desc_q + (uint64_t)((r * 128 * ROWB + h * 64) >> 4);
        const uint64_t db = desc_k + (uint64_t)((h * 64) >> 4);
        tc_mma_bf16(tmem_base, da, db, idesc_s, 0u);
        tc_mma_bf16(tmem_base, da + 2, db + 2, idesc_s, 1u);
        tc_commit(&mma_bar);
      }
      __syncwarp();
    }
    const int R = r * 128 + rih;                 // query token within the window
    const int qi = R >> 4, qj = R & 15;
    const uint32_t tabp = sbase + (uint32_t)SMEM_TAB + 4u * (uint32_t)(h * TROWS * TSTRIDE + (qi + 15 - ch * 8) * TSTRIDE + (qj + 15));
    uint32_t bad_x = 0, bad_y = 0;
    if (need_mask) {
      const int rqy = region3(wy * 16 + qi, p.H, 16, p.shift_y), rqx = region3(wx * 16 + qj, p.W, 16, p.shift_x);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        bad_y |= (uint32_t)(region3(wy * 16 + k, p.H, 16, p.shift_y) != rqy) << k;
        bad_x |= (uint32_t)(region3(wx * 16 + k, p.W, 16, p.shift_x) != rqx) << k;
      }
      bad_y >>= ch * 8;
    }
    mbar_wait(&mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    PROF(1)

    // ---- pass 1: row max of the raw logits (the next chunk's TMEM load is in flight while a chunk is reduced) ----
    uint32_t raw[2][32];
    tmem_ld32(t_s, raw[0]);
    float mx0 = -1e30f, mx1 = -1e30f;
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      tc_wait_ld();
      if (cb < 3) tmem_ld32(t_s + (cb + 1) * 32, raw[(cb + 1) & 1]);
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        mx0 = fmax3(mx0, __uint_as_float(raw[cb & 1][c]), __uint_as_float(raw[cb & 1][c + 1]));
        mx1 = fmax3(mx1, __uint_as_float(raw[cb & 1][c + 2]), __uint_as_float(raw[cb & 1][c + 3]));
      }
    }
    tmem_ld32(t_s, raw[0]);            // first chunk of pass 2 travels under the exchange
    sMax[ch * 128 + rih] = fmaxf(mx0, mx1);
    __syncthreads();
    PROF(2)
    // softmax shift = max_k(q.k) + max(bias table): an upper bound of the true row max that exceeds it by at most the
    // spread of the table (softmax is shift invariant; exp2 has 126 binades of headroom), which spares a bias pass
    const float mshift = fmaxf(fmaxf(mx0, mx1), sMax[(ch ^ 1) * 128 + rih]) + sTabMax[h];

    // ---- pass 2: P = exp2(s + bias (+mask) - shift) as bf16 pairs over the first half of this thread's own S columns ----
    if (need_mask) softmax_pass2<true>(t_s, raw, mshift, tabp, bad_y, bad_x);     // CTA-uniform branch
    else softmax_pass2<false>(t_s, raw, mshift, tabp, 0u, 0u);
    tc_wait_st();
    PROF(3)
    if (unit == 0) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");    // V has landed (this thread's part; the barrier covers the rest)
      fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    PROF(4)

    // ---- O = P V ----
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t ta = tmem_base + (j < 8 ? j * 8 : 128 + (j - 8) * 8);
          tc_mma_bf16_ts(tmem_base + O_COL, ta, desc_v + (uint64_t)((j * 16 * ROWB) >> 4), idesc_o, j != 0 ? 1u : 0u);
        }
        tc_commit(&mma_bar);
      }
      __syncwarp();
    }
    mbar_wait(&mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    PROF(5)

    // ---- normalise and store: warps 0-3 dims 0-15, warps 4-7 dims 16-31 of head h ----
    {
      uint32_t o[16], os[1];
      tmem_ld16(t_lane + O_COL + h * 32 + ch * 16, o);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(os[0]) : "r"(t_lane + O_COL + h * 32 + 31) : "memory");
      tc_wait_ld();
      const float inv = 1.f / __uint_as_float(os[0]);
      int y = wy * 16 + qi + p.shift_y; if (y >= p.H) y -= p.H;
      int x = wx * 16 + qj + p.shift_x; if (x >= p.W) x -= p.W;
      bf16* dst = outp + (img0 + (long long)y * p.W + x) * p.out_ld + p.out_off + (head0 + h) * 32 + ch * 16;
      uint4 v0, v1;
      v0.x = pack_bf16(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
      v0.y = pack_bf16(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
      v0.z = pack_bf16(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
      v0.w = pack_bf16(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
      v1.x = pack_bf16(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
      v1.y = pack_bf16(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
      v1.z = pack_bf16(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
      v1.w = pack_bf16(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
      reinterpret_cast<uint4*>(dst)[0] = v0;
      reinterpret_cast<uint4*>(dst)[1] = v1;
    }
    tc_fence_before();
    __syncthreads();      // O read out before the next unit's S overwrites the columns
    if (warp == 0) tc_fence_after();
    PROF(6)
  }
  PROF_FLUSH

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int g_mode = -1;   // -1 unread, 0 off, 1 on

}  // namespace

#ifdef FF_ATTN_PROF
extern "C" int ff_debug_attn_prof(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_attn_prof, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; cudaMemcpyToSymbol(g_attn_prof, z, sizeof(z)); }
  return 0;
}
#endif

// Returns FF_OK when the tensor-core kernel was launched, 1 when the shape is not covered (caller falls back to the
// mma.sync kernel of window_attention.cu), < 0 on error.
int ff_window_attention_tc_try(const FFWinAttn& p, cudaStream_t st) {
  if (g_mode < 0) {
    const char* e = getenv("FFB200_ATTN_TC");
    g_mode = (e && e[0] == '0') ? 0 : 1;
  }
  if (!g_mode) return 1;
  const bool ok = (p.Hp == 0 || p.Hp == p.H) && (p.Wp == 0 || p.Wp == p.W) && p.H % 16 == 0 && p.W % 16 == 0 && p.wh == 16 && p.ww == 16 && p.kh == 16 && p.kw == 16 && p.kpad_y == 0 && p.kpad_x == 0 && p.rel_sign == 1 &&
                  p.rel_stride == 31 && p.rel_off_y == 15 && p.rel_off_x == 15 && p.T == 961 && (p.heads & 1) == 0 &&
                  (p.head_off & 1) == 0 && p.q_off % 8 == 0 && p.k_off % 8 == 0 && p.v_off % 8 == 0 && p.ld % 8 == 0 &&
                  p.out_ld % 8 == 0 && p.out_off % 8 == 0 && p.shift_y >= 0 && p.shift_y < 16 && p.shift_x >= 0 && p.shift_x < 16 &&
                  ((uintptr_t)p.qkv & 15) == 0 && ((uintptr_t)p.out & 15) == 0;
  if (!ok) return 1;
  static FFPerDeviceFlag configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(window_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) {
      ff_set_error("ff_window_attention(tc): smem %zu: %s", SMEM_BYTES, cudaGetErrorString(e));
      return FF_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid((unsigned)(p.B * (p.H / 16) * (p.W / 16) * (p.heads / 2)));
  window_attention_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(p);
  FF_CHECK_LAUNCH("ff_window_attention(tc)");
  return FF_OK;
}
