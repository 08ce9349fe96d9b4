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
#include <string.h>

namespace {

constexpr int NT = 256;            // tokens per window
constexpr int ROWB = 128;          // bytes per token row in smem (2 heads x 32 dims bf16)
constexpr int NTHREADS = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float MASKV = 100.0f * 1.4426950408889634f;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t O_COL = 64;     // O accumulator columns [64,128): S columns that are dead once P is written

// Window geometry (compile time so the key offsets of the bias reads become LDS immediates):
//   16 x 16  HAT (S)W-MSA;  8 x 32 / 32 x 8  the two branches of DAT's spatial attention (dat_arch.py:250-253).
// The relative-position table of one head is [2WH-1][2WW-1]; in smem its rows are TSTRIDE apart, chosen so that the 32 query
// rows of a warp (consecutive window tokens) hit 32 different banks for a fixed key.
template <int WH_, int WW_>
struct Geo {
  static constexpr int WH = WH_, WW = WW_;
  static_assert(WH * WW == NT, "windows hold 256 tokens");
  static constexpr int TROWS = 2 * WH - 1, TCOLS = 2 * WW - 1;
  static constexpr int TSTRIDE = WW == 16 ? 48 : (WW == 32 ? 64 : 24);
  static constexpr int LOG_WW = WW == 8 ? 3 : (WW == 16 ? 4 : 5);
  static constexpr int HALF_ROWS = 128 / WW;          // key rows in one 128-key half
  static constexpr size_t SMEM_Q = 0, SMEM_K = NT * ROWB, SMEM_V = 2 * NT * ROWB;
  static constexpr size_t SMEM_TAB = 3 * NT * ROWB;                        // 2 heads x TROWS x TSTRIDE floats
  static constexpr size_t SMEM_MAX = SMEM_TAB + 2 * TROWS * TSTRIDE * 4;   // [2][128] floats
  static constexpr size_t SMEM_TMAX = SMEM_MAX + 2 * 128 * 4;              // [2] floats: max of each head's table (x log2 e)
  static constexpr size_t SMEM_END = SMEM_TMAX + 16;
  static constexpr size_t SMEM_BYTES = SMEM_END + 1024;                    // + slack for the 1024-byte alignment of the operand tiles
};

__device__ __forceinline__ void cp_async16z(uint32_t smem_dst, const void* gsrc, int src_bytes) {      // src_bytes = 0 zero-fills
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
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
template <class G, bool MASK>
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
      const int kidx = cb * 32 + c;                                  // key within this 128-key half (pairs: packed FADD2)
      const int kil = kidx >> G::LOG_WW, kj = kidx & (G::WW - 1);     // key row within the half, key column
      const float2 r2 = make_float2(__uint_as_float(raw[cb & 1][c]), __uint_as_float(raw[cb & 1][c + 1]));
      const float2 b2 = make_float2(lds_f32(tabp - 4u * (uint32_t)(kil * G::TSTRIDE + kj)), lds_f32(tabp - 4u * (uint32_t)(kil * G::TSTRIDE + kj + 1)));
      float2 s2 = __fadd2_rn(__fadd2_rn(r2, nshift), b2);
      if (MASK) {
        const uint32_t eff = ((bad_y >> kil) & 1u) ? 0xFFFFFFFFu : bad_x;
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

template <class G>
__global__ void __launch_bounds__(NTHREADS, 2) window_attention_tc_kernel(const __grid_constant__ FFWinAttn p, const __grid_constant__ CUtensorMap tmQKV,
                                                                          const int use_tma) {
  constexpr int WH = G::WH, WW = G::WW, TROWS = G::TROWS, TCOLS = G::TCOLS, TSTRIDE = G::TSTRIDE;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mma_bar, qk_bar, v_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  float* sTab = reinterpret_cast<float*>(smem + G::SMEM_TAB);
  float* sMax = reinterpret_cast<float*>(smem + G::SMEM_MAX);
  float* sTabMax = reinterpret_cast<float*>(smem + G::SMEM_TMAX);
  __shared__ float sRed[2][NTHREADS / 32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int npairs = (p.heads + 1) >> 1;
  const int pair = blockIdx.x % npairs;          // head pairs are the fast index: the CTAs sharing a window run together
  const int head0_l = pair * 2;                  // local head index (bias table row = bias_head_off + local)
  const int head0 = p.head_off + head0_l;        // absolute head (channel block)
  const int nh = min(2, p.heads - head0_l);      // an odd head count (DAT: 3 per branch) leaves the last pair with one head
  // Padded geometry (DAT, dat_arch.py:505-528): windows / shift / mask regions on the Hp x Wp grid, tokens beyond H x W are
  // all-zero q / k / v rows (their V row keeps the all-ones column: as keys they still take softmax mass), never stored.
  const int Hp = p.Hp > 0 ? p.Hp : p.H, Wp = p.Wp > 0 ? p.Wp : p.W;
  int win = blockIdx.x / npairs;
  const int nwx = Wp / WW, nwy = Hp / WH;
  const int b = win / (nwx * nwy);
  win -= b * nwx * nwy;
  const int wy = win / nwx, wx = win - wy * nwx;
  const long long img0 = (long long)b * p.H * p.W;
  const bool shifted = (p.shift_y | p.shift_x) != 0;
  const bool need_mask = shifted && (wy == nwy - 1 || wx == nwx - 1);

  PROF_DECL
  // A window that neither wraps around the (padded) image under the cyclic shift nor holds padded tokens is one dense
  // [WH][WW][64 channels] box per operand: three TMA loads with the 128-byte swizzle land it in exactly the layout the cp.async
  // gather below builds (token = y * WW + x -> 128-byte row, 1024-byte atoms of 8 rows).  Border windows of shifted blocks, the
  // padded windows of DAT keep the gather.  (The single-head tail of an odd head count loads 64 channels as well: the second
  // half of its rows holds whatever follows in the qkv row -- never read by Q K^T, and it only feeds ignored columns of P V.)
  const int wy0 = wy * WH + p.shift_y, wx0 = wx * WW + p.shift_x;
  const bool tma_win = use_tma && wy0 + WH <= p.H && wx0 + WW <= p.W;      // CTA-uniform
  if (tid == 0) {
    mbar_init(&mma_bar, 1);
    mbar_init(&qk_bar, 1);
    mbar_init(&v_bar, 1);
    fence_mbar_init();
    if (tma_win) {
      mbar_arrive_expect_tx(&qk_bar, 2 * NT * ROWB);
      tma_load_4d(smem + G::SMEM_Q, &tmQKV, &qk_bar, p.q_off + head0 * 32, wx0, wy0, b);
      tma_load_4d(smem + G::SMEM_K, &tmQKV, &qk_bar, p.k_off + head0 * 32, wx0, wy0, b);
      mbar_arrive_expect_tx(&v_bar, NT * ROWB);
      tma_load_4d(smem + G::SMEM_V, &tmQKV, &v_bar, p.v_off + head0 * 32, wx0, wy0, b);
    }
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
      if (!tma_win)
      for (int idx = tid; idx < NT * 8; idx += NTHREADS) {
        const int t = idx >> 3, c = idx & 7;
        int y = wy * WH + (t >> G::LOG_WW) + p.shift_y; if (y >= Hp) y -= Hp;
        int x = wx * WW + (t & (WW - 1)) + p.shift_x; if (x >= Wp) x -= Wp;
        const bool real = y < p.H && x < p.W && (c < 4 || nh == 2);      // padded token, or the absent second head of an odd pair
        const bf16* src = real ? base + (img0 + (long long)y * p.W + x) * p.ld + head0 * 32 + c * 8 : base;
        const uint32_t dst = sbase + t * ROWB + ((c ^ (t & 7)) << 4);
        if (part == 0) {
          cp_async16z(dst + G::SMEM_Q, src + (real ? p.q_off : 0), real ? 16 : 0);
          cp_async16z(dst + G::SMEM_K, src + (real ? p.k_off : 0), real ? 16 : 0);
        } else if (real || (c & 3) != 3) {
          cp_async16z(dst + G::SMEM_V, src + (real ? p.v_off : 0), real ? 16 : 0);
        } else {
          // zero V row of a padded token: dim 31 (the all-ones column that accumulates the softmax row sums) stays 1.0
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %2};" ::"r"(dst + (uint32_t)G::SMEM_V), "r"(0u), "r"(0x3F800000u) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // bias tables of the two heads, x log2(e), re-laid with row stride TSTRIDE
    const float* tb = p.bias_table + (long long)(p.bias_head_off + head0_l) * p.T;
    float tm[2] = {-1e30f, -1e30f};
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (h < nh)
        for (int r = tid; r < TROWS * TCOLS; r += NTHREADS) {
          const int di = r / TCOLS, dj = r - di * TCOLS;
          const float v = LOG2E * __ldg(tb + h * TROWS * TCOLS + r);
          sTab[h * TROWS * TSTRIDE + di * TSTRIDE + dj] = v;
          tm[h] = fmaxf(tm[h], v);
        }
    tm[0] = warp_max(tm[0]);
    tm[1] = warp_max(tm[1]);
    if (lane == 0) { sRed[0][warp] = tm[0]; sRed[1][warp] = tm[1]; }
    asm volatile("cp.async.wait_group 1;" ::: "memory");      // Q and K have landed
    fence_proxy_async_smem();      // generic/cp.async writes -> visible to the tensor core's async-proxy reads
    if (tma_win && warp == 0) mbar_wait(&qk_bar, 0);      // the warp that issues the MMAs sees the TMA transaction complete
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
  const uint64_t desc_q = umma_desc_k_sw128(sbase + G::SMEM_Q);
  const uint64_t desc_k = umma_desc_k_sw128(sbase + G::SMEM_K);
  const uint64_t desc_v = umma_desc_k_sw128(sbase + G::SMEM_V);   // same fields: SBO = 1024 B between 8-key groups, one 64-wide MN atom
  bf16* outp = reinterpret_cast<bf16*>(p.out);
  uint32_t phase = 0;
  const int nunits = 2 * nh;

#pragma unroll 1
  for (int unit = 0; unit < nunits; ++unit) {
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
    const int qi = R >> G::LOG_WW, qj = R & (WW - 1);
    const uint32_t tabp = sbase + (uint32_t)G::SMEM_TAB + 4u * (uint32_t)(h * TROWS * TSTRIDE + (qi + WH - 1 - ch * G::HALF_ROWS) * TSTRIDE + (qj + WW - 1));
    uint32_t bad_x = 0, bad_y = 0;
    if (need_mask) {
      const int rqy = region3(wy * WH + qi, Hp, WH, p.shift_y), rqx = region3(wx * WW + qj, Wp, WW, p.shift_x);
#pragma unroll
      for (int k = 0; k < WH; ++k) bad_y |= (uint32_t)(region3(wy * WH + k, Hp, WH, p.shift_y) != rqy) << k;
#pragma unroll
      for (int k = 0; k < WW; ++k) bad_x |= (uint32_t)(region3(wx * WW + k, Wp, WW, p.shift_x) != rqx) << k;
      bad_y >>= ch * G::HALF_ROWS;
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
    if (need_mask) softmax_pass2<G, true>(t_s, raw, mshift, tabp, bad_y, bad_x);     // CTA-uniform branch
    else softmax_pass2<G, false>(t_s, raw, mshift, tabp, 0u, 0u);
    tc_wait_st();
    PROF(3)
    if (unit == 0) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");    // V has landed (this thread's part; the barrier covers the rest)
      fence_proxy_async_smem();
      if (tma_win && warp == 0) mbar_wait(&v_bar, 0);
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
      int y = wy * WH + qi + p.shift_y; if (y >= Hp) y -= Hp;
      int x = wx * WW + qj + p.shift_x; if (x >= Wp) x -= Wp;
      if (y < p.H && x < p.W) {      // (padded query positions are dropped)
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
int g_tma = 1;     // FFB200_ATTN_TMA=0: every window through the cp.async gather

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    return reinterpret_cast<EncodeTiledFn>(ptr);
  return nullptr;
}

template <class G>
int launch_tc_attn(const FFWinAttn& p, cudaStream_t st) {
  static FFPerDeviceFlag configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(window_attention_tc_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES);
    if (e != cudaSuccess) {
      ff_set_error("ff_window_attention(tc): smem %zu: %s", (size_t)G::SMEM_BYTES, cudaGetErrorString(e));
      return FF_ERR_CUDA;
    }
    configured = true;
  }
  const int Hp = p.Hp > 0 ? p.Hp : p.H, Wp = p.Wp > 0 ? p.Wp : p.W;
  // one tensor map over the qkv rows: box = a whole window of one head pair of one operand ([WH][WW][64 channels])
  CUtensorMap tm;
  int use_tma = g_tma;
  if (use_tma) {
    static EncodeTiledFn enc = get_encode();
    cuuint64_t dims[4] = {(cuuint64_t)p.ld, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    cuuint64_t strides[3] = {(cuuint64_t)p.ld * 2, (cuuint64_t)p.ld * 2 * p.W, (cuuint64_t)p.ld * 2 * p.W * p.H};
    cuuint32_t box[4] = {64, (cuuint32_t)G::WW, (cuuint32_t)G::WH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (!enc || p.H < G::WH || p.W < G::WW ||
        enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.qkv), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      use_tma = 0;
  }
  if (!use_tma) memset(&tm, 0, sizeof(tm));
  dim3 grid((unsigned)(p.B * (Hp / G::WH) * (Wp / G::WW) * ((p.heads + 1) / 2)));
  window_attention_tc_kernel<G><<<grid, NTHREADS, G::SMEM_BYTES, st>>>(p, tm, use_tma);
  FF_CHECK_LAUNCH("ff_window_attention(tc)");
  return FF_OK;
}

}  // namespace

// Returns FF_OK when the tensor-core kernel was launched, 1 when the shape is not covered (caller falls back to the
// mma.sync kernel of window_attention.cu), < 0 on error.
int ff_window_attention_tc_try(const FFWinAttn& p, cudaStream_t st) {
  if (g_mode < 0) {
    const char* e = getenv("FFB200_ATTN_TC");
    g_mode = (e && e[0] == '0') ? 0 : 1;
    const char* t = getenv("FFB200_ATTN_TMA");
    g_tma = (t && t[0] == '0') ? 0 : 1;
  }
  if (!g_mode) return 1;
  const int Hp = p.Hp > 0 ? p.Hp : p.H, Wp = p.Wp > 0 ? p.Wp : p.W;
  // self-attention windows of 256 tokens (keys = the query window) with the standard relative-position table
  const bool shape_ok = (p.wh == 16 && p.ww == 16) || (p.wh == 8 && p.ww == 32) || (p.wh == 32 && p.ww == 8);
  const bool ok = shape_ok && p.kh == p.wh && p.kw == p.ww && p.kpad_y == 0 && p.kpad_x == 0 && p.rel_sign == 1 && p.rel_stride == 2 * p.ww - 1 &&
                  p.rel_off_y == p.wh - 1 && p.rel_off_x == p.ww - 1 && p.T == (2 * p.wh - 1) * (2 * p.ww - 1) && p.heads > 0 && Hp >= p.H && Wp >= p.W &&
                  Hp % p.wh == 0 && Wp % p.ww == 0 && p.q_off % 8 == 0 && p.k_off % 8 == 0 && p.v_off % 8 == 0 && p.ld % 8 == 0 && p.out_ld % 8 == 0 &&
                  p.out_off % 8 == 0 && p.shift_y >= 0 && p.shift_y < p.wh && p.shift_x >= 0 && p.shift_x < p.ww && ((uintptr_t)p.qkv & 15) == 0 &&
                  ((uintptr_t)p.out & 15) == 0;
  if (!ok) return 1;
  if (p.ww == 16) return launch_tc_attn<Geo<16, 16>>(p, st);
  if (p.ww == 32) return launch_tc_attn<Geo<8, 32>>(p, st);
  return launch_tc_attn<Geo<32, 8>>(p, st);
}

#ifdef FF_ATTN_PROF
extern "C" int ff_debug_attn_prof(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_attn_prof, sizeof(unsigned long long) * 8);
  if (reset) { unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; cudaMemcpyToSymbol(g_attn_prof, z, sizeof(z)); }
  return 0;
}
#endif
