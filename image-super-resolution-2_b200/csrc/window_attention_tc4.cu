// Window attention on tcgen05 / TMEM, four CTAs per SM: the 256-key self-attention windows of HAT's (S)W-MSA (16 x 16,
// hat_arch.py:120-176, :281-306) and of DAT's spatial attention (8 x 32 / 32 x 8, dat_arch.py:290-342, 505-540).
//
// window_attention_tc.cu runs one CTA per (window, head pair) with a whole 128 x 256 logit tile in TMEM: 256 TMEM columns and 105 KB
// of shared memory per CTA cap the SM at two CTAs, whose serial phases (gather, MMA waits, read-out) leave the MUFU -- the
// unit that bounds a softmax at head dim 30 -- idle about half of the time.  This kernel makes the unit of work small enough for FOUR
// independent CTAs per SM:
//
//   one CTA = one (window, head); 128 threads (thread = query row = TMEM lane); 128 TMEM columns; 54 KB of shared memory
//   operands  Q, K, V of ONE head as 64-byte rows in the 64-byte-swizzled K-major layout: three TMA boxes [WH][WW][32 channels] when
//             the window is dense in the image, the cp.async gather for windows that wrap under the cyclic shift / hold padded tokens
//   per query half (128 rows), keys in chunks of 96 / 96 / 64 with an ONLINE softmax (the logits of a chunk live in 96 columns):
//     S_c = Q K_c^T        tcgen05.mma M128 N96|64 K16 x2                                     -> TMEM cols [0, 96)
//     pass 1               row max of the raw chunk logits -> running max m; O and the row sum are rescaled by 2^(m_old - m_new)
//                          in TMEM when a later chunk raises the maximum
//     pass 2               p = exp2(s + bias (+ mask) - m - max(table)) -> bf16 pairs in place (cols [0, 48))
//     O (+)= P_c V_c       tcgen05.mma M128 N32 K16 x6|4, A = P from TMEM, B = V (MN-major)     -> TMEM cols [96, 128)
//   out       O[:, 0..31] / O[:, 31] (V carries 1.0 in padding dim 31 -> softmax row sums), 64 contiguous bytes per thread
// The next chunk's S MMAs are issued right behind P V of the current one (the tensor pipe executes in issue order), one mbarrier
// phase per chunk.  A thread owns its whole row, so the softmax needs no cross-warp exchange.
#include "ff_common.cuh"
#include "../../include/ffb200.h"
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int NT = 256;            // tokens per window
constexpr int ROWB = 64;           // bytes per token row in smem (one head: 32 dims bf16)
constexpr int NTHREADS = 128;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float MASKV = 100.0f * 1.4426950408889634f;
constexpr uint32_t TMEM_COLS = 128;
constexpr uint32_t O_COL = 96;
constexpr int CHUNK = 96;
#ifndef FF_ATTN_POLY16
#define FF_ATTN_POLY16 0      // key pairs out of 16 whose exponentials run on the FMA pipe (0: all on the MUFU)
#endif

template <int WH_, int WW_>
struct Geo {
  static constexpr int WH = WH_, WW = WW_;
  static_assert(WH * WW == NT, "windows hold 256 tokens");
  static constexpr int TROWS = 2 * WH - 1, TCOLS = 2 * WW - 1;
  static constexpr int TSTRIDE = WW == 16 ? 48 : (WW == 32 ? 64 : 24);      // 32 consecutive query tokens of a warp hit 32 different banks
  static constexpr int LOG_WW = WW == 8 ? 3 : (WW == 16 ? 4 : 5);
  static constexpr int PIECE_ROWS = 32 / WW;          // key rows per 32-key piece
  static constexpr size_t SMEM_Q = 0, SMEM_K = NT * ROWB, SMEM_V = 2 * NT * ROWB;
  static constexpr size_t SMEM_TAB = 3 * NT * ROWB;
  static constexpr size_t SMEM_END = SMEM_TAB + TROWS * TSTRIDE * 4;
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
__device__ __forceinline__ int region3(int p, int size, int win, int shift) { return p < size - win ? 0 : (p < size - shift ? 1 : 2); }
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
// UMMA shared-memory descriptor, 64-byte swizzle: rows of 32 bf16 (= 64 B), 8-row swizzle atoms 512 B apart (SBO)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;      // SWIZZLE_64B
  return d;
}

// exp2 of two arguments on the FMA / integer pipes (no MUFU): Cody-Waite split x = n + f, f in [-0.5, 0.5], 2^f by a cubic
// (relative error < 1.1e-4, an order of magnitude below the bf16 rounding of P), the exponent added to the bit pattern.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);      // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float2 t = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
  float2 q = __ffma2_rn(f, make_float2(0.0555041087f, 0.0555041087f), make_float2(0.2402265070f, 0.2402265070f));
  q = __ffma2_rn(q, f, make_float2(0.6931471806f, 0.6931471806f));
  q = __ffma2_rn(q, f, make_float2(1.0f, 1.0f));
  return make_float2(__int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23)), __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23)));
}

// One 32-key piece of the second softmax pass: p = exp2(s + bias (+ mask) - shift) as 16 bf16 pairs.  `tabp` / `bad_y` are already
// advanced to the piece's first key row, so every key offset below is a compile-time constant (LDS immediates).  The MUFU does
// 16 exponentials per clock and SM -- the floor of a softmax at head dim 30 -- so -DFF_ATTN_POLY16=n sends n of every 16 key pairs
// through the polynomial on the FMA pipe instead.  Measured at the bench shape (W-MSA / SW-MSA / DAT 8x32 / 32x8, us):
// n = 0: 182 / 191 / 97 / 110;  4: 179 / 195 / 96 / 111;  6: 179 / 201 / 97 / 114;  8: 186 / 209 / 99 / 118 -- no gain: the kernel is
// bound by the serial latency chain of a chunk with 16 resident warps, not by the MUFU (54 % busy), so the default stays 0.
template <class G, bool MASK>
__device__ __forceinline__ void exp_piece(const uint32_t (&raw)[32], uint32_t (&pk)[16], float shift, uint32_t tabp, uint32_t bad_y, uint32_t bad_x) {
  const float2 nshift = make_float2(-shift, -shift);
#pragma unroll
  for (int c = 0; c < 32; c += 2) {
    const int kil = c >> G::LOG_WW, kj = c & (G::WW - 1);      // key row within the piece, key column
    const float2 r2 = make_float2(__uint_as_float(raw[c]), __uint_as_float(raw[c + 1]));
    const float2 b2 = make_float2(lds_f32(tabp - 4u * (uint32_t)(kil * G::TSTRIDE + kj)), lds_f32(tabp - 4u * (uint32_t)(kil * G::TSTRIDE + kj + 1)));
    float2 s2 = __fadd2_rn(__fadd2_rn(r2, nshift), b2);
    if (MASK) {
      const uint32_t eff = ((bad_y >> kil) & 1u) ? 0xFFFFFFFFu : bad_x;
      if ((eff >> kj) & 1u) s2.x -= MASKV;
      if ((eff >> (kj + 1)) & 1u) s2.y -= MASKV;
    }
    const int pr = c >> 1;
    const bool poly = FF_ATTN_POLY16 > 0 && ((pr * FF_ATTN_POLY16) / 16 != ((pr + 1) * FF_ATTN_POLY16) / 16);      // POLY16 of 16 pairs, evenly spread
    if (poly) {
      const float2 e = ex2_poly2(s2);
      pk[pr] = pack_bf16(e.x, e.y);
    } else {
      pk[pr] = pack_bf16(ex2(s2.x), ex2(s2.y));
    }
  }
}

#ifdef FF_ATTN_PROF
// development build: per-phase cycle counters of thread 0 of every CTA
__device__ unsigned long long g_attn4_prof[12];
#define PROF_DECL long long prof_t = clock64(); unsigned long long prof_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PROF(i) { const long long t_ = clock64(); prof_acc[i] += (unsigned long long)(t_ - prof_t); prof_t = t_; }
#define PROF_FLUSH if (threadIdx.x == 0) { for (int i_ = 0; i_ < 11; ++i_) atomicAdd(&g_attn4_prof[i_], prof_acc[i_]); atomicAdd(&g_attn4_prof[11], 1ull); }
#define PROF_ARGS , long long& prof_t, unsigned long long (&prof_acc)[12]
#define PROF_PASS , prof_t, prof_acc
#else
#define PROF_DECL
#define PROF(i)
#define PROF_FLUSH
#define PROF_ARGS
#define PROF_PASS
#endif

// One key chunk of one query half for this thread's row: online-softmax update of (m, O) and P_c written in place.
template <class G, bool MASK>
__device__ __forceinline__ void chunk_softmax(uint32_t t_lane, int c, float& m_run, float tmax, uint32_t tabp, uint32_t bad_y, uint32_t bad_x PROF_ARGS) {
  const int np = c == 2 ? 2 : 3;      // 32-key pieces in this chunk (96 / 96 / 64 keys)
  uint32_t raw[2][32];
  // pass 1: row max of the raw logits of the chunk (the next piece's TMEM load is in flight while a piece is reduced)
  float mx0 = -1e30f, mx1 = -1e30f;
  tmem_ld32(t_lane, raw[0]);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j < np) {
      tc_wait_ld();
      if (j + 1 < np) tmem_ld32(t_lane + (j + 1) * 32, raw[(j + 1) & 1]);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmax3(mx0, __uint_as_float(raw[j & 1][i]), __uint_as_float(raw[j & 1][i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(raw[j & 1][i + 2]), __uint_as_float(raw[j & 1][i + 3]));
      }
    }
  }
  const float m_new = fmaxf(m_run, fmaxf(mx0, mx1));
  PROF(2)
  tmem_ld32(t_lane, raw[0]);      // first piece of pass 2
  if (c > 0) {
    // a later chunk may raise the running maximum: O and the row sum (column 31) follow (P V of the previous chunk has completed:
    // the mbarrier phase this chunk waited for was committed behind it)
    const float sc = ex2(m_run - m_new);
    uint32_t (&o)[32] = raw[1];
    tmem_ld32(t_lane + O_COL, o);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
    tmem_st16(t_lane + O_COL, *reinterpret_cast<uint32_t(*)[16]>(&o[0]));
    tmem_st16(t_lane + O_COL + 16, *reinterpret_cast<uint32_t(*)[16]>(&o[16]));
  }
  PROF(3)
  m_run = m_new;
  // softmax shift = running max of q.k + max(bias table): an upper bound of the true row maximum that exceeds it by at most the
  // spread of the table (softmax is shift invariant; exp2 has 126 binades of headroom)
  const float shift = m_new + tmax;
  const int row0 = c * (CHUNK >> G::LOG_WW);      // first key row of the chunk
  // pass 2: the piece's 32 logits -> 16 bf16 pairs over columns [16 j, 16 j + 16): always inside columns this thread has already consumed
  auto piece = [&](int j, uint32_t (&cur)[32], uint32_t (&nxt)[32]) {
    tc_wait_ld();
    if (j + 1 < np) tmem_ld32(t_lane + (j + 1) * 32, nxt);
    const int prow = row0 + j * G::PIECE_ROWS;
    const uint32_t tp = tabp - 4u * (uint32_t)(prow * G::TSTRIDE);
    uint32_t pk[16];
    exp_piece<G, MASK>(cur, pk, shift, tp, MASK ? (bad_y >> prow) : 0u, bad_x);
    tmem_st16(t_lane + j * 16, pk);
  };
  piece(0, raw[0], raw[1]);
  piece(1, raw[1], raw[0]);
  if (np == 3) piece(2, raw[0], raw[1]);
  PROF(4)
  tc_wait_st();
  PROF(5)
}

template <class G>
__global__ void __launch_bounds__(NTHREADS, 4) window_attention_tc4_kernel(const __grid_constant__ FFWinAttn p, const __grid_constant__ CUtensorMap tmQKV,
                                                                           const int use_tma) {
  constexpr int WH = G::WH, WW = G::WW, TROWS = G::TROWS, TCOLS = G::TCOLS, TSTRIDE = G::TSTRIDE;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mma_bar, qk_bar, v_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float sRed[NTHREADS / 32];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  float* sTab = reinterpret_cast<float*>(smem + G::SMEM_TAB);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head_l = blockIdx.x % p.heads;       // heads are the fast index: the CTAs sharing a window run together
  const int head = p.head_off + head_l;
  pdl_launch_dependents();
  pdl_wait();      // the first thing this kernel does is fetch q / k / v, which the predecessor wrote
  // Padded geometry (DAT, dat_arch.py:505-528): windows / shift / mask regions on the Hp x Wp grid, tokens beyond H x W are
  // all-zero q / k / v rows (their V row keeps the all-ones column: as keys they still take softmax mass), never stored.
  const int Hp = p.Hp > 0 ? p.Hp : p.H, Wp = p.Wp > 0 ? p.Wp : p.W;
  int win = blockIdx.x / p.heads;
  const int nwx = Wp / WW, nwy = Hp / WH;
  const int b = win / (nwx * nwy);
  win -= b * nwx * nwy;
  const int wy = win / nwx, wx = win - wy * nwx;
  const long long img0 = (long long)b * p.H * p.W;
  const bool shifted = (p.shift_y | p.shift_x) != 0;
  const bool need_mask = shifted && (wy == nwy - 1 || wx == nwx - 1);

  // a window that neither wraps around the (padded) image under the cyclic shift nor holds padded tokens is one dense box per operand
  const int wy0 = wy * WH + p.shift_y, wx0 = wx * WW + p.shift_x;
  const bool tma_win = use_tma && wy0 + WH <= p.H && wx0 + WW <= p.W;      // CTA-uniform
  if (tid == 0) {
    mbar_init(&mma_bar, 1);
    mbar_init(&qk_bar, 1);
    mbar_init(&v_bar, 1);
    fence_mbar_init();
    if (tma_win) {
      mbar_arrive_expect_tx(&qk_bar, 2 * NT * ROWB);
      tma_load_4d(smem + G::SMEM_Q, &tmQKV, &qk_bar, p.q_off + head * 32, wx0, wy0, b);
      tma_load_4d(smem + G::SMEM_K, &tmQKV, &qk_bar, p.k_off + head * 32, wx0, wy0, b);
      mbar_arrive_expect_tx(&v_bar, NT * ROWB);
      tma_load_4d(smem + G::SMEM_V, &tmQKV, &v_bar, p.v_off + head * 32, wx0, wy0, b);
    }
  }
  PROF_DECL
  if (warp == 1) {
    tmem_alloc(&tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  float tmax = -1e30f;
  {
    if (!tma_win) {
      // gather: 3 operands x 256 tokens x 4 chunks of 16 B; Q and K in one cp.async group, V in a second one
      const bf16* base = reinterpret_cast<const bf16*>(p.qkv);
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        for (int idx = tid; idx < NT * 4; idx += NTHREADS) {
          const int t = idx >> 2, c = idx & 3;
          int y = wy * WH + (t >> G::LOG_WW) + p.shift_y; if (y >= Hp) y -= Hp;
          int x = wx * WW + (t & (WW - 1)) + p.shift_x; if (x >= Wp) x -= Wp;
          const bool real = y < p.H && x < p.W;
          const bf16* src = real ? base + (img0 + (long long)y * p.W + x) * p.ld + head * 32 + c * 8 : base;
          const uint32_t dst = sbase + t * ROWB + ((c ^ ((t >> 1) & 3)) << 4);
          if (part == 0) {
            cp_async16z(dst + (uint32_t)G::SMEM_Q, src + (real ? p.q_off : 0), real ? 16 : 0);
            cp_async16z(dst + (uint32_t)G::SMEM_K, src + (real ? p.k_off : 0), real ? 16 : 0);
          } else if (real || c != 3) {
            cp_async16z(dst + (uint32_t)G::SMEM_V, src + (real ? p.v_off : 0), real ? 16 : 0);
          } else {
            // zero V row of a padded token: dim 31 (the all-ones column that accumulates the softmax row sums) stays 1.0
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %2};" ::"r"(dst + (uint32_t)G::SMEM_V), "r"(0u), "r"(0x3F800000u) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }
    // bias table of this head, x log2(e), re-laid with row stride TSTRIDE
    const float* tb = p.bias_table + (long long)(p.bias_head_off + head_l) * p.T;
    constexpr int TITER = (TROWS * TCOLS + NTHREADS - 1) / NTHREADS;
    float tv[TITER];
#pragma unroll
    for (int i = 0; i < TITER; ++i) {      // all loads of a thread in flight together
      const int r = tid + i * NTHREADS;
      tv[i] = r < TROWS * TCOLS ? __ldg(tb + r) : -1e30f;
    }
#pragma unroll
    for (int i = 0; i < TITER; ++i) {
      const int r = tid + i * NTHREADS;
      if (r < TROWS * TCOLS) {
        const int di = r / TCOLS, dj = r - di * TCOLS;
        const float v = LOG2E * tv[i];
        sTab[di * TSTRIDE + dj] = v;
        tmax = fmaxf(tmax, v);
      }
    }
    tmax = warp_max(tmax);
    if (lane == 0) sRed[warp] = tmax;
    if (!tma_win) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");      // Q and K have landed
      fence_proxy_async_smem();      // generic/cp.async writes -> visible to the tensor core's async-proxy reads
    } else if (warp == 0) {
      mbar_wait(&qk_bar, 0);      // the warp that issues the MMAs sees the TMA transaction complete
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  tmax = fmaxf(fmaxf(sRed[0], sRed[1]), fmaxf(sRed[2], sRed[3]));
  PROF(0)

  const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);      // this thread's TMEM lane (= row of the query half)
  constexpr uint32_t idesc_s96 = umma_idesc_bf16(128, 96), idesc_s64 = umma_idesc_bf16(128, 64);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, 32) | (1u << 16);    // B (= V) is MN-major
  const uint64_t desc_q = umma_desc_sw64(sbase + (uint32_t)G::SMEM_Q);
  const uint64_t desc_k = umma_desc_sw64(sbase + (uint32_t)G::SMEM_K);
  const uint64_t desc_v = umma_desc_sw64(sbase + (uint32_t)G::SMEM_V);
  bf16* outp = reinterpret_cast<bf16*>(p.out);
  uint32_t phase = 0;

  auto issue_s = [&](int r, int c) {      // elected thread of warp 0
    const uint64_t da = desc_q + (uint64_t)((r * 128 * ROWB) >> 4);
    const uint64_t db = desc_k + (uint64_t)((c * CHUNK * ROWB) >> 4);
    const uint32_t idesc = c == 2 ? idesc_s64 : idesc_s96;
    tc_mma_bf16(tmem_base, da, db, idesc, 0u);
    tc_mma_bf16(tmem_base, da + 2, db + 2, idesc, 1u);
  };
  auto read_out = [&](int r) {      // O / row sum of query half r -> 64 bytes at the un-shifted token position
    const int R = r * 128 + tid;
    const int qi = R >> G::LOG_WW, qj = R & (WW - 1);
    uint32_t o[32];
    tmem_ld32(t_lane + O_COL, o);
    tc_wait_ld();
    const float inv = 1.f / __uint_as_float(o[31]);
    int y = wy * WH + qi + p.shift_y; if (y >= Hp) y -= Hp;
    int x = wx * WW + qj + p.shift_x; if (x >= Wp) x -= Wp;
    if (y < p.H && x < p.W) {      // (padded query positions are dropped)
      uint4* dst = reinterpret_cast<uint4*>(outp + (img0 + (long long)y * p.W + x) * p.out_ld + p.out_off + head * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        dst[q] = make_uint4(pack_bf16(__uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv),
                            pack_bf16(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv),
                            pack_bf16(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv),
                            pack_bf16(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv));
    }
  };

  if (warp == 0) {
    if (elect_one()) { issue_s(0, 0); tc_commit(&mma_bar); }
    __syncwarp();
  }
#pragma unroll 1
  for (int r = 0; r < 2; ++r) {
    const int R = r * 128 + tid;                 // query token within the window
    const int qi = R >> G::LOG_WW, qj = R & (WW - 1);
    const uint32_t tabp = sbase + (uint32_t)G::SMEM_TAB + 4u * (uint32_t)((qi + WH - 1) * TSTRIDE + (qj + WW - 1));
    uint32_t bad_x = 0, bad_y = 0;
    if (need_mask) {
      const int rqy = region3(wy * WH + qi, Hp, WH, p.shift_y), rqx = region3(wx * WW + qj, Wp, WW, p.shift_x);
#pragma unroll
      for (int k = 0; k < WH; ++k) bad_y |= (uint32_t)(region3(wy * WH + k, Hp, WH, p.shift_y) != rqy) << k;
#pragma unroll
      for (int k = 0; k < WW; ++k) bad_x |= (uint32_t)(region3(wx * WW + k, Wp, WW, p.shift_x) != rqx) << k;
    }
    float m_run = -1e30f;
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      mbar_wait(&mma_bar, phase);      // S_c is in TMEM, and every earlier MMA (P V of the previous chunk) has completed
      phase ^= 1;
      tc_fence_after();
      PROF(1)
      if (r == 1 && c == 0) { read_out(0); PROF(8) }      // O of the first half is complete; its columns are overwritten by P V below, after the barrier
      if (need_mask) chunk_softmax<G, true>(t_lane, c, m_run, tmax, tabp, bad_y, bad_x PROF_PASS);     // CTA-uniform branch
      else chunk_softmax<G, false>(t_lane, c, m_run, tmax, tabp, 0u, 0u PROF_PASS);
      if (r == 0 && c == 0) {
        if (!tma_win) {
          asm volatile("cp.async.wait_group 0;" ::: "memory");    // V has landed (this thread's part; the barrier covers the rest)
          fence_proxy_async_smem();
        } else if (warp == 0) {
          mbar_wait(&v_bar, 0);
        }
      }
      tc_fence_before();
      __syncthreads();
      PROF(6)
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          const int nk = c == 2 ? 4 : 6;      // 16-key steps of the chunk
#pragma unroll 1
          for (int j = 0; j < nk; ++j)
            tc_mma_bf16_ts(tmem_base + O_COL, tmem_base + j * 8, desc_v + (uint64_t)(((c * CHUNK + j * 16) * ROWB) >> 4), idesc_o, (c | j) != 0 ? 1u : 0u);
          // the next chunk's logits go straight behind (issue order = execution order: they overwrite P only after P V has read it)
          if (c < 2) issue_s(r, c + 1);
          else if (r == 0) issue_s(1, 0);
          tc_commit(&mma_bar);
        }
        __syncwarp();
      }
      PROF(7)
    }
  }
  mbar_wait(&mma_bar, phase);
  tc_fence_after();
  PROF(1)
  read_out(1);
  PROF(8)
  PROF_FLUSH

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int g_mode4 = -1;   // -1 unread, 0 off, 1 on
int g_tma4 = 1;

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
int launch_tc4(const FFWinAttn& p, cudaStream_t st) {
  static FFPerDeviceFlag configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(window_attention_tc4_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES);
    if (e != cudaSuccess) {
      ff_set_error("ff_window_attention(tc4): smem %zu: %s", (size_t)G::SMEM_BYTES, cudaGetErrorString(e));
      return FF_ERR_CUDA;
    }
    configured = true;
  }
  const int Hp = p.Hp > 0 ? p.Hp : p.H, Wp = p.Wp > 0 ? p.Wp : p.W;
  // one tensor map over the qkv rows: box = a whole window of one head of one operand ([WH][WW][32 channels], 64-byte rows)
  CUtensorMap tm;
  int use_tma = g_tma4;
  if (use_tma) {
    static EncodeTiledFn enc = get_encode();
    cuuint64_t dims[4] = {(cuuint64_t)p.ld, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    cuuint64_t strides[3] = {(cuuint64_t)p.ld * 2, (cuuint64_t)p.ld * 2 * p.W, (cuuint64_t)p.ld * 2 * p.W * p.H};
    cuuint32_t box[4] = {32, (cuuint32_t)G::WW, (cuuint32_t)G::WH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (!enc || p.H < G::WH || p.W < G::WW ||
        enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.qkv), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      use_tma = 0;
  }
  if (!use_tma) memset(&tm, 0, sizeof(tm));
  dim3 grid((unsigned)(p.B * (Hp / G::WH) * (Wp / G::WW) * p.heads));
  const cudaError_t le = ff_launch_pdl(window_attention_tc4_kernel<G>, grid, dim3(NTHREADS), G::SMEM_BYTES, st, p, tm, use_tma);
  if (le != cudaSuccess) { ff_set_error("ff_window_attention(tc4): launch failed: %s", cudaGetErrorString(le)); return FF_ERR_CUDA; }
  FF_CHECK_LAUNCH("ff_window_attention(tc4)");
  return FF_OK;
}

}  // namespace

// Returns FF_OK when the kernel was launched, 1 when the shape is not covered or the variant is switched off (FFB200_ATTN_TC4=0:
// the caller falls back to the two-CTA kernel of window_attention_tc.cu), < 0 on error.
int ff_window_attention_tc4_try(const FFWinAttn& p, cudaStream_t st) {
  if (g_mode4 < 0) {
    const char* e = getenv("FFB200_ATTN_TC4");
    const char* e0 = getenv("FFB200_ATTN_TC");
    g_mode4 = ((e && e[0] == '0') || (e0 && e0[0] == '0')) ? 0 : 1;
    const char* t = getenv("FFB200_ATTN_TMA");
    g_tma4 = (t && t[0] == '0') ? 0 : 1;
  }
  if (!g_mode4) return 1;
  const int Hp = p.Hp > 0 ? p.Hp : p.H, Wp = p.Wp > 0 ? p.Wp : p.W;
  // self-attention windows of 256 tokens (keys = the query window) with the standard relative-position table
  const bool shape_ok = (p.wh == 16 && p.ww == 16) || (p.wh == 8 && p.ww == 32) || (p.wh == 32 && p.ww == 8);
  const bool ok = shape_ok && p.kh == p.wh && p.kw == p.ww && p.kpad_y == 0 && p.kpad_x == 0 && p.rel_sign == 1 && p.rel_stride == 2 * p.ww - 1 &&
                  p.rel_off_y == p.wh - 1 && p.rel_off_x == p.ww - 1 && p.T == (2 * p.wh - 1) * (2 * p.ww - 1) && p.heads > 0 && Hp >= p.H && Wp >= p.W &&
                  Hp % p.wh == 0 && Wp % p.ww == 0 && p.q_off % 8 == 0 && p.k_off % 8 == 0 && p.v_off % 8 == 0 && p.ld % 8 == 0 && p.out_ld % 8 == 0 &&
                  p.out_off % 8 == 0 && p.shift_y >= 0 && p.shift_y < p.wh && p.shift_x >= 0 && p.shift_x < p.ww && ((uintptr_t)p.qkv & 15) == 0 &&
                  ((uintptr_t)p.out & 15) == 0;
  if (!ok) return 1;
  if (p.ww == 16) return launch_tc4<Geo<16, 16>>(p, st);
  if (p.ww == 32) return launch_tc4<Geo<8, 32>>(p, st);
  return launch_tc4<Geo<32, 8>>(p, st);
}

#ifdef FF_ATTN_PROF
extern "C" int ff_debug_attn4_prof(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_attn4_prof, sizeof(unsigned long long) * 12);
  if (reset) { unsigned long long z[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; cudaMemcpyToSymbol(g_attn4_prof, z, sizeof(z)); }
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, window_attention_tc4_kernel<Geo<16, 16>>, NTHREADS, Geo<16, 16>::SMEM_BYTES);
  return nb;
}
#endif
