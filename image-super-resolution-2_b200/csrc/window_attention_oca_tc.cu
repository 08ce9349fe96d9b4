// Overlapping cross-attention (OCAB) of HAT on the 5th-generation tensor cores (tcgen05 + TMEM):
// hat_arch.py:392-438 -- queries = one 16x16 window, keys / values = the 24x24 window around it that nn.Unfold(kernel 24,
// stride 16, padding 4) cuts from the zero-padded k / v images (keys outside the image are all-zero rows that still take
// softmax mass, :377,408), relative-position bias through the 39x39 table that the reference indexes with NEGATIVE offsets
// (wrap-around, :896-919).
//
// One CTA = one (window, head); 256 threads; two CTAs per SM (each owns 256 of the 512 TMEM columns, 97 KB of smem).
//
//   gather   Q (256 tokens) and K, V (576 tokens) rows of ONE head: 64 B per token, 16-byte cp.async into the canonical
//            64-byte-swizzled K-major layout (512-byte atoms of 8 rows).  Zero-padded keys are zero-filled; their V row keeps
//            the all-ones column (dim 31) that accumulates the softmax row sums.
//   bias     the head's table is re-laid at load as B2[ki-qi+15][kj-qj+15] (39 x 39, row stride 48) with the reference's
//            negative-index wrap applied once, so the bias of (query, key) is an affine address again: a per-thread base plus a
//            compile-time key offset (LDS with immediate offsets, conflict-free across the 32 query rows of a warp).
//   per query half r (128 rows = TMEM lanes), keys in six chunks of 96 (4 key rows), two S buffers X / Y of 96 columns:
//     max pass   S_c = Q[r] K_c^T (tcgen05.mma M=128 N=96 K=16 x2) for c = 0..5, pipelined one chunk ahead through X / Y;
//                thread = (query row, half of the chunk's columns): tcgen05.ld, running row max of the raw logits.
//     exp pass   chunks in the order 4, 5 (still resident), 0, 1, 2, 3 (recomputed into the buffer whose P V has drained):
//                p = exp2(s + bias - shift), shift = max_k(q.k) + max(table) >= the row max (softmax is shift invariant, exp2 has
//                126 binades of headroom) -> bf16 pairs over the first half of the thread's own S columns (tcgen05.st) ->
//                O += P_c V_c (tcgen05.mma M=128 N=32 K=16 x6, A = P from TMEM, B = V from smem, MN-major) into columns [192,224).
//     out        O[:, 0..29] / O[:, 31], bf16 store at the token position.
// The logits never exist outside TMEM (the reference materialises [windows, 6, 256, 576] fp32).  Recomputing four of the six
// S chunks costs 8 small MMAs per half on an otherwise idle tensor pipe and spares a second 576-column TMEM buffer.
// q is pre-scaled by head_dim^-0.5 * log2(e) in the packed qkv weights, so the softmax is exp2.
#include "ff_common.cuh"
#include "../../include/ffb200.h"
#include <stdlib.h>

namespace {

constexpr int NQ = 256, KWIN = 24, NK = KWIN * KWIN;      // 576 keys
constexpr int ROWB = 64;                                   // bytes per token row in smem (one head: 32 dims bf16)
constexpr int CHUNK = 96, NCHUNK = NK / CHUNK;             // 6 chunks of 4 key rows
constexpr int NTHREADS = 256;
constexpr int TDIM = 39, TSTRIDE = 48;                     // re-laid bias table [39][48]
constexpr float LOG2E = 1.4426950408889634f;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t O_COL = 192;

constexpr size_t SMEM_Q = 0, SMEM_K = NQ * ROWB, SMEM_V = SMEM_K + NK * ROWB;
constexpr size_t SMEM_TAB = SMEM_V + NK * ROWB;
constexpr size_t SMEM_MAX = SMEM_TAB + TDIM * TSTRIDE * 4;      // [2][128] floats
constexpr size_t SMEM_END = SMEM_MAX + 2 * 128 * 4;
constexpr size_t SMEM_BYTES = SMEM_END + 1024;

__device__ __forceinline__ void cp_async16z(uint32_t smem_dst, const void* gsrc, int src_bytes) {
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
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

// this thread's 48 columns of one S chunk -> registers (the three loads are in flight together)
__device__ __forceinline__ void ld48(uint32_t taddr, uint32_t (&v)[48]) {
  tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
  tmem_ld16(taddr + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
  tmem_ld16(taddr + 32, *reinterpret_cast<uint32_t(*)[16]>(&v[32]));
}

__global__ void __launch_bounds__(NTHREADS, 2) ocab_attention_tc_kernel(const __grid_constant__ FFWinAttn p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_s[2];      // S chunk ready in buffer X / Y
  __shared__ __align__(8) uint64_t bar_p[2];      // P V of buffer X / Y has drained (the buffer may be overwritten)
  __shared__ __align__(8) uint64_t bar_o;         // every MMA of the unit has completed
  __shared__ uint32_t tmem_slot;
  __shared__ float sRed[NTHREADS / 32];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  float* sTab = reinterpret_cast<float*>(smem + SMEM_TAB);
  float* sMax = reinterpret_cast<float*>(smem + SMEM_MAX);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head_l = blockIdx.x % p.heads;       // heads are the fast index: the CTAs sharing a window run together
  const int head = p.head_off + head_l;
  pdl_launch_dependents();
  pdl_wait();      // the first thing this kernel does is fetch q / k / v, which the predecessor wrote
  int win = blockIdx.x / p.heads;
  const int nwx = p.W >> 4, nwy = p.H >> 4;
  const int b = win / (nwx * nwy);
  win -= b * nwx * nwy;
  const int wy = win / nwx, wx = win - wy * nwx;
  const long long img0 = (long long)b * p.H * p.W;

  if (tid == 0) {
    mbar_init(&bar_s[0], 1); mbar_init(&bar_s[1], 1);
    mbar_init(&bar_p[0], 1); mbar_init(&bar_p[1], 1);
    mbar_init(&bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }

  // ---- gather ----
  float tmax = -1e30f;
  {
    const bf16* base = reinterpret_cast<const bf16*>(p.qkv);
    // group 0: Q and K (needed by the first S), group 1: V (first needed by the first P V)
    for (int idx = tid; idx < NQ * 4; idx += NTHREADS) {
      const int t = idx >> 2, c = idx & 3;
      const int y = wy * 16 + (t >> 4), x = wx * 16 + (t & 15);
      const bf16* src = base + (img0 + (long long)y * p.W + x) * p.ld + p.q_off + head * 32 + c * 8;
      cp_async16z(sbase + (uint32_t)SMEM_Q + t * ROWB + ((c ^ ((t >> 1) & 3)) << 4), src, 16);
    }
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      for (int idx = tid; idx < NK * 4; idx += NTHREADS) {
        const int t = idx >> 2, c = idx & 3;
        const int i = t / KWIN, j = t - i * KWIN;
        const int y = wy * 16 - p.kpad_y + i, x = wx * 16 - p.kpad_x + j;
        const bool inside = y >= 0 && y < p.H && x >= 0 && x < p.W;
        const bf16* src = inside ? base + (img0 + (long long)y * p.W + x) * p.ld + head * 32 + c * 8 : base;
        const uint32_t dst = sbase + t * ROWB + ((c ^ ((t >> 1) & 3)) << 4);
        if (part == 0) {
          cp_async16z(dst + (uint32_t)SMEM_K, src + (inside ? p.k_off : 0), inside ? 16 : 0);
        } else if (inside || c != 3) {
          cp_async16z(dst + (uint32_t)SMEM_V, src + (inside ? p.v_off : 0), inside ? 16 : 0);
        } else {
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %2};" ::"r"(dst + (uint32_t)SMEM_V), "r"(0u), "r"(0x3F800000u) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // bias table of this head (x log2 e), re-laid as B2[A][Bc], A = ki - qi + 15, Bc = kj - qj + 15 in [0, 38]:
    // reference index ((ki - qi) - 7) * 39 + (kj - qj) - 7 = (A - 22) * 39 + (Bc - 22), negative values wrap by +T (hat_arch.py:896-919)
    const float* tb = p.bias_table + (long long)(p.bias_head_off + head_l) * p.T;
    for (int r = tid; r < TDIM * TDIM; r += NTHREADS) {
      const int A = r / TDIM, Bc = r - A * TDIM;
      int idx = (A - 22) * TDIM + (Bc - 22);
      idx += (idx >> 31) & p.T;
      const float v = LOG2E * __ldg(tb + idx);
      sTab[A * TSTRIDE + Bc] = v;
      tmax = fmaxf(tmax, v);
    }
    tmax = warp_max(tmax);
    if (lane == 0) sRed[warp] = tmax;
    asm volatile("cp.async.wait_group 1;" ::: "memory");      // Q and K have landed
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  {
    float m = sRed[0];
#pragma unroll
    for (int w = 1; w < NTHREADS / 32; ++w) m = fmaxf(m, sRed[w]);
    tmax = m;
  }

  const int quad = warp & 3, ch = warp >> 2;     // TMEM lane quadrant; half of a chunk's columns (2 of its 4 key rows)
  const int rih = quad * 32 + lane;              // row within the query half
  const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, CHUNK);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, 32) | (1u << 16);    // B (= V) is MN-major
  const uint64_t desc_q = umma_desc_sw64(sbase + (uint32_t)SMEM_Q);
  const uint64_t desc_k = umma_desc_sw64(sbase + (uint32_t)SMEM_K);
  const uint64_t desc_v = umma_desc_sw64(sbase + (uint32_t)SMEM_V);
  bf16* outp = reinterpret_cast<bf16*>(p.out);
  uint32_t ph_s[2] = {0, 0}, ph_p[2] = {0, 0}, ph_o = 0;

  // S_c = Q[r] K_c^T into buffer `buf` (elected thread of warp 0)
  auto issue_s = [&](int r, int c, int buf) {
    const uint64_t da = desc_q + (uint64_t)((r * 128 * ROWB) >> 4);
    const uint64_t db = desc_k + (uint64_t)((c * CHUNK * ROWB) >> 4);
    const uint32_t d = tmem_base + buf * CHUNK;
    tc_mma_bf16(d, da, db, idesc_s, 0u);
    tc_mma_bf16(d, da + 2, db + 2, idesc_s, 1u);
    tc_commit(&bar_s[buf]);
  };

#pragma unroll 1
  for (int r = 0; r < 2; ++r) {
    const int R = r * 128 + rih;                 // query token within the window
    const int qi = R >> 4, qj = R & 15;
    // bias address of (this query, key (ki, kj)) = tab_q + 4 * (ki * TSTRIDE + kj)
    const uint32_t tab_q = sbase + (uint32_t)SMEM_TAB + 4u * (uint32_t)((15 - qi) * TSTRIDE + (15 - qj));

    // ================= max pass =================
    if (warp == 0) {
      if (elect_one()) { issue_s(r, 0, 0); issue_s(r, 1, 1); }
      __syncwarp();
    }
    float mx0 = -1e30f, mx1 = -1e30f;
#pragma unroll 1
    for (int c = 0; c < NCHUNK; ++c) {
      const int buf = c & 1;
      mbar_wait(&bar_s[buf], ph_s[buf]);
      ph_s[buf] ^= 1;
      tc_fence_after();
      uint32_t raw[48];
      ld48(t_lane + buf * CHUNK + ch * 48, raw);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 48; i += 4) {
        mx0 = fmax3(mx0, __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(raw[i + 2]), __uint_as_float(raw[i + 3]));
      }
      if (c + 2 < NCHUNK) {
        tc_fence_before();
        __syncthreads();      // everybody has read this buffer
        if (warp == 0) {
          tc_fence_after();
          if (elect_one()) issue_s(r, c + 2, buf);
          __syncwarp();
        }
      }
    }
    sMax[ch * 128 + rih] = fmaxf(mx0, mx1);
    __syncthreads();
    const float mshift = fmaxf(fmaxf(mx0, mx1), sMax[(ch ^ 1) * 128 + rih]) + tmax;

    // ================= exp pass: chunks 4 (X), 5 (Y) are resident, then 0..3 are recomputed =================
#pragma unroll 1
    for (int i = 0; i < NCHUNK; ++i) {
      const int c = i < 2 ? 4 + i : i - 2;
      const int buf = i & 1;
      if (i >= 2) {
        mbar_wait(&bar_s[buf], ph_s[buf]);
        ph_s[buf] ^= 1;
        tc_fence_after();
      }
      const uint32_t t_s = t_lane + buf * CHUNK + ch * 48;
      uint32_t raw[48];
      ld48(t_s, raw);
      // keys of this thread in chunk c: key rows 4c + 2ch and 4c + 2ch + 1, all 24 columns
      const uint32_t tabp = tab_q + 4u * (uint32_t)((4 * c + 2 * ch) * TSTRIDE);
      tc_wait_ld();
      uint32_t pk[24];
#pragma unroll
      for (int k = 0; k < 48; k += 2) {
        const int kil = k / KWIN, kj = k % KWIN;      // compile-time after unrolling (pairs never straddle a key row: 24 is even)
        const float b0 = lds_f32(tabp + 4u * (uint32_t)(kil * TSTRIDE + kj)), b1 = lds_f32(tabp + 4u * (uint32_t)(kil * TSTRIDE + kj + 1));
        const float2 s2 = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(raw[k]), __uint_as_float(raw[k + 1])), make_float2(-mshift, -mshift)), make_float2(b0, b1));
        pk[k >> 1] = pack_bf16(ex2(s2.x), ex2(s2.y));
      }
      tmem_st16(t_s, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
      tmem_st8(t_s + 16, *reinterpret_cast<uint32_t(*)[8]>(&pk[16]));
      tc_wait_st();
      if (r == 0 && i == 0) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");    // V has landed (this thread's part; the barrier covers the rest)
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          // O += P_c V_c: 96 keys = 6 k-steps of 16; P of keys [0,48) of the chunk sits in columns [0,24) of the buffer, keys [48,96) in [48,72)
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const uint32_t ta = tmem_base + buf * CHUNK + (j < 3 ? j * 8 : 48 + (j - 3) * 8);
            tc_mma_bf16_ts(tmem_base + O_COL, ta, desc_v + (uint64_t)(((c * CHUNK + j * 16) * ROWB) >> 4), idesc_o, (i | j) != 0 ? 1u : 0u);
          }
          tc_commit(&bar_p[buf]);
          if (i == NCHUNK - 1) tc_commit(&bar_o);
          // the chunk that follows the next one goes into the OTHER buffer's predecessor... i.e. chunk order[i+1] needs buffer
          // (i+1)&1 = buf^1, whose P V was issued one step ago: wait for it to drain, then recompute S there
          if (i >= 1 && i + 1 < NCHUNK) {
            mbar_wait(&bar_p[buf ^ 1], ph_p[buf ^ 1]);
            ph_p[buf ^ 1] ^= 1;
            issue_s(r, i + 1 - 2, buf ^ 1);
          }
        }
        __syncwarp();
      }
    }
    // drain the phase bookkeeping of the elected thread: P V of the last two chunks
    // (every MMA of the unit has completed once bar_o flips)
    mbar_wait(&bar_o, ph_o);
    ph_o ^= 1;
    tc_fence_after();
    if (warp == 0 && elect_one()) {
      // bar_p[X] and bar_p[Y] each completed one more phase (steps 4 and 5) that nobody waited for
      ph_p[0] ^= 1;
      ph_p[1] ^= 1;
    }
    __syncwarp();

    // ---- normalise and store: warps 0-3 dims 0-15, warps 4-7 dims 16-31 ----
    {
      uint32_t o[16], os[1];
      tmem_ld16(t_lane + O_COL + ch * 16, o);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(os[0]) : "r"(t_lane + O_COL + 31) : "memory");
      tc_wait_ld();
      const float inv = 1.f / __uint_as_float(os[0]);
      const int y = wy * 16 + qi, x = wx * 16 + qj;
      bf16* dst = outp + (img0 + (long long)y * p.W + x) * p.out_ld + p.out_off + head * 32 + ch * 16;
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
    __syncthreads();      // O read out before the next half's MMAs overwrite the columns
    if (warp == 0) tc_fence_after();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

// Returns FF_OK when the tensor-core kernel was launched, 1 when the call is not HAT's overlapping cross-attention geometry
// (caller falls back), < 0 on error.
int ff_window_attention_oca_tc_try(const FFWinAttn& p, cudaStream_t st) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("FFB200_ATTN_TC");
    mode = (e && e[0] == '0') ? 0 : 1;
  }
  if (!mode) return 1;
  const bool ok = p.wh == 16 && p.ww == 16 && p.kh == 24 && p.kw == 24 && p.kpad_y == 4 && p.kpad_x == 4 && p.rel_sign == -1 && p.rel_stride == 39 &&
                  p.rel_off_y == -7 && p.rel_off_x == -7 && p.T == 1521 && p.shift_y == 0 && p.shift_x == 0 && p.heads > 0 && (p.Hp == 0 || p.Hp == p.H) &&
                  (p.Wp == 0 || p.Wp == p.W) && p.H % 16 == 0 && p.W % 16 == 0 && p.q_off % 8 == 0 && p.k_off % 8 == 0 && p.v_off % 8 == 0 && p.ld % 8 == 0 &&
                  p.out_ld % 8 == 0 && p.out_off % 8 == 0 && ((uintptr_t)p.qkv & 15) == 0 && ((uintptr_t)p.out & 15) == 0;
  if (!ok) return 1;
  static FFPerDeviceFlag configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ocab_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) {
      ff_set_error("ff_window_attention(oca tc): smem %zu: %s", (size_t)SMEM_BYTES, cudaGetErrorString(e));
      return FF_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid((unsigned)(p.B * (p.H / 16) * (p.W / 16) * p.heads));
  const cudaError_t le = ff_launch_pdl(ocab_attention_tc_kernel, grid, dim3(NTHREADS), SMEM_BYTES, st, p);
  if (le != cudaSuccess) { ff_set_error("ff_window_attention(oca tc): launch failed: %s", cudaGetErrorString(le)); return FF_ERR_CUDA; }
  FF_CHECK_LAUNCH("ff_window_attention(oca tc)");
  return FF_OK;
}
