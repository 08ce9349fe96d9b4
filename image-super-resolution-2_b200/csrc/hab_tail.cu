// ff_hab_tail: everything of a HAT block that follows the attention, as ONE kernel for sm_100a (hat_arch.py:303-309, :435-438):
//
//     x1 = shortcut + proj(attn) (+ 0.01 * cab * se)          x2 = x1 + fc2(GELU(fc1(LayerNorm2(x1))))        t' = LayerNorm'(x2)
//
// The two-kernel form (ff_conv_gemm with the residual + LayerNorm epilogue, then ff_mlp_fused) writes x1 (fp32) and LayerNorm2(x1)
// (bf16) to HBM and reads both back: 603 MB per block at the bench shape, 84 blocks per step.  Here x1 lives in TMEM for the whole
// block -- it is the *initial value of the fc2 accumulator* -- and LayerNorm2(x1) is written straight into shared memory as the
// K-major A operand of fc1.
//
//   per 128-token tile (8 x 16 pixels), accumulator buffer Y = acc2[tile & 1] (192 TMEM columns):
//   G0     Y  = [attn | cab] . [W_proj ; diag(0.01 se_b)]^T     tcgen05.mma M128 N192 K16 x12 (+ M128 N64 K16 x12 for the diagonal
//                                                               block: only its 64 x 64 diagonal slabs are multiplied), operands
//                                                               streamed in k-blocks through the weight ring
//   MID    8 warps: pass 1  x1 = Y + b_proj + shortcut (fp32 residual sub-blocks TMA-loaded two ahead), row sums, x1 written back
//                           over Y (tcgen05.st);  pass 2  LayerNorm2(x1) -> bf16 -> A tile in smem (128B-swizzled K-major)
//   G1/G2  the MLP chain: G1(c) = A . W1c^T (N64, acc1[c & 1]) -> 8 GELU warps: + b1, tanh-form GELU, bf16 pairs written back IN PLACE
//          over the chunk's own TMEM columns (tcgen05.st) -> G2(c): Y += H_c . W2c^T with A = H_c from TMEM (as P in the attention
//          kernels): the hidden tile has no shared-memory round trip; acc1[c & 1] is reused by G1(c + 2) behind G2(c) in the tensor
//          pipe's issue order, so it needs no "empty" barrier
//   FIN    8 warps: x2 = Y + b2 -> fp32 TMA store (+ optional bf16 copy), row sums, written back, LayerNorm' -> bf16 TMA store
// Persistent CTAs (one per SM), 26 warps: TMA producer, MMA issuer, 8 GELU, 8 MID, 8 FIN (two per TMEM lane quadrant each).  G0 of
// tile i+1 is issued early in the MLP chain of tile i, so MID(i+1) runs under the chain of tile i and FIN(i) under the chain of tile
// i+1 (Y is double buffered: 2 x 64 + 2 x 192 = 512 TMEM columns).  Per-channel vectors (biases, LayerNorm weights) are staged in
// shared memory once per CTA; epilogue TMEM loads run one 16-column sub-block ahead of the math.
// Measured (bench shape, 262 144 tokens): 219 us with the cab term / 198 us without, against 260 / 252 us for the two-kernel form.
// The kernel is shared-memory-bandwidth bound (tools/tail_phases.py on a -DFF_TAIL_PROF build): ~2 MB of smem traffic per tile --
// MMA operand reads 0.72 MB (single-CTA M128 tiles re-read the A tile for each of the six hidden chunks), the weight ring 0.49 MB
// written + read again by the MMAs, epilogue staging 0.5 MB -- at ~75 B/clk; moving H from smem to TMEM bought 12 us.
#include "ff_common.cuh"
#include "../../include/ffb200.h"
#include <stdlib.h>

extern long long g_ff_launches;

namespace {

constexpr int TM = 128, TW_ = 16, TH_ = 8;      // tile: 8 rows x 16 pixels
constexpr int CP = 192, HID = 384;
constexpr int CH = 64, NCH = HID / CH;           // hidden chunks
constexpr int KB = 64;                            // bf16 per k-block = 128 B
constexpr int KBLK = TM * KB * 2;                 // 16 KB: one 128 x 64 bf16 operand k-block
constexpr int A_BYTES = 3 * KBLK;                 // LayerNorm2(x1) as the A operand of fc1
constexpr int WSTAGE = 24 * 1024, NST = 4;        // ring stage: a W1 chunk, a W2 / W_proj k-block, an attn k-block, or cab k-block + diagonal slab
constexpr int SUB = 16, NSUB = CP / SUB;          // epilogue sub-blocks of 16 columns; the two warps of a quadrant alternate over them
constexpr int SUB_PER_WARP = NSUB / 2;
constexpr int EPI_WARPS = 8;
constexpr int MID_WARP_BYTES = 2 * 2048;          // two residual sub-blocks [32 rows][16 fp32] (64B swizzle)
constexpr int FIN_WARP_BYTES = 2048 + 1024;       // one fp32 staging sub-block (= two bf16 ones for the LayerNorm pass) + one bf16 sub-block [32 rows][16 bf16] (32B swizzle)
constexpr int OFF_A = 0, OFF_W = OFF_A + A_BYTES, OFF_MID = OFF_W + NST * WSTAGE;
constexpr int OFF_FIN = OFF_MID + EPI_WARPS * MID_WARP_BYTES;
// per-channel vectors staged once per CTA (floats): every epilogue thread needs all of them for every tile
constexpr int P_BP = 0, P_G2 = 192, P_BE2 = 384, P_B2 = 576, P_LNG = 768, P_LNB = 960, P_B1 = 1152, P_FLOATS = 1536;
constexpr int OFF_P = OFF_FIN + EPI_WARPS * FIN_WARP_BYTES;
constexpr int SMEM_BYTES = OFF_P + P_FLOATS * 4 + 1024;
constexpr int W_GELU0 = 2, W_MID0 = 10, W_FIN0 = 18;
constexpr int NTHREADS = 32 * (W_FIN0 + EPI_WARPS);      // 832
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t ACC1_COL = 0, ACC2_COL = 128;

struct Args {
  int B, H, W;
  int tiles_x, tiles_per_img, m_tiles;
  int has_cab, wp_batch_rows;
  const float* diag;      // [B][diag_ld] per-sample channel scale of the a1 term (the diagonal K block is generated on chip), or null
  int diag_ld;
  float diag_alpha;
  const float* bp;        // [192] proj bias
  const float* g2;        // LayerNorm2 gamma / beta, zero-padded to 192
  const float* be2;
  const float* b1;        // [384]
  const float* b2;        // [192]
  const float* ln_gamma;  // next LayerNorm, or null
  const float* ln_beta;
  float ln_eps;
  int ln_cols;
  int has_bf16, has_ln;
  int g0_pos;             // G0 of the next tile is issued after G2(g0_pos) / G1(g0_pos + 2) of the current one
};

#ifdef FF_TAIL_PROF
// development build: cycle counters of block 0 (lane 0 of the MMA warp, of MID warp 0 and of FIN warp 0), summed over its tiles
__device__ unsigned long long g_tail_prof[32];
#define TPROF_T0 long long tp_t = clock64();
#define TPROF(i) { const long long t_ = clock64(); if (blockIdx.x == 0 && lane == 0) atomicAdd(&g_tail_prof[i], (unsigned long long)(t_ - tp_t)); tp_t = t_; }
#else
#define TPROF_T0
#define TPROF(i) {}
#endif

__device__ __forceinline__ float gelu_tanh_hw(float x) {
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

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

struct TailMaps {
  CUtensorMap A0, A1, Wp, WpD, W1, W2, R, O32, O16, LN;
};

__global__ void __launch_bounds__(NTHREADS, 1) hab_tail_kernel(const __grid_constant__ TailMaps tm, const __grid_constant__ Args a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_full[NST], w_empty[NST];
  __shared__ __align__(8) uint64_t g0_full[2], a_full, a_empty;
  __shared__ __align__(8) uint64_t acc1_full[2], h_full[2], acc2_full[2], acc2_empty[2];
  __shared__ __align__(8) uint64_t res_bar[EPI_WARPS][2];
  __shared__ float2 ln_part[2][TM];      // per-row exchange slots of the MID / FIN warp pairs
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = a.m_tiles;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.A0); tma_prefetch_desc(&tm.Wp); tma_prefetch_desc(&tm.W1); tma_prefetch_desc(&tm.W2);
    if (a.has_cab) { tma_prefetch_desc(&tm.A1); tma_prefetch_desc(&tm.WpD); }
    for (int s = 0; s < NST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&g0_full[s], 1);
      mbar_init(&acc1_full[s], 1); mbar_init(&h_full[s], EPI_WARPS);
      mbar_init(&acc2_full[s], 1); mbar_init(&acc2_empty[s], EPI_WARPS);
    }
    mbar_init(&a_full, EPI_WARPS); mbar_init(&a_empty, 1);
    for (int w = 0; w < EPI_WARPS; ++w) { mbar_init(&res_bar[w][0], 1); mbar_init(&res_bar[w][1], 1); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  float* sP = reinterpret_cast<float*>(smem + OFF_P);
  for (int i = threadIdx.x; i < P_FLOATS; i += NTHREADS) {
    float v;
    if (i < P_G2) v = a.bp[i];
    else if (i < P_BE2) v = a.g2[i - P_G2];
    else if (i < P_B2) v = a.be2[i - P_BE2];
    else if (i < P_LNG) v = a.b2[i - P_B2];
    else if (i < P_LNB) v = a.has_ln ? a.ln_gamma[i - P_LNG] : 0.f;
    else if (i < P_B1) v = a.has_ln ? a.ln_beta[i - P_LNB] : 0.f;
    else v = a.b1[i - P_B1];
    sP[i] = v;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();      // the prologue above reads static parameters only; the predecessor's tensors are touched from here on
  const uint32_t tmem_base = tmem_slot;

  auto tile_coords = [&](int tile, int& b, int& y0, int& x0) {
    b = tile / a.tiles_per_img;
    const int t = tile - b * a.tiles_per_img;
    const int ty = t / a.tiles_x;
    y0 = ty * TH_;
    x0 = (t - ty * a.tiles_x) * TW_;
  };

  if (warp == 0) {
    // ================= TMA producer: ring stages in the MMA warp's consumption order =================
    int st = 0;
    uint32_t st_phase = 0;
    auto stage_begin = [&]() -> uint8_t* {      // all lanes wait; the caller's elected lane fills
      mbar_wait(&w_empty[st], st_phase ^ 1);
      return smem + OFF_W + st * WSTAGE;
    };
    auto stage_end = [&]() {
      __syncwarp();
      if (++st == NST) { st = 0; st_phase ^= 1; }
    };
    auto load_g0 = [&](int tile) {
      int b, y0, x0;
      tile_coords(tile, b, y0, x0);
      const int wrow = b * a.wp_batch_rows;
      float dv[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};      // this lane's diagonal entries (rows lane, lane + 32 of each slab): requested before the ring waits
      if (a.has_cab && a.diag) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int h = 0; h < 2; ++h) dv[j][h] = __ldg(a.diag + (long long)b * a.diag_ld + j * CH + lane + 32 * h);
      }
      for (int kb = 0; kb < 3; ++kb) {
        uint8_t* d0 = stage_begin();
        if (elect_one()) {
          mbar_arrive_expect_tx(&w_full[st], KBLK);
          tma_load_4d(d0, &tm.A0, &w_full[st], kb * KB, x0, y0, b);
        }
        stage_end();
        uint8_t* d1 = stage_begin();
        if (elect_one()) {
          mbar_arrive_expect_tx(&w_full[st], WSTAGE);
          tma_load_2d(d1, &tm.Wp, &w_full[st], kb * KB, wrow);                           // [192 rows][64 k]
        }
        stage_end();
      }
      if (a.has_cab)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          uint8_t* d0 = stage_begin();
          if (a.diag) {
            // the 64 x 64 diagonal slab diag(alpha * s_b[64 j ..]) written in place (128B-swizzled K-major rows): no per-sample weight
            // tensor in HBM, no builder launch.  Lane l owns rows l and l + 32.
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int r = lane + 32 * h;
              uint8_t* rowp = d0 + KBLK + r * 128;
              const uint32_t v = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a.diag_alpha * dv[j][h]));
#pragma unroll
              for (int q = 0; q < 8; ++q) {      // logical 16-byte chunk q holds columns 8 q .. 8 q + 7
                uint4 z = make_uint4(0u, 0u, 0u, 0u);
                if (q == (r >> 3)) {
                  const uint32_t word = (r & 1) ? (v << 16) : v;
                  const int wi = (r & 7) >> 1;
                  z.x = wi == 0 ? word : 0u; z.y = wi == 1 ? word : 0u; z.z = wi == 2 ? word : 0u; z.w = wi == 3 ? word : 0u;
                }
                *reinterpret_cast<uint4*>(rowp + ((q ^ (r & 7)) << 4)) = z;
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
          }
          if (elect_one()) {
            mbar_arrive_expect_tx(&w_full[st], a.diag ? KBLK : KBLK + CH * KB * 2);
            tma_load_4d(d0, &tm.A1, &w_full[st], j * KB, x0, y0, b);
            if (!a.diag) tma_load_2d(d0 + KBLK, &tm.WpD, &w_full[st], CP + j * KB, wrow + j * CH);      // the 64 x 64 diagonal slab
          }
          stage_end();
        }
    };
    auto load_w1 = [&](int c) {
      uint8_t* d = stage_begin();
      if (elect_one()) {
        mbar_arrive_expect_tx(&w_full[st], WSTAGE);
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) tma_load_2d(d + kb * (CH * KB * 2), &tm.W1, &w_full[st], kb * KB, c * CH);      // 3 x [64 rows][64 k]
      }
      stage_end();
    };
    auto load_w2 = [&](int c) {
      uint8_t* d = stage_begin();
      if (elect_one()) {
        mbar_arrive_expect_tx(&w_full[st], WSTAGE);
        tma_load_2d(d, &tm.W2, &w_full[st], c * KB, 0);                                    // [192 rows][64 k]
      }
      stage_end();
    };
    if ((int)blockIdx.x < num_tiles) load_g0(blockIdx.x);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int next = tile + gridDim.x;
      load_w1(0);
      load_w1(1);
      for (int c = 0; c < NCH; ++c) {
        load_w2(c);
        if (c + 2 < NCH) load_w1(c + 2);
        if (c == a.g0_pos && next < num_tiles) load_g0(next);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc_n64 = umma_idesc_bf16(TM, CH);
    constexpr uint32_t idesc_n192 = umma_idesc_bf16(TM, CP);
    const uint64_t desc_a = umma_desc_k_sw128(smem_u32(smem + OFF_A));
    const uint64_t desc_w = umma_desc_k_sw128(smem_u32(smem + OFF_W));
    int st = 0;
    uint32_t st_phase = 0, p_af = 0;
    uint32_t p_hf[2] = {0, 0}, p_acc2e[2] = {0, 0};
    TPROF_T0
    auto next_stage = [&]() -> int {      // waits for the ring stage to be full, returns its index, advances
      const int s = st;
      TPROF(7)
      mbar_wait(&w_full[st], st_phase);
      TPROF(0)
      if (++st == NST) { st = 0; st_phase ^= 1; }
      return s;
    };
    auto g0 = [&](int s2) {
      TPROF(7)
      mbar_wait(&acc2_empty[s2], p_acc2e[s2] ^ 1);
      TPROF(1)
      p_acc2e[s2] ^= 1;
      const uint32_t d = tmem_base + ACC2_COL + s2 * CP;
      for (int kb = 0; kb < 3; ++kb) {
        const int sa = next_stage();
        const int sw = next_stage();
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = desc_w + (uint64_t)((sa * WSTAGE) >> 4);
          const uint64_t db = desc_w + (uint64_t)((sw * WSTAGE) >> 4);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k) tc_mma_bf16(d, da + 2 * k, db + 2 * k, idesc_n192, (kb | k) != 0 ? 1u : 0u);
          tc_commit(&w_empty[sa]);
          tc_commit(&w_empty[sw]);
        }
        __syncwarp();
      }
      if (a.has_cab)
        for (int j = 0; j < 3; ++j) {
          const int sc = next_stage();
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = desc_w + (uint64_t)((sc * WSTAGE) >> 4);
            const uint64_t db = desc_w + (uint64_t)((sc * WSTAGE + KBLK) >> 4);
#pragma unroll
            for (int k = 0; k < KB / 16; ++k) tc_mma_bf16(d + j * CH, da + 2 * k, db + 2 * k, idesc_n64, 1u);
            tc_commit(&w_empty[sc]);
          }
          __syncwarp();
        }
      if (elect_one()) tc_commit(&g0_full[s2]);
      __syncwarp();
    };
    auto g1 = [&](int c, bool last) {
      const int s1 = c & 1;      // (acc1[s1] was last read by G2(c - 2) as its A operand: the tensor pipe executes in issue order)
      const int sw = next_stage();
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + ACC1_COL + s1 * CH;
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          const uint64_t da = desc_a + (uint64_t)((kb * KBLK) >> 4);
          const uint64_t db = desc_w + (uint64_t)((sw * WSTAGE + kb * CH * KB * 2) >> 4);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k) tc_mma_bf16(d, da + 2 * k, db + 2 * k, idesc_n64, (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit(&w_empty[sw]);
        tc_commit(&acc1_full[s1]);
        if (last) tc_commit(&a_empty);
      }
      __syncwarp();
    };
    auto g2 = [&](int c, int s2, bool last) {
      const int s1 = c & 1;
      const int sw = next_stage();
      TPROF(7)
      mbar_wait(&h_full[s1], p_hf[s1]);
      TPROF(3)
      p_hf[s1] ^= 1;
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + ACC2_COL + s2 * CP;      // holds x1: always accumulate
        const uint32_t ha = tmem_base + ACC1_COL + s1 * CH;      // H = GELU(acc1) as bf16 pairs, written in place by the GELU warps
        const uint64_t db = desc_w + (uint64_t)((sw * WSTAGE) >> 4);
#pragma unroll
        for (int k = 0; k < KB / 16; ++k) tc_mma_bf16_ts(d, ha + (k < 2 ? k * 8 : 32 + (k - 2) * 8), db + 2 * k, idesc_n192, 1u);
        tc_commit(&w_empty[sw]);
        if (last) tc_commit(&acc2_full[s2]);
      }
      __syncwarp();
    };
    int it = 0;
    if ((int)blockIdx.x < num_tiles) g0(0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s2 = it & 1;
      const bool has_next = tile + (int)gridDim.x < num_tiles;
      TPROF(7)
      mbar_wait(&a_full, p_af);      // MID: LayerNorm2(x1) is in the A tile and x1 is back in Y
      TPROF(4)
#ifdef FF_TAIL_PROF
      if (blockIdx.x == 0 && lane == 0) atomicAdd(&g_tail_prof[31], 1ull);
#endif
      p_af ^= 1;
      g1(0, false);
      g1(1, false);
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        g2(c, s2, c == NCH - 1);
        if (c + 2 < NCH) g1(c + 2, c + 2 == NCH - 1);
        if (c == a.g0_pos && has_next) g0(s2 ^ 1);
      }
    }
  } else if (warp < W_MID0) {
    // ================= GELU warps: acc1 chunk -> + b1 -> GELU -> bf16 pairs, written back over the chunk's own TMEM columns =================
    // (the A operand of G2 comes from TMEM, as P does in the attention kernels: no shared-memory round trip for the hidden tile)
    const int ew = warp - W_GELU0;
    const int quad = warp & 3;
    const int half = ew >> 2;                   // which 32 of the chunk's 64 columns; the packed result lands in the first 16 of them
    uint32_t p_a1f[2] = {0, 0};
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int s1 = c & 1;
        mbar_wait(&acc1_full[s1], p_a1f[s1]);
        p_a1f[s1] ^= 1;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ACC1_COL + s1 * CH + half * 32 + ((uint32_t)(quad * 32) << 16);
        const float* bias = sP + P_B1 + c * CH + half * 32;
        uint32_t w[16];
        uint32_t raw[2][16];
        tmem_ld16(taddr, raw[0]);
        tmem_ld16(taddr + 16, raw[1]);
        tc_wait_ld();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 bv = *reinterpret_cast<const float2*>(bias + hh * 16 + 2 * i);
            w[hh * 8 + i] = pack_bf16(gelu_tanh_hw(__uint_as_float(raw[hh][2 * i]) + bv.x), gelu_tanh_hw(__uint_as_float(raw[hh][2 * i + 1]) + bv.y));
          }
        tmem_st16(taddr, w);
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_full[s1]);
      }
    }
  } else if (warp < W_FIN0) {
    // ================= MID warps: x1 = Y + b_proj + shortcut -> back into Y; LayerNorm2(x1) -> A tile =================
    const int mw = warp - W_MID0;
    const int quad = warp & 3;
    const int half = mw >> 2;
    const int row = quad * 32 + lane;
    uint8_t* wbase = smem + OFF_MID + mw * MID_WARP_BYTES;
    const int sw3 = (lane >> 1) & 3;
    uint32_t ph[2] = {0, 0}, p_g0[2] = {0, 0}, p_ae = 0;
    if (lane == 0) tma_prefetch_desc(&tm.R);
    auto issue_load = [&](int tile, int sb, int bsel) {      // lane 0
      int b, y0, x0;
      tile_coords(tile, b, y0, x0);
      mbar_arrive_expect_tx(&res_bar[mw][bsel], 2048);
      tma_load_4d(wbase + bsel * 2048, &tm.R, &res_bar[mw][bsel], sb * SUB, x0, y0 + quad * 2, b);
    };
    if (lane == 0 && (int)blockIdx.x < num_tiles) {
      issue_load(blockIdx.x, half, 0);
      issue_load(blockIdx.x, half + 2, 1);
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s2 = it & 1;
      TPROF_T0
      mbar_wait(&g0_full[s2], p_g0[s2]);
      if (mw == 0) TPROF(8)
      p_g0[s2] ^= 1;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ACC2_COL + s2 * CP + ((uint32_t)(quad * 32) << 16);
      float s1sum = 0.f, s2sum = 0.f;
      uint32_t raw[2][16];
      tmem_ld16(taddr + half * SUB, raw[0]);
#pragma unroll 2
      for (int j = 0; j < SUB_PER_WARP; ++j) {      // (the next sub-block's TMEM load is in flight while this one is updated)
        const int sb = half + 2 * j;
        const int buf = j & 1;
        if (mw == 0) TPROF(9)
        mbar_wait(&res_bar[mw][buf], ph[buf]);
        if (mw == 0) TPROF(10)
        ph[buf] ^= 1;
        tc_wait_ld();
        if (j + 1 < SUB_PER_WARP) tmem_ld16(taddr + (sb + 2) * SUB, raw[(j + 1) & 1]);
        const uint8_t* rrow = wbase + buf * 2048 + lane * 64;
        uint32_t (&rw)[16] = raw[j & 1];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 bb = *reinterpret_cast<const float4*>(sP + P_BP + sb * SUB + c * 4);
          const float4 r = *reinterpret_cast<const float4*>(rrow + ((c ^ sw3) << 4));
          const float v0 = r.x + (__uint_as_float(rw[4 * c]) + bb.x);
          const float v1 = r.y + (__uint_as_float(rw[4 * c + 1]) + bb.y);
          const float v2 = r.z + (__uint_as_float(rw[4 * c + 2]) + bb.z);
          const float v3 = r.w + (__uint_as_float(rw[4 * c + 3]) + bb.w);
          s1sum += (v0 + v1) + (v2 + v3);
          s2sum += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
          rw[4 * c] = __float_as_uint(v0); rw[4 * c + 1] = __float_as_uint(v1);
          rw[4 * c + 2] = __float_as_uint(v2); rw[4 * c + 3] = __float_as_uint(v3);
        }
        tmem_st16(taddr + sb * SUB, rw);
        __syncwarp();      // every lane has read the residual buffer: refill it with the sub-block two items ahead
        int ntile = tile, nj = j + 2;
        if (nj >= SUB_PER_WARP) { nj -= SUB_PER_WARP; ntile = tile + gridDim.x; }
        if (lane == 0 && ntile < num_tiles) issue_load(ntile, half + 2 * nj, buf);
      }
      // row statistics = both warps' partial sums (two 64-thread named barriers per tile)
      float rstd, nmr;
      {
        float2* slot = &ln_part[0][row];
        if (half == 1) *slot = make_float2(s1sum, s2sum);
        named_bar_sync(1 + quad, 64);
        if (half == 0) {
          const float2 other = *slot;
          const float inv_c = 1.0f / (float)a.ln_cols;
          const float mean = (s1sum + other.x) * inv_c;
          const float var = fmaxf((s2sum + other.y) * inv_c - mean * mean, 0.f);
          rstd = rsqrtf(var + a.ln_eps);
          nmr = -mean * rstd;
          *slot = make_float2(rstd, nmr);
        }
        named_bar_sync(1 + quad, 64);
        if (half == 1) { const float2 st2 = *slot; rstd = st2.x; nmr = st2.y; }
      }
      tc_wait_st();
      tmem_ld16(taddr + half * SUB, raw[0]);
      if (mw == 0) TPROF(9)
      if (it > 0) {      // G1(5) of the previous tile has drained the A tile
        mbar_wait(&a_empty, p_ae);
        p_ae ^= 1;
      }
      if (mw == 0) TPROF(11)
#pragma unroll 2
      for (int j = 0; j < SUB_PER_WARP; ++j) {
        const int sb = half + 2 * j;
        tc_wait_ld();
        if (j + 1 < SUB_PER_WARP) tmem_ld16(taddr + (sb + 2) * SUB, raw[(j + 1) & 1]);
        const uint32_t (&xr)[16] = raw[j & 1];
        uint32_t pk[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 g = *reinterpret_cast<const float4*>(sP + P_G2 + sb * SUB + c * 4);
          const float4 be = *reinterpret_cast<const float4*>(sP + P_BE2 + sb * SUB + c * 4);
          const float y0v = fmaf(fmaf(__uint_as_float(xr[4 * c]), rstd, nmr), g.x, be.x);
          const float y1v = fmaf(fmaf(__uint_as_float(xr[4 * c + 1]), rstd, nmr), g.y, be.y);
          const float y2v = fmaf(fmaf(__uint_as_float(xr[4 * c + 2]), rstd, nmr), g.z, be.z);
          const float y3v = fmaf(fmaf(__uint_as_float(xr[4 * c + 3]), rstd, nmr), g.w, be.w);
          pk[2 * c] = pack_bf16(y0v, y1v);
          pk[2 * c + 1] = pack_bf16(y2v, y3v);
        }
        // columns 16 sb .. +16 = 16-byte chunks (sb % 4) * 2, +1 of k-block sb / 4
        uint8_t* arow = smem + OFF_A + (sb >> 2) * KBLK + row * 128;
        const int q0 = (sb & 3) * 2;
        *reinterpret_cast<uint4*>(arow + ((q0 ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(arow + (((q0 + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full);
      if (mw == 0) TPROF(12)
    }
  } else {
    // ================= FIN warps: x2 = Y + b2 -> fp32 store (+ bf16 copy); LayerNorm'(x2) -> bf16 store =================
    const int fw = warp - W_FIN0;
    const int quad = warp & 3;
    const int half = fw >> 2;
    const int row = quad * 32 + lane;
    uint8_t* wbase = smem + OFF_FIN + fw * FIN_WARP_BYTES;      // [F 2K][S 1K]
    const int sw3 = (lane >> 1) & 3, sw1 = (lane >> 2) & 1;
    uint32_t p_a2f[2] = {0, 0};
    if (lane == 0) { tma_prefetch_desc(&tm.O32); if (a.has_bf16) tma_prefetch_desc(&tm.O16); if (a.has_ln) tma_prefetch_desc(&tm.LN); }
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s2 = it & 1;
      int b, y0, x0;
      tile_coords(tile, b, y0, x0);
      TPROF_T0
      mbar_wait(&acc2_full[s2], p_a2f[s2]);
      if (fw == 0) TPROF(16)
      p_a2f[s2] ^= 1;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ACC2_COL + s2 * CP + ((uint32_t)(quad * 32) << 16);
      float s1sum = 0.f, s2sum = 0.f;
      uint32_t raw[2][16];
      tmem_ld16(taddr + half * SUB, raw[0]);
#pragma unroll 2
      for (int j = 0; j < SUB_PER_WARP; ++j) {
        const int sb = half + 2 * j;
        tc_wait_ld();
        if (j + 1 < SUB_PER_WARP) tmem_ld16(taddr + (sb + 2) * SUB, raw[(j + 1) & 1]);
        uint32_t (&rw)[16] = raw[j & 1];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 bb = *reinterpret_cast<const float4*>(sP + P_B2 + sb * SUB + c * 4);
          const float v0 = __uint_as_float(rw[4 * c]) + bb.x, v1 = __uint_as_float(rw[4 * c + 1]) + bb.y;
          const float v2 = __uint_as_float(rw[4 * c + 2]) + bb.z, v3 = __uint_as_float(rw[4 * c + 3]) + bb.w;
          s1sum += (v0 + v1) + (v2 + v3);
          s2sum += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
          rw[4 * c] = __float_as_uint(v0); rw[4 * c + 1] = __float_as_uint(v1);
          rw[4 * c + 2] = __float_as_uint(v2); rw[4 * c + 3] = __float_as_uint(v3);
        }
        if (a.has_ln) tmem_st16(taddr + sb * SUB, rw);
        if (lane == 0) tma_store_wait_read<0>();      // the previous sub-block's stores have read F / S
        __syncwarp();
        uint8_t* frow = wbase + lane * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(frow + ((c ^ sw3) << 4)) = make_uint4(rw[4 * c], rw[4 * c + 1], rw[4 * c + 2], rw[4 * c + 3]);
        if (a.has_bf16) {
          uint8_t* srow = wbase + 2048 + lane * 32;
#pragma unroll
          for (int c = 0; c < 2; ++c)
            *reinterpret_cast<uint4*>(srow + ((c ^ sw1) << 4)) =
                make_uint4(pack_bf16(__uint_as_float(rw[8 * c]), __uint_as_float(rw[8 * c + 1])), pack_bf16(__uint_as_float(rw[8 * c + 2]), __uint_as_float(rw[8 * c + 3])),
                           pack_bf16(__uint_as_float(rw[8 * c + 4]), __uint_as_float(rw[8 * c + 5])), pack_bf16(__uint_as_float(rw[8 * c + 6]), __uint_as_float(rw[8 * c + 7])));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tm.O32, wbase, sb * SUB, x0, y0 + quad * 2, b);
          if (a.has_bf16) tma_store_4d(&tm.O16, wbase + 2048, sb * SUB, x0, y0 + quad * 2, b);
          tma_store_commit();
        }
      }
      if (fw == 0) TPROF(17)
      if (a.has_ln) {
        float rstd, nmr;
        {
          float2* slot = &ln_part[1][row];
          if (half == 1) *slot = make_float2(s1sum, s2sum);
          named_bar_sync(5 + quad, 64);
          if (half == 0) {
            const float2 other = *slot;
            const float inv_c = 1.0f / (float)a.ln_cols;
            const float mean = (s1sum + other.x) * inv_c;
            const float var = fmaxf((s2sum + other.y) * inv_c - mean * mean, 0.f);
            rstd = rsqrtf(var + a.ln_eps);
            nmr = -mean * rstd;
            *slot = make_float2(rstd, nmr);
          }
          named_bar_sync(5 + quad, 64);
          if (half == 1) { const float2 st2 = *slot; rstd = st2.x; nmr = st2.y; }
        }
        tc_wait_st();
        tmem_ld16(taddr + half * SUB, raw[0]);
        if (lane == 0) tma_store_wait_read<0>();      // F is reused as two bf16 stages
        __syncwarp();
#pragma unroll 2
        for (int j = 0; j < SUB_PER_WARP; ++j) {
          const int sb = half + 2 * j;
          tc_wait_ld();
          if (j + 1 < SUB_PER_WARP) tmem_ld16(taddr + (sb + 2) * SUB, raw[(j + 1) & 1]);
          const uint32_t (&xr)[16] = raw[j & 1];
          uint32_t pk[8];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 g = *reinterpret_cast<const float4*>(sP + P_LNG + sb * SUB + c * 4);
            const float4 be = *reinterpret_cast<const float4*>(sP + P_LNB + sb * SUB + c * 4);
            const float y0v = fmaf(fmaf(__uint_as_float(xr[4 * c]), rstd, nmr), g.x, be.x);
            const float y1v = fmaf(fmaf(__uint_as_float(xr[4 * c + 1]), rstd, nmr), g.y, be.y);
            const float y2v = fmaf(fmaf(__uint_as_float(xr[4 * c + 2]), rstd, nmr), g.z, be.z);
            const float y3v = fmaf(fmaf(__uint_as_float(xr[4 * c + 3]), rstd, nmr), g.w, be.w);
            pk[2 * c] = pack_bf16(y0v, y1v);
            pk[2 * c + 1] = pack_bf16(y2v, y3v);
          }
          uint8_t* stg = wbase + (j & 1) * 1024;
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          uint8_t* srow = stg + lane * 32;
          *reinterpret_cast<uint4*>(srow + ((0 ^ sw1) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(srow + ((1 ^ sw1) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tm.LN, stg, sb * SUB, x0, y0 + quad * 2, b);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc2_empty[s2]);
      if (fw == 0) TPROF(18)
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace

extern "C" int ff_hab_tail(const FFHabTail* pp, void* stream) {
  FF_CHECK_ARG(pp != nullptr, "ff_hab_tail: null params");
  const FFHabTail& p = *pp;
  FF_CHECK_ARG(p.a0 && p.wp && p.bp && p.res && p.ln2_gamma && p.ln2_beta && p.w1 && p.b1 && p.w2 && p.b2 && p.x, "ff_hab_tail: null buffer");
  FF_CHECK_ARG(p.B > 0 && p.H > 0 && p.W > 0, "ff_hab_tail: bad size");
  FF_CHECK_ARG(p.a0_ld % 8 == 0 && p.a0_ld >= CP && p.res_ld % 4 == 0 && p.res_ld >= CP && p.x_ld % 4 == 0 && p.x_ld >= CP,
               "ff_hab_tail: bad pitches (a0_ld=%d res_ld=%d x_ld=%d)", p.a0_ld, p.res_ld, p.x_ld);
  FF_CHECK_ARG(p.wp_batch_rows == 0 || p.wp_batch_rows == CP, "ff_hab_tail: wp_batch_rows must be 0 or %d", CP);
  FF_CHECK_ARG(p.ln_cols > 0 && p.ln_cols <= CP && p.ln_eps > 0.f, "ff_hab_tail: bad LayerNorm width / eps");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  FF_CHECK_ARG(al16(p.a0) && al16(p.wp) && al16(p.bp) && al16(p.res) && al16(p.ln2_gamma) && al16(p.ln2_beta) && al16(p.w1) && al16(p.w2) && al16(p.b1) &&
                   al16(p.b2) && al16(p.x), "ff_hab_tail: operands must be 16-byte aligned");
  if (p.a1) FF_CHECK_ARG(al16(p.a1) && p.a1_ld % 8 == 0 && p.a1_ld >= CP, "ff_hab_tail: bad a1 / a1_ld");
  if (p.a1 && p.a1_diag) FF_CHECK_ARG(p.a1_diag_ld >= CP && p.wp_batch_rows == 0, "ff_hab_tail: a1_diag needs a1_diag_ld >= %d and a shared wp [192][192]", CP);
  if (p.out_bf16) FF_CHECK_ARG(al16(p.out_bf16) && p.out_ld % 8 == 0 && p.out_ld >= CP, "ff_hab_tail: bad out_bf16 / out_ld");
  if (p.ln_out) FF_CHECK_ARG(al16(p.ln_out) && p.ln_out_ld % 8 == 0 && p.ln_out_ld >= CP && p.ln_gamma && p.ln_beta && al16(p.ln_gamma) && al16(p.ln_beta),
                             "ff_hab_tail: bad LayerNorm output operands");
  EncodeTiledFn enc = get_encode();
  if (!enc) { ff_set_error("ff_hab_tail: cuTensorMapEncodeTiled entry point unavailable"); return FF_ERR_DRIVER; }
  TailMaps tm;
  auto img_map = [&](CUtensorMap* m, const void* ptr, int ld, int esz, CUtensorMapDataType dt, CUtensorMapSwizzle sw, int box_c, int box_w, int box_h, int ncols) {
    cuuint64_t dims[4] = {(cuuint64_t)ncols, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    cuuint64_t strides[3] = {(cuuint64_t)ld * esz, (cuuint64_t)ld * esz * p.W, (cuuint64_t)ld * esz * p.W * p.H};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  auto w_map = [&](CUtensorMap* m, const void* ptr, int rows, int K, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  const int K0 = (p.a1 && !p.a1_diag) ? 2 * CP : CP;
  const int wp_rows = p.wp_batch_rows ? p.B * CP : CP;
  const CUtensorMapDataType BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, F32T = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  bool ok = img_map(&tm.A0, p.a0, p.a0_ld, 2, BF, CU_TENSOR_MAP_SWIZZLE_128B, KB, TW_, TH_, CP) && w_map(&tm.Wp, p.wp, wp_rows, K0, CP) &&
            w_map(&tm.W1, p.w1, HID, CP, CH) && w_map(&tm.W2, p.w2, CP, HID, CP) &&
            img_map(&tm.R, p.res, p.res_ld, 4, F32T, CU_TENSOR_MAP_SWIZZLE_64B, SUB, TW_, 2, CP) &&
            img_map(&tm.O32, p.x, p.x_ld, 4, F32T, CU_TENSOR_MAP_SWIZZLE_64B, SUB, TW_, 2, CP);
  tm.A1 = tm.A0; tm.WpD = tm.Wp; tm.O16 = tm.A0; tm.LN = tm.A0;
  if (ok && p.a1) ok = img_map(&tm.A1, p.a1, p.a1_ld, 2, BF, CU_TENSOR_MAP_SWIZZLE_128B, KB, TW_, TH_, CP) && (p.a1_diag || w_map(&tm.WpD, p.wp, wp_rows, K0, CH));
  if (ok && p.out_bf16) ok = img_map(&tm.O16, p.out_bf16, p.out_ld, 2, BF, CU_TENSOR_MAP_SWIZZLE_32B, SUB, TW_, 2, CP);
  if (ok && p.ln_out) ok = img_map(&tm.LN, p.ln_out, p.ln_out_ld, 2, BF, CU_TENSOR_MAP_SWIZZLE_32B, SUB, TW_, 2, CP);
  FF_CHECK_ARG(ok, "ff_hab_tail: cuTensorMapEncodeTiled failed");
  Args a;
  a.B = p.B; a.H = p.H; a.W = p.W;
  a.tiles_x = ff_cdiv(p.W, TW_);
  a.tiles_per_img = a.tiles_x * ff_cdiv(p.H, TH_);
  a.m_tiles = a.tiles_per_img * p.B;
  a.has_cab = p.a1 ? 1 : 0;
  a.wp_batch_rows = p.wp_batch_rows;
  a.diag = p.a1 ? p.a1_diag : nullptr; a.diag_ld = p.a1_diag_ld; a.diag_alpha = p.a1_alpha;
  a.bp = p.bp; a.g2 = p.ln2_gamma; a.be2 = p.ln2_beta; a.b1 = p.b1; a.b2 = p.b2;
  a.ln_gamma = p.ln_gamma; a.ln_beta = p.ln_beta; a.ln_eps = p.ln_eps; a.ln_cols = p.ln_cols;
  a.has_bf16 = p.out_bf16 ? 1 : 0;
  a.has_ln = p.ln_out ? 1 : 0;
  static const int g0_pos = []() { const char* e = getenv("FFB200_TAIL_G0POS"); return e ? atoi(e) : 1; }();
  a.g0_pos = g0_pos < 0 ? 0 : (g0_pos > 5 ? 5 : g0_pos);
  static FFPerDeviceFlag configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(hab_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { ff_set_error("ff_hab_tail: cudaFuncSetAttribute(%d) failed: %s", SMEM_BYTES, cudaGetErrorString(e)); return FF_ERR_CUDA; }
    configured = true;
  }
  const int grid = a.m_tiles < ff_num_sms() ? a.m_tiles : ff_num_sms();
  const cudaError_t le = ff_launch_pdl(hab_tail_kernel, dim3(grid), dim3(NTHREADS), SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream), tm, a);
  if (le != cudaSuccess) { ff_set_error("ff_hab_tail: launch failed: %s", cudaGetErrorString(le)); return FF_ERR_CUDA; }
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_hab_tail");
  return FF_OK;
}

#ifdef FF_TAIL_PROF
extern "C" int ff_debug_tail_prof(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_tail_prof, sizeof(unsigned long long) * 32);
  if (reset) { unsigned long long z[32] = {}; cudaMemcpyToSymbol(g_tail_prof, z, sizeof(z)); }
  return 0;
}
#endif
