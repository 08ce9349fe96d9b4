// ff_mlp_fused: the transformer MLP  x <- x + fc2(GELU(fc1(t)))  of HAT's HAB / OCAB (hat_arch.py:77-94, :308, :437) as ONE
// kernel for sm_100a: the 128 x 384 hidden tile never leaves the SM (the two-kernel version writes and re-reads it through HBM:
// 2 x 201 MB per block at the bench shape, 84 blocks per step).
//
//   t (bf16, LayerNorm output, [M][192])  --TMA-->  A tile 128 x 192 (three 128B-swizzled k-blocks)
//   for hidden chunk c = 0..5 (64 columns each):
//     G1(c)  acc1[c&1] (TMEM, 64 cols)  = A . W1[64c..64c+64, :]^T          tcgen05.mma M=128 N=64  K=16 x 12
//     GELU   8 warps: tcgen05.ld -> + b1 -> tanh-form GELU (hardware tanh.approx, as the stand-alone fc1 epilogue) -> bf16 ->
//            H[c&1] in smem, written directly in the canonical 128B-swizzled K-major layout (= an A operand)
//     G2(c)  acc2 (TMEM, 192 cols) += H[c&1] . W2[:, 64c..64c+64]^T          tcgen05.mma M=128 N=192 K=16 x 4
//   final    8 warps: x_new = acc2 + b2 + x (fp32 residual block TMA-loaded one block ahead, updated in 128B-swizzled smem,
//            TMA-stored), optional bf16 copy, optional fused LayerNorm of x_new for the next consumer (row statistics exchanged
//            between the two warps of a quadrant, x_new written back over the accumulator, re-read, normalised, TMA-stored).
// Persistent CTAs (one per SM), warp-specialised: warp 0 = TMA producer (A + a 3-stage ring of 24 KB weight stages: the three
// k-blocks of a W1 chunk, or one k-block of W2), warp 1 = MMA issuer, warps 2-9 = GELU, warps 10-17 = final epilogue (two per TMEM
// lane quadrant).  acc1 and acc2 are double-buffered (2 x 64 + 2 x 192 = 512 TMEM columns), so G1 of chunk c+1 / the next tile
// overlaps the GELU of chunk c / the final epilogue of this tile.  The weights (2 x 147 KB) stream from L2 once per tile.
#include "ff_common.cuh"
#include "../../include/ffb200.h"

extern long long g_ff_launches;

namespace {

constexpr int TM = 128, TW_ = 16, TH_ = 8;      // output tile: 8 rows x 16 pixels
constexpr int CIN = 192, HID = 384, COUT = 192;
constexpr int CH = 64, NCH = HID / CH;           // hidden chunks
constexpr int KB = 64;                            // bf16 per k-block = 128 B
constexpr int A_BYTES = 3 * TM * KB * 2;          // 48 KB
constexpr int WSTAGE = 24 * 1024, NST = 3;
constexpr int H_BYTES = TM * KB * 2;              // 16 KB, single buffer (the GELU math of chunk c+1 runs while G2(c) drains it)
constexpr int FIN_WARPS = 8;                      // two per TMEM lane quadrant
constexpr int FIN_WARP_BYTES = 10 * 1024;         // [R0 4K][R1 4K][S 2K]
constexpr int OFF_A = 0, OFF_W = OFF_A + A_BYTES, OFF_H = OFF_W + NST * WSTAGE, OFF_FIN = OFF_H + H_BYTES;
constexpr int SMEM_BYTES = OFF_FIN + FIN_WARPS * FIN_WARP_BYTES + 1024;
constexpr int NTHREADS = 32 * (2 + 8 + FIN_WARPS);
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t ACC1_COL = 0, ACC2_COL = 128;

struct Args {
  int B, H, W;
  int tiles_x, tiles_per_img, m_tiles;
  const float* b1;       // [384]
  const float* b2;       // [192]
  const float* ln_gamma; // [192] or null
  const float* ln_beta;
  float ln_eps;
  int ln_cols;
  int has_bf16, has_ln;
};

#ifdef FF_MLP_PROF
// development build: cycle counters of block 0 (MMA warp waits per barrier class, GELU warp 2, final warp 10, producer)
__device__ unsigned long long g_mlp_prof[32];
#define MPROF_T0 long long mp_t = clock64();
#define MPROF(i) { const long long t_ = clock64(); if (blockIdx.x == 0 && lane == 0) atomicAdd(&g_mlp_prof[i], (unsigned long long)(t_ - mp_t)); mp_t = t_; }
#else
#define MPROF_T0
#define MPROF(i) {}
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

__global__ void __launch_bounds__(NTHREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmO16,
                 const __grid_constant__ CUtensorMap tmLN, const __grid_constant__ Args a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full, a_empty;
  __shared__ __align__(8) uint64_t w_full[NST], w_empty[NST];
  __shared__ __align__(8) uint64_t acc1_full[2], acc1_empty[2], h_full, h_empty, acc2_full[2], acc2_empty[2];
  __shared__ __align__(8) uint64_t res_bar[FIN_WARPS][2];
  __shared__ float2 ln_part[TM];      // per-row exchange slot between the two final warps of a lane quadrant
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = a.m_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    mbar_init(&a_full, 1); mbar_init(&a_empty, 1);
    for (int s = 0; s < NST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc1_full[s], 1); mbar_init(&acc1_empty[s], 8);
      mbar_init(&acc2_full[s], 1); mbar_init(&acc2_empty[s], FIN_WARPS);
    }
    mbar_init(&h_full, 8); mbar_init(&h_empty, 1);
    for (int w = 0; w < FIN_WARPS; ++w) { mbar_init(&res_bar[w][0], 1); mbar_init(&res_bar[w][1], 1); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  auto tile_coords = [&](int tile, int& b, int& y0, int& x0) {
    b = tile / a.tiles_per_img;
    const int t = tile - b * a.tiles_per_img;
    const int ty = t / a.tiles_x;
    y0 = ty * TH_;
    x0 = (t - ty * a.tiles_x) * TW_;
  };

  if (warp == 0) {
    // ================= TMA producer =================
    int st = 0;
    uint32_t st_phase = 0, a_phase = 0;
    auto load_a = [&](int tile) {      // elected lane
      int b, y0, x0;
      tile_coords(tile, b, y0, x0);
      mbar_arrive_expect_tx(&a_full, A_BYTES);
#pragma unroll
      for (int kb = 0; kb < 3; ++kb) tma_load_4d(smem + OFF_A + kb * (TM * KB * 2), &tmA, &a_full, kb * KB, x0, y0, b);
    };
    // weight stage sequence of one tile, in the MMA warp's consumption order: W1(0) W1(1) W2(0) W1(2) W2(1) ... W2(4) W2(5)
    bool a_next_loaded = false;
    if (blockIdx.x < num_tiles) {
      if (elect_one()) load_a(blockIdx.x);      // first tile: the buffer is free
      __syncwarp();
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int next = tile + gridDim.x;
      a_next_loaded = next >= num_tiles;
      for (int i = 0; i < 2 * NCH; ++i) {
        // i -> (is_w2, chunk): 0:W1(0) 1:W1(1) then pairs (W2(c), W1(c+2))
        int is_w2, c;
        if (i < 2) { is_w2 = 0; c = i; }
        else { const int j = i - 2; is_w2 = (j & 1) == 0 || j >= 8; c = j >= 8 ? j - 4 : (is_w2 ? j >> 1 : (j >> 1) + 2); }
        { MPROF_T0 mbar_wait(&w_empty[st], st_phase ^ 1); MPROF(0) }
        if (elect_one()) {
          uint8_t* dst = smem + OFF_W + st * WSTAGE;
          mbar_arrive_expect_tx(&w_full[st], WSTAGE);
          if (is_w2) {
            tma_load_2d(dst, &tmW2, &w_full[st], c * KB, 0);                                   // [192 rows][64 k]
          } else {
#pragma unroll
            for (int kb = 0; kb < 3; ++kb) tma_load_2d(dst + kb * (CH * KB * 2), &tmW1, &w_full[st], kb * KB, c * CH);   // 3 x [64 rows][64 k]
          }
        }
        __syncwarp();
        if (++st == NST) { st = 0; st_phase ^= 1; }
        // the A tile of the next work item as soon as G1(5) of this one has drained the buffer
        if (!a_next_loaded && mbar_test_wait(&a_empty, a_phase)) {      // (test_wait: try_wait would park the producer until its time-out)
          a_phase ^= 1;
          a_next_loaded = true;
          if (elect_one()) load_a(next);
          __syncwarp();
        }
      }
      if (!a_next_loaded) {
        mbar_wait(&a_empty, a_phase);
        a_phase ^= 1;
        if (elect_one()) load_a(next);
        __syncwarp();
      } else if (next >= num_tiles) {
        // keep the phase bookkeeping consistent (nobody waits on the last a_empty)
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc1 = umma_idesc_bf16(TM, CH);
    constexpr uint32_t idesc2 = umma_idesc_bf16(TM, COUT);
    const uint64_t desc_a = umma_desc_k_sw128(smem_u32(smem + OFF_A));
    const uint64_t desc_w = umma_desc_k_sw128(smem_u32(smem + OFF_W));
    const uint64_t desc_h = umma_desc_k_sw128(smem_u32(smem + OFF_H));
    int st = 0;
    uint32_t st_phase = 0, a_phase = 0;
    uint32_t p_acc1e[2] = {0, 0}, p_hf = 0;
    uint32_t p_acc2e[2] = {0, 0};
    int it = 0;
    auto g1 = [&](int c, bool last) {
      const int s1 = c & 1;
      MPROF_T0
      mbar_wait(&w_full[st], st_phase);
      MPROF(1)
      mbar_wait(&acc1_empty[s1], p_acc1e[s1] ^ 1);
      MPROF(2)
      p_acc1e[s1] ^= 1;
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + ACC1_COL + s1 * CH;
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          const uint64_t da = desc_a + (uint64_t)((kb * TM * KB * 2) >> 4);
          const uint64_t db = desc_w + (uint64_t)((st * WSTAGE + kb * CH * KB * 2) >> 4);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k) tc_mma_bf16(d, da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit(&w_empty[st]);
        tc_commit(&acc1_full[s1]);
        if (last) tc_commit(&a_empty);
      }
      __syncwarp();
      MPROF(3)
      if (++st == NST) { st = 0; st_phase ^= 1; }
    };
    auto g2 = [&](int c, int s2, bool last) {
      MPROF_T0
      mbar_wait(&w_full[st], st_phase);
      MPROF(4)
      mbar_wait(&h_full, p_hf);
      MPROF(5)
      p_hf ^= 1;
      if (c == 0) {
        mbar_wait(&acc2_empty[s2], p_acc2e[s2] ^ 1);
        p_acc2e[s2] ^= 1;
      }
      MPROF(6)
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + ACC2_COL + s2 * COUT;
        const uint64_t da = desc_h;
        const uint64_t db = desc_w + (uint64_t)((st * WSTAGE) >> 4);
#pragma unroll
        for (int k = 0; k < KB / 16; ++k) tc_mma_bf16(d, da + 2 * k, db + 2 * k, idesc2, (c | k) != 0 ? 1u : 0u);
        tc_commit(&w_empty[st]);
        tc_commit(&h_empty);
        if (last) tc_commit(&acc2_full[s2]);
      }
      __syncwarp();
      MPROF(7)
      if (++st == NST) { st = 0; st_phase ^= 1; }
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s2 = it & 1;
      { MPROF_T0 mbar_wait(&a_full, a_phase); MPROF(8) }
      a_phase ^= 1;
      g1(0, false);
      g1(1, false);
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        g2(c, s2, c == NCH - 1);
        if (c + 2 < NCH) g1(c + 2, c + 2 == NCH - 1);
      }
    }
  } else if (warp < 10) {
    // ================= GELU warps: acc1 chunk -> + b1 -> GELU -> bf16 -> H (K-major, 128B swizzle) =================
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int half = ew >> 2;                   // which 32 of the chunk's 64 columns
    const int row = quad * 32 + lane;           // tile row = TMEM lane
    uint32_t p_a1f[2] = {0, 0}, p_he = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int s1 = c & 1;
        MPROF_T0
        mbar_wait(&acc1_full[s1], p_a1f[s1]);
        if (ew == 0) MPROF(10)
        p_a1f[s1] ^= 1;
        tc_fence_after();
        uint32_t raw[32];
        const uint32_t taddr = tmem_base + ACC1_COL + s1 * CH + half * 32 + ((uint32_t)(quad * 32) << 16);
        tmem_ld16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&raw[0]));
        tmem_ld16(taddr + 16, *reinterpret_cast<uint32_t(*)[16]>(&raw[16]));
        float4 bq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(a.b1 + c * CH + half * 32) + j);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc1_empty[s1]);      // the accumulator chunk is in registers
        if (ew == 0) MPROF(11)
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 bv = bq[i >> 1];
          const float b0 = (i & 1) ? bv.z : bv.x, b1v = (i & 1) ? bv.w : bv.y;
          w[i] = pack_bf16(gelu_tanh_hw(__uint_as_float(raw[2 * i]) + b0), gelu_tanh_hw(__uint_as_float(raw[2 * i + 1]) + b1v));
        }
        if (ew == 0) MPROF(13)
        // the single H buffer must have been consumed by G2 of the previous chunk (the math above ran under that wait)
        mbar_wait(&h_empty, p_he ^ 1);
        if (ew == 0) MPROF(12)
        p_he ^= 1;
        uint8_t* hrow = smem + OFF_H + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q)               // 16-byte chunk q of this warp's 64 bytes: columns 8q .. 8q+7
          *reinterpret_cast<uint4*>(hrow + (((half * 4 + q) ^ (row & 7)) << 4)) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_full);
      }
    }
  } else {
    // ================= final epilogue warps: x_new = acc2 + b2 + x; stores; fused LayerNorm =================
    // Two warps per TMEM lane quadrant; the six 32-column blocks of a tile alternate between them (as the residual epilogue of
    // ff_conv_gemm): residual block TMA-loaded one item ahead -> in-place update in 128B-swizzled smem -> TMA store.
    const int fw = warp - 10;                   // 0..7
    const int quad = warp & 3;
    const int half = fw >> 2;
    uint8_t* wbase = smem + OFF_FIN + fw * FIN_WARP_BYTES;      // [R0 4K][R1 4K][S 2K]
    constexpr int NBLK = COUT / 32;             // 6 column blocks per tile, 3 per warp
    int buf = 0;
    uint32_t ph[2] = {0, 0};
    uint32_t p_a2f[2] = {0, 0};
    if (lane == 0) { tma_prefetch_desc(&tmR); tma_prefetch_desc(&tmO32); if (a.has_bf16) tma_prefetch_desc(&tmO16); if (a.has_ln) tma_prefetch_desc(&tmLN); }
    auto issue_load = [&](int tile, int cb, int bsel) {      // lane 0
      int b, y0, x0;
      tile_coords(tile, b, y0, x0);
      mbar_arrive_expect_tx(&res_bar[fw][bsel], 4096);
      tma_load_4d(wbase + bsel * 4096, &tmR, &res_bar[fw][bsel], cb * 32, x0, y0 + quad * 2, b);
    };
    if (lane == 0 && (int)blockIdx.x < num_tiles) issue_load(blockIdx.x, half, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s2 = it & 1;
      int b, y0, x0;
      tile_coords(tile, b, y0, x0);
      MPROF_T0
      mbar_wait(&acc2_full[s2], p_a2f[s2]);
      if (fw == 0) MPROF(16)
      p_a2f[s2] ^= 1;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ACC2_COL + s2 * COUT + ((uint32_t)(quad * 32) << 16);
      float s1sum = 0.f, s2sum = 0.f;
      const int sw7 = lane & 7, sw3 = (lane >> 1) & 3;
#pragma unroll 1
      for (int cb = half; cb < NBLK; cb += 2) {
        // prefetch the residual block of the next item (possibly the next tile) into the other buffer
        int ntile = tile, ncb = cb + 2;
        if (ncb >= NBLK) { ncb = half; ntile = tile + gridDim.x; }
        if (lane == 0) {
          tma_store_wait_read<0>();      // the stores that read R[buf^1] / S (issued one item ago) have drained
          if (ntile < num_tiles) issue_load(ntile, ncb, buf ^ 1);
        }
        uint32_t raw[32];
        tmem_ld16(taddr + cb * 32, *reinterpret_cast<uint32_t(*)[16]>(&raw[0]));
        tmem_ld16(taddr + cb * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&raw[16]));
        tc_wait_ld();
        mbar_wait(&res_bar[fw][buf], ph[buf]);
        ph[buf] ^= 1;
        __syncwarp();
        uint8_t* rrow = wbase + buf * 4096 + lane * 128;
        uint8_t* orow = wbase + 8192 + lane * 64;      // S: bf16 copy staging
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(a.b2 + cb * 32 + c * 4));
          float4* rp = reinterpret_cast<float4*>(rrow + ((c ^ sw7) << 4));
          float4 r = *rp;
          r.x += __uint_as_float(raw[4 * c]) + bb.x;
          r.y += __uint_as_float(raw[4 * c + 1]) + bb.y;
          r.z += __uint_as_float(raw[4 * c + 2]) + bb.z;
          r.w += __uint_as_float(raw[4 * c + 3]) + bb.w;
          *rp = r;
          if (a.has_bf16) *reinterpret_cast<uint2*>(orow + (((c >> 1) ^ sw3) << 4) + ((c & 1) << 3)) = make_uint2(pack_bf16(r.x, r.y), pack_bf16(r.z, r.w));
          if (a.has_ln) {
            s1sum += (r.x + r.y) + (r.z + r.w);
            s2sum += (r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w);
            raw[4 * c] = __float_as_uint(r.x); raw[4 * c + 1] = __float_as_uint(r.y);
            raw[4 * c + 2] = __float_as_uint(r.z); raw[4 * c + 3] = __float_as_uint(r.w);
          }
        }
        if (a.has_ln) {
          tmem_st16(taddr + cb * 32, *reinterpret_cast<uint32_t(*)[16]>(&raw[0]));
          tmem_st16(taddr + cb * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&raw[16]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmO32, wbase + buf * 4096, cb * 32, x0, y0 + quad * 2, b);
          if (a.has_bf16) tma_store_4d(&tmO16, wbase + 8192, cb * 32, x0, y0 + quad * 2, b);
          tma_store_commit();
        }
        buf ^= 1;
      }
      if (a.has_ln) {
        // row statistics = both warps' partial sums: the second warp publishes its partial, the first one combines and
        // publishes (rstd, -mean * rstd) in the same slot (two 64-thread named barriers per tile)
        float rstd, nmr;
        {
          float2* slot = &ln_part[quad * 32 + lane];
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
        if (lane == 0) tma_store_wait_read<0>();      // R[buf^1] / S of the last item have been read
        __syncwarp();
        tc_wait_st();
        int sb = 0;
#pragma unroll 1
        for (int cb = half; cb < NBLK; cb += 2) {
          uint32_t xr[32];
          tmem_ld16(taddr + cb * 32, *reinterpret_cast<uint32_t(*)[16]>(&xr[0]));
          tmem_ld16(taddr + cb * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&xr[16]));
          uint8_t* stg = sb == 0 ? (wbase + (buf ^ 1) * 4096) : (wbase + 8192);      // (the prefetched residual block sits in R[buf])
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          tc_wait_ld();
          uint8_t* srow = stg + lane * 64;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(a.ln_gamma + cb * 32 + c * 4));
            const float4 be = __ldg(reinterpret_cast<const float4*>(a.ln_beta + cb * 32 + c * 4));
            const float y0v = fmaf(fmaf(__uint_as_float(xr[4 * c]), rstd, nmr), g.x, be.x);
            const float y1v = fmaf(fmaf(__uint_as_float(xr[4 * c + 1]), rstd, nmr), g.y, be.y);
            const float y2v = fmaf(fmaf(__uint_as_float(xr[4 * c + 2]), rstd, nmr), g.z, be.z);
            const float y3v = fmaf(fmaf(__uint_as_float(xr[4 * c + 3]), rstd, nmr), g.w, be.w);
            *reinterpret_cast<uint2*>(srow + (((c >> 1) ^ sw3) << 4) + ((c & 1) << 3)) = make_uint2(pack_bf16(y0v, y1v), pack_bf16(y2v, y3v));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmLN, stg, cb * 32, x0, y0 + quad * 2, b);
            tma_store_commit();
          }
          sb ^= 1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc2_empty[s2]);
      if (fw == 0) MPROF(17)
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

extern "C" int ff_mlp_fused(const FFMlpFused* pp, void* stream) {
  FF_CHECK_ARG(pp != nullptr, "ff_mlp_fused: null params");
  const FFMlpFused& p = *pp;
  FF_CHECK_ARG(p.t && p.w1 && p.b1 && p.w2 && p.b2 && p.x, "ff_mlp_fused: null buffer");
  FF_CHECK_ARG(p.B > 0 && p.H > 0 && p.W > 0, "ff_mlp_fused: bad size");
  FF_CHECK_ARG(p.t_ld % 8 == 0 && p.t_ld >= CIN && p.x_ld % 4 == 0 && p.x_ld >= COUT, "ff_mlp_fused: bad pitches (t_ld=%d x_ld=%d)", p.t_ld, p.x_ld);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  FF_CHECK_ARG(al16(p.t) && al16(p.w1) && al16(p.w2) && al16(p.b1) && al16(p.b2) && al16(p.x), "ff_mlp_fused: operands must be 16-byte aligned");
  if (p.out_bf16) FF_CHECK_ARG(al16(p.out_bf16) && p.out_ld % 8 == 0 && p.out_ld >= COUT, "ff_mlp_fused: bad out_bf16 / out_ld");
  if (p.ln_out) FF_CHECK_ARG(al16(p.ln_out) && p.ln_out_ld % 8 == 0 && p.ln_out_ld >= COUT && p.ln_gamma && p.ln_beta && al16(p.ln_gamma) && al16(p.ln_beta) &&
                                 p.ln_cols > 0 && p.ln_cols <= COUT && p.ln_eps > 0.f, "ff_mlp_fused: bad LayerNorm operands");
  EncodeTiledFn enc = get_encode();
  if (!enc) { ff_set_error("ff_mlp_fused: cuTensorMapEncodeTiled entry point unavailable"); return FF_ERR_DRIVER; }
  CUtensorMap tmA, tmW1, tmW2, tmR, tmO32, tmO16, tmLN;
  auto img_map = [&](CUtensorMap* tm, const void* ptr, int ld, int esz, CUtensorMapDataType dt, CUtensorMapSwizzle sw, int box_c, int box_w, int box_h, int ncols) {
    cuuint64_t dims[4] = {(cuuint64_t)ncols, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    cuuint64_t strides[3] = {(cuuint64_t)ld * esz, (cuuint64_t)ld * esz * p.W, (cuuint64_t)ld * esz * p.W * p.H};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(tm, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  auto w_map = [&](CUtensorMap* tm, const void* ptr, int rows, int K, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  bool ok = img_map(&tmA, p.t, p.t_ld, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B, KB, TW_, TH_, CIN) &&
            w_map(&tmW1, p.w1, HID, CIN, CH) && w_map(&tmW2, p.w2, COUT, HID, COUT) &&
            img_map(&tmR, p.x, p.x_ld, 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B, 32, TW_, 2, COUT) &&
            img_map(&tmO32, p.x, p.x_ld, 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B, 32, TW_, 2, COUT);
  tmO16 = tmA; tmLN = tmA;
  if (ok && p.out_bf16) ok = img_map(&tmO16, p.out_bf16, p.out_ld, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B, 32, TW_, 2, COUT);
  if (ok && p.ln_out) ok = img_map(&tmLN, p.ln_out, p.ln_out_ld, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B, 32, TW_, 2, COUT);
  FF_CHECK_ARG(ok, "ff_mlp_fused: cuTensorMapEncodeTiled failed");
  Args a;
  a.B = p.B; a.H = p.H; a.W = p.W;
  a.tiles_x = ff_cdiv(p.W, TW_);
  a.tiles_per_img = a.tiles_x * ff_cdiv(p.H, TH_);
  a.m_tiles = a.tiles_per_img * p.B;
  a.b1 = p.b1; a.b2 = p.b2; a.ln_gamma = p.ln_gamma; a.ln_beta = p.ln_beta; a.ln_eps = p.ln_eps; a.ln_cols = p.ln_cols;
  a.has_bf16 = p.out_bf16 ? 1 : 0;
  a.has_ln = p.ln_out ? 1 : 0;
  static FFPerDeviceFlag configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { ff_set_error("ff_mlp_fused: cudaFuncSetAttribute(%d) failed: %s", SMEM_BYTES, cudaGetErrorString(e)); return FF_ERR_CUDA; }
    configured = true;
  }
  const int grid = a.m_tiles < ff_num_sms() ? a.m_tiles : ff_num_sms();
  mlp_fused_kernel<<<grid, NTHREADS, SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(tmA, tmW1, tmW2, tmR, tmO32, tmO16, tmLN, a);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_mlp_fused");
  return FF_OK;
}

#ifdef FF_MLP_PROF
extern "C" int ff_debug_mlp_prof(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_mlp_prof, sizeof(unsigned long long) * 32);
  if (reset) { unsigned long long z[32] = {}; cudaMemcpyToSymbol(g_mlp_prof, z, sizeof(z)); }
  return 0;
}
#endif
