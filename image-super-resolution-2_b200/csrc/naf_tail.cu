// ff_naf_tail: everything of a 64-channel NAFBlock that follows the SimpleGate depthwise conv, as ONE kernel for sm_100a
// (nafnet_arch.py:118-131):
//
//     y = inp + (conv3(g * sca) ) * beta          t = LayerNorm2d(y)          u = conv4(t)          z = y + conv5(u1 * u2) * gamma
//     t' = LayerNorm2d'(z)   (norm1 of the next block)
//
// At the full-resolution levels of NAFNet-SR the three 1x1 convs are HBM-bound passes over an fp32 stream (conv3 + residual +
// LayerNorm 768 B per pixel, conv4 + gate 256 B, conv5 + residual + LayerNorm 768 B).  All five steps are per-pixel, so one
// kernel reads g (bf16) and inp (fp32) once and writes z (fp32) and t' (bf16) once: 768 B per pixel instead of 1 792.
// beta / gamma / sca are folded into the weights by the caller (w3 = beta * conv3 * sca per sample, w5 = gamma * conv5).
//
//   per 128-pixel tile (8 rows x 16 pixels), thread = pixel = TMEM lane, so a LayerNorm row never leaves its thread:
//   G0     Y  = g . w3_b^T                  tcgen05.mma M128 N64  K16 x4   (A0 tile + the sample's w3 by TMA)
//   EPI1   y = Y + b3 + inp (fp32 tile TMA-loaded as two 128B-swizzled boxes) -> written back over Y (tcgen05.st: y stays in TMEM
//          as the initial value of the conv5 accumulator); LayerNorm2d(y) -> bf16 -> A1 tile (128B-swizzled K-major) in smem
//   G1     U  = A1 . w4^T                   tcgen05.mma M128 N128 K16 x4
//   EPI2   (U[:, j] + b4[j]) * (U[:, 64 + j] + b4[64 + j]) -> bf16 -> written over the A1 tile
//   G2     Y += A1 . w5^T                   tcgen05.mma M128 N64  K16 x4
//   EPI3   z = Y + b5 -> fp32 TMA stores (16-column sub-blocks staged per warp); LayerNorm2d'(z) (or z itself) -> bf16 -> the warp's
//          own rows of the A1 tile -> one TMA store per warp
// Two CTAs per SM (107 KB of shared memory, 256 TMEM columns each) of 6 warps: TMA producer, MMA issuer, 4 epilogue warps.  A CTA's
// chain over one tile is serial; the second CTA and the loads of the next tile (A0 / w3 as soon as G0 has read them, the fp32 tile
// as soon as EPI1 has) cover it.  Y is double buffered so G0 of the next tile is issued behind G2.
#include "ff_common.cuh"
#include "../../include/ffb200.h"

extern long long g_ff_launches;

namespace {

constexpr int TM = 128, TW_ = 16, TH_ = 8;
constexpr int C = 64, C2 = 128;
constexpr int A_BYTES = TM * C * 2;               // 16 KB: one 128 x 64 bf16 operand tile
constexpr int X_BYTES = TM * C * 4;               // 32 KB: the fp32 residual tile as two [128][32 fp32] boxes
constexpr int W3_BYTES = C * C * 2, W4_BYTES = C2 * C * 2, W5_BYTES = C * C * 2;
constexpr int Z_WARP_BYTES = 2048;                // fp32 staging sub-block [32 rows][16 fp32] (64B swizzle) per epilogue warp
constexpr int OFF_A0 = 0, OFF_X = OFF_A0 + A_BYTES, OFF_A1 = OFF_X + X_BYTES, OFF_W3 = OFF_A1 + A_BYTES, OFF_W4 = OFF_W3 + W3_BYTES,
              OFF_W5 = OFF_W4 + W4_BYTES, OFF_Z = OFF_W5 + W5_BYTES, OFF_P = OFF_Z + 4 * Z_WARP_BYTES;
constexpr int P_B3 = 0, P_G2 = 64, P_BE2 = 128, P_B4 = 192, P_B5 = 320, P_LNG = 384, P_LNB = 448, P_FLOATS = 512;
constexpr int SMEM_BYTES = OFF_P + P_FLOATS * 4 + 1024;
constexpr int NTHREADS = 192;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t Y_COL = 0, U_COL = 128;

struct Args {
  int B, H, W;
  int tiles_x, tiles_per_img, m_tiles;
  int w3_batch_rows;
  const float *b3, *g2, *be2, *b4, *b5, *lng, *lnb;
  float eps;
  int has_bf16, has_ln;
};

struct NafMaps {
  CUtensorMap A0, W3, W4, W5, R, O32, O16;
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(NTHREADS, 2) naf_tail_kernel(const __grid_constant__ NafMaps tm, const __grid_constant__ Args a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_full, a0_full, a0_empty, x_full, x_empty, g0_full[2], a1_full, u_full, a2_full, y2_full;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = a.m_tiles;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.A0); tma_prefetch_desc(&tm.W3); tma_prefetch_desc(&tm.W4); tma_prefetch_desc(&tm.W5); tma_prefetch_desc(&tm.R);
    mbar_init(&w_full, 1);
    mbar_init(&a0_full, 1); mbar_init(&a0_empty, 1);
    mbar_init(&x_full, 1); mbar_init(&x_empty, 4);
    mbar_init(&g0_full[0], 1); mbar_init(&g0_full[1], 1);
    mbar_init(&a1_full, 4); mbar_init(&u_full, 1); mbar_init(&a2_full, 4); mbar_init(&y2_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  float* sP = reinterpret_cast<float*>(smem + OFF_P);
  for (int i = threadIdx.x; i < P_FLOATS; i += NTHREADS) {
    float v;
    if (i < P_G2) v = a.b3[i];
    else if (i < P_BE2) v = a.g2[i - P_G2];
    else if (i < P_B4) v = a.be2[i - P_BE2];
    else if (i < P_B5) v = a.b4[i - P_B4];
    else if (i < P_LNG) v = a.b5[i - P_B5];
    else if (i < P_LNB) v = a.has_ln ? a.lng[i - P_LNG] : 1.f;
    else v = a.has_ln ? a.lnb[i - P_LNB] : 0.f;
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
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_arrive_expect_tx(&w_full, W4_BYTES + W5_BYTES);
      tma_load_2d(smem + OFF_W4, &tm.W4, &w_full, 0, 0);
      tma_load_2d(smem + OFF_W5, &tm.W5, &w_full, 0, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        int b, y0, x0;
        tile_coords(tile, b, y0, x0);
        if (it > 0) mbar_wait(&a0_empty, (it - 1) & 1);      // G0 of the previous tile has read A0 and w3
        mbar_arrive_expect_tx(&a0_full, A_BYTES + W3_BYTES);
        tma_load_4d(smem + OFF_A0, &tm.A0, &a0_full, 0, x0, y0, b);
        tma_load_2d(smem + OFF_W3, &tm.W3, &a0_full, 0, b * a.w3_batch_rows);
        if (it > 0) mbar_wait(&x_empty, (it - 1) & 1);       // EPI1 of the previous tile has read the fp32 tile
        mbar_arrive_expect_tx(&x_full, X_BYTES);
        tma_load_4d(smem + OFF_X, &tm.R, &x_full, 0, x0, y0, b);
        tma_load_4d(smem + OFF_X + X_BYTES / 2, &tm.R, &x_full, 32, x0, y0, b);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc_n64 = umma_idesc_bf16(TM, C);
      constexpr uint32_t idesc_n128 = umma_idesc_bf16(TM, C2);
      const uint64_t desc_a0 = umma_desc_k_sw128(smem_u32(smem + OFF_A0));
      const uint64_t desc_a1 = umma_desc_k_sw128(smem_u32(smem + OFF_A1));
      const uint64_t desc_w3 = umma_desc_k_sw128(smem_u32(smem + OFF_W3));
      const uint64_t desc_w4 = umma_desc_k_sw128(smem_u32(smem + OFF_W4));
      const uint64_t desc_w5 = umma_desc_k_sw128(smem_u32(smem + OFF_W5));
      auto g0 = [&](int s, int it) {
        mbar_wait(&a0_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < C / 16; ++k) tc_mma_bf16(tmem_base + Y_COL + s * C, desc_a0 + 2 * k, desc_w3 + 2 * k, idesc_n64, k != 0 ? 1u : 0u);
        tc_commit(&g0_full[s]);
        tc_commit(&a0_empty);
      };
      int it = 0;
      if ((int)blockIdx.x < num_tiles) g0(0, 0);
      mbar_wait(&w_full, 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(&a1_full, it & 1);      // EPI1: LayerNorm2d(y) is in the A1 tile and y is back in Y[s]
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < C / 16; ++k) tc_mma_bf16(tmem_base + U_COL, desc_a1 + 2 * k, desc_w4 + 2 * k, idesc_n128, k != 0 ? 1u : 0u);
        tc_commit(&u_full);
        mbar_wait(&a2_full, it & 1);      // EPI2: the gated tile is in A1
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < C / 16; ++k) tc_mma_bf16(tmem_base + Y_COL + s * C, desc_a1 + 2 * k, desc_w5 + 2 * k, idesc_n64, 1u);
        tc_commit(&y2_full);
        // G0 of the next tile into the other Y buffer (its last reader, EPI3 of tile it - 1, precedes this tile's a1_full arrivals)
        if (tile + (int)gridDim.x < num_tiles) g0(s ^ 1, it + 1);
      }
    }
  } else {
    // ================= epilogue warps: thread = pixel = TMEM lane =================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int r7 = row & 7;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    uint8_t* zst = smem + OFF_Z + (warp - 2) * Z_WARP_BYTES;
    uint8_t* a1row = smem + OFF_A1 + row * 128;
    const uint8_t* xrow = smem + OFF_X + row * 128;
    const int sw3 = (lane >> 1) & 3;
    if (lane == 0) { tma_prefetch_desc(&tm.O32); if (a.has_bf16) tma_prefetch_desc(&tm.O16); }
    const float inv_c = 1.0f / (float)C;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      int b, y0, x0;
      tile_coords(tile, b, y0, x0);
      const uint32_t ty = tmem_base + Y_COL + s * C + lane_addr;
      uint32_t raw[4][16];
      // ---------------- EPI1 ----------------
      mbar_wait(&g0_full[s], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_ld16(ty + q * 16, raw[q]);
      mbar_wait(&x_full, it & 1);
      tc_wait_ld();
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {      // float4 j = channels 4 j .. 4 j + 3
        const float4 xv = *reinterpret_cast<const float4*>(xrow + (j >> 3) * (X_BYTES / 2) + (((j & 7) ^ r7) << 4));
        const float4 bb = *reinterpret_cast<const float4*>(sP + P_B3 + 4 * j);
        uint32_t (&rw)[16] = raw[j >> 2];
        const int o = (j & 3) * 4;
        const float v0 = xv.x + (__uint_as_float(rw[o]) + bb.x), v1 = xv.y + (__uint_as_float(rw[o + 1]) + bb.y);
        const float v2 = xv.z + (__uint_as_float(rw[o + 2]) + bb.z), v3 = xv.w + (__uint_as_float(rw[o + 3]) + bb.w);
        s1 += (v0 + v1) + (v2 + v3);
        s2 += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
        rw[o] = __float_as_uint(v0); rw[o + 1] = __float_as_uint(v1); rw[o + 2] = __float_as_uint(v2); rw[o + 3] = __float_as_uint(v3);
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&x_empty);
        tma_store_wait_read0();      // the previous tile's stores have read this warp's staging block and its rows of the A1 tile
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_st16(ty + q * 16, raw[q]);
      {
        const float mean = s1 * inv_c;
        const float var = fmaxf(s2 * inv_c - mean * mean, 0.f);
        const float rstd = rsqrtf(var + a.eps);
        const float nmr = -mean * rstd;
#pragma unroll
        for (int q = 0; q < 8; ++q) {      // 16-byte chunk q = channels 8 q .. 8 q + 7
          uint32_t pk[4];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 g = *reinterpret_cast<const float4*>(sP + P_G2 + 8 * q + 4 * h);
            const float4 be = *reinterpret_cast<const float4*>(sP + P_BE2 + 8 * q + 4 * h);
            const uint32_t (&rw)[16] = raw[q >> 1];
            const int o = (q & 1) * 8 + 4 * h;
            pk[2 * h] = pack_bf16(fmaf(fmaf(__uint_as_float(rw[o]), rstd, nmr), g.x, be.x), fmaf(fmaf(__uint_as_float(rw[o + 1]), rstd, nmr), g.y, be.y));
            pk[2 * h + 1] = pack_bf16(fmaf(fmaf(__uint_as_float(rw[o + 2]), rstd, nmr), g.z, be.z), fmaf(fmaf(__uint_as_float(rw[o + 3]), rstd, nmr), g.w, be.w));
          }
          *reinterpret_cast<uint4*>(a1row + ((q ^ r7) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      fence_proxy_async_smem();
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a1_full);
      // ---------------- EPI2: SimpleGate ----------------
      mbar_wait(&u_full, it & 1);
      tc_fence_after();
      {
        const uint32_t tu = tmem_base + U_COL + lane_addr;
        tmem_ld16(tu, raw[0]);
        tmem_ld16(tu + C, raw[1]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {      // channels 16 q .. 16 q + 15 and their partners 64 + ...
          tc_wait_ld();
          if (q + 1 < 4) {
            tmem_ld16(tu + (q + 1) * 16, raw[2 * ((q + 1) & 1)]);
            tmem_ld16(tu + C + (q + 1) * 16, raw[2 * ((q + 1) & 1) + 1]);
          }
          const uint32_t (&ua)[16] = raw[2 * (q & 1)];
          const uint32_t (&ub)[16] = raw[2 * (q & 1) + 1];
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 ba = *reinterpret_cast<const float4*>(sP + P_B4 + 16 * q + 4 * i);
            const float4 bb = *reinterpret_cast<const float4*>(sP + P_B4 + C + 16 * q + 4 * i);
            const float g0v = (__uint_as_float(ua[4 * i]) + ba.x) * (__uint_as_float(ub[4 * i]) + bb.x);
            const float g1v = (__uint_as_float(ua[4 * i + 1]) + ba.y) * (__uint_as_float(ub[4 * i + 1]) + bb.y);
            const float g2v = (__uint_as_float(ua[4 * i + 2]) + ba.z) * (__uint_as_float(ub[4 * i + 2]) + bb.z);
            const float g3v = (__uint_as_float(ua[4 * i + 3]) + ba.w) * (__uint_as_float(ub[4 * i + 3]) + bb.w);
            pk[2 * i] = pack_bf16(g0v, g1v);
            pk[2 * i + 1] = pack_bf16(g2v, g3v);
          }
          *reinterpret_cast<uint4*>(a1row + (((2 * q) ^ r7) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(a1row + (((2 * q + 1) ^ r7) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a2_full);
      // ---------------- EPI3 ----------------
      mbar_wait(&y2_full, it & 1);
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_ld16(ty + q * 16, raw[q]);
      tc_wait_ld();
      tc_fence_before();
      s1 = 0.f; s2 = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {      // fp32 sub-block q = channels 16 q .. 16 q + 15
        uint32_t (&rw)[16] = raw[q];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 bb = *reinterpret_cast<const float4*>(sP + P_B5 + 16 * q + 4 * c);
          const float v0 = __uint_as_float(rw[4 * c]) + bb.x, v1 = __uint_as_float(rw[4 * c + 1]) + bb.y;
          const float v2 = __uint_as_float(rw[4 * c + 2]) + bb.z, v3 = __uint_as_float(rw[4 * c + 3]) + bb.w;
          s1 += (v0 + v1) + (v2 + v3);
          s2 += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
          rw[4 * c] = __float_as_uint(v0); rw[4 * c + 1] = __float_as_uint(v1); rw[4 * c + 2] = __float_as_uint(v2); rw[4 * c + 3] = __float_as_uint(v3);
        }
        if (q > 0) {
          if (lane == 0) tma_store_wait_read0();
          __syncwarp();
        }
        uint8_t* frow = zst + lane * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(frow + ((c ^ sw3) << 4)) = make_uint4(rw[4 * c], rw[4 * c + 1], rw[4 * c + 2], rw[4 * c + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tm.O32, zst, q * 16, x0, y0 + quad * 2, b);
          tma_store_commit();
        }
      }
      if (a.has_bf16) {
        float rstd = 1.f, nmr = 0.f;
        if (a.has_ln) {
          const float mean = s1 * inv_c;
          const float var = fmaxf(s2 * inv_c - mean * mean, 0.f);
          rstd = rsqrtf(var + a.eps);
          nmr = -mean * rstd;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {      // (has_ln = 0: gamma = 1, beta = 0, rstd = 1, nmr = 0 -> a plain bf16 copy of z)
          uint32_t pk[4];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 g = *reinterpret_cast<const float4*>(sP + P_LNG + 8 * q + 4 * h);
            const float4 be = *reinterpret_cast<const float4*>(sP + P_LNB + 8 * q + 4 * h);
            const uint32_t (&rw)[16] = raw[q >> 1];
            const int o = (q & 1) * 8 + 4 * h;
            pk[2 * h] = pack_bf16(fmaf(fmaf(__uint_as_float(rw[o]), rstd, nmr), g.x, be.x), fmaf(fmaf(__uint_as_float(rw[o + 1]), rstd, nmr), g.y, be.y));
            pk[2 * h + 1] = pack_bf16(fmaf(fmaf(__uint_as_float(rw[o + 2]), rstd, nmr), g.z, be.z), fmaf(fmaf(__uint_as_float(rw[o + 3]), rstd, nmr), g.w, be.w));
          }
          *reinterpret_cast<uint4*>(a1row + ((q ^ r7) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);      // (G2 has drained the A1 tile: y2_full)
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tm.O16, smem + OFF_A1 + quad * 4096, 0, x0, y0 + quad * 2, b);
          tma_store_commit();
        }
      }
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

extern "C" int ff_naf_tail(const FFNafTail* pp, void* stream) {
  FF_CHECK_ARG(pp != nullptr, "ff_naf_tail: null params");
  const FFNafTail& p = *pp;
  FF_CHECK_ARG(p.a0 && p.w3 && p.b3 && p.res && p.ln2_gamma && p.ln2_beta && p.w4 && p.b4 && p.w5 && p.b5 && p.x, "ff_naf_tail: null buffer");
  FF_CHECK_ARG(p.B > 0 && p.H > 0 && p.W > 0, "ff_naf_tail: bad size");
  FF_CHECK_ARG(p.a0_ld % 8 == 0 && p.a0_ld >= C && p.res_ld % 4 == 0 && p.res_ld >= C && p.x_ld % 4 == 0 && p.x_ld >= C,
               "ff_naf_tail: bad pitches (a0_ld=%d res_ld=%d x_ld=%d)", p.a0_ld, p.res_ld, p.x_ld);
  FF_CHECK_ARG(p.w3_batch_rows == 0 || p.w3_batch_rows == C, "ff_naf_tail: w3_batch_rows must be 0 or %d", C);
  FF_CHECK_ARG(p.ln_eps > 0.f, "ff_naf_tail: bad LayerNorm eps");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  FF_CHECK_ARG(al16(p.a0) && al16(p.w3) && al16(p.b3) && al16(p.res) && al16(p.ln2_gamma) && al16(p.ln2_beta) && al16(p.w4) && al16(p.b4) && al16(p.w5) &&
                   al16(p.b5) && al16(p.x), "ff_naf_tail: operands must be 16-byte aligned");
  if (p.out_bf16) FF_CHECK_ARG(al16(p.out_bf16) && p.out_ld % 8 == 0 && p.out_ld >= C, "ff_naf_tail: bad out_bf16 / out_ld");
  if (p.ln_gamma || p.ln_beta) FF_CHECK_ARG(p.ln_gamma && p.ln_beta && p.out_bf16, "ff_naf_tail: the next LayerNorm needs gamma, beta and out_bf16");
  EncodeTiledFn enc = get_encode();
  if (!enc) { ff_set_error("ff_naf_tail: cuTensorMapEncodeTiled entry point unavailable"); return FF_ERR_DRIVER; }
  NafMaps tm;
  auto img_map = [&](CUtensorMap* m, const void* ptr, int ld, int esz, CUtensorMapDataType dt, CUtensorMapSwizzle sw, int box_c, int box_h) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    cuuint64_t strides[3] = {(cuuint64_t)ld * esz, (cuuint64_t)ld * esz * p.W, (cuuint64_t)ld * esz * p.W * p.H};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)TW_, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  auto w_map = [&](CUtensorMap* m, const void* ptr, int rows, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  const CUtensorMapDataType BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, F32T = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  bool ok = img_map(&tm.A0, p.a0, p.a0_ld, 2, BF, CU_TENSOR_MAP_SWIZZLE_128B, C, TH_) &&
            w_map(&tm.W3, p.w3, p.w3_batch_rows ? p.B * C : C, C) && w_map(&tm.W4, p.w4, C2, C2) && w_map(&tm.W5, p.w5, C, C) &&
            img_map(&tm.R, p.res, p.res_ld, 4, F32T, CU_TENSOR_MAP_SWIZZLE_128B, 32, TH_) &&
            img_map(&tm.O32, p.x, p.x_ld, 4, F32T, CU_TENSOR_MAP_SWIZZLE_64B, 16, 2);
  tm.O16 = tm.A0;
  if (ok && p.out_bf16) ok = img_map(&tm.O16, p.out_bf16, p.out_ld, 2, BF, CU_TENSOR_MAP_SWIZZLE_128B, C, 2);
  FF_CHECK_ARG(ok, "ff_naf_tail: cuTensorMapEncodeTiled failed");
  Args a;
  a.B = p.B; a.H = p.H; a.W = p.W;
  a.tiles_x = ff_cdiv(p.W, TW_);
  a.tiles_per_img = a.tiles_x * ff_cdiv(p.H, TH_);
  a.m_tiles = a.tiles_per_img * p.B;
  a.w3_batch_rows = p.w3_batch_rows;
  a.b3 = p.b3; a.g2 = p.ln2_gamma; a.be2 = p.ln2_beta; a.b4 = p.b4; a.b5 = p.b5;
  a.lng = p.ln_gamma; a.lnb = p.ln_beta; a.eps = p.ln_eps;
  a.has_bf16 = p.out_bf16 ? 1 : 0;
  a.has_ln = p.ln_gamma ? 1 : 0;
  static FFPerDeviceFlag configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(naf_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { ff_set_error("ff_naf_tail: cudaFuncSetAttribute(%d) failed: %s", SMEM_BYTES, cudaGetErrorString(e)); return FF_ERR_CUDA; }
    configured = true;
  }
  const int cap = 2 * ff_num_sms();
  const int grid = a.m_tiles < cap ? a.m_tiles : cap;
  const cudaError_t le = ff_launch_pdl(naf_tail_kernel, dim3(grid), dim3(NTHREADS), SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream), tm, a);
  if (le != cudaSuccess) { ff_set_error("ff_naf_tail: launch failed: %s", cudaGetErrorString(le)); return FF_ERR_CUDA; }
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_naf_tail");
  return FF_OK;
}
