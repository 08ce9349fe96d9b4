// ff_conv_gemm: implicit-GEMM convolution / linear layer for sm_100a.
//   TMA (4-D shifted boxes, hardware zero fill)  ->  128B-swizzled smem ring
//   -> tcgen05.mma (M=128, N=BN, K=16, bf16 x bf16 -> fp32 in TMEM, double-buffered accumulators)
//   -> tcgen05.ld epilogue (bias / activation / gates / residual / PixelShuffle store).
// Persistent CTAs (one per SM), warp-specialised: warp0 = TMA producer, warp1 = TMEM alloc + MMA
// issuer, warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4).
#include "ff_common.cuh"
#include "../../include/ffb200.h"

namespace {

constexpr int TILE_M = 128;   // output pixels per tile (8 rows x 16 cols)
constexpr int TILE_W = 16;
constexpr int TILE_H = 8;
constexpr int BLOCK_K = 64;   // bf16 elements per k-block = 128 B = one swizzle row
constexpr int A_STAGE_BYTES = TILE_M * BLOCK_K * 2;
constexpr int NUM_EPI_WARPS = 8;    // epilogue warps of every variant except the TMA-store ones (see Cfg::EPI_WARPS)
// HALO variant (3x3 convs with few output channels, where the A operand re-read per tap is the L2 bottleneck):
// output tile = 16 rows x 8 cols; per 64-channel chunk the producer loads three column-shifted halo slabs
// [18 rows][8 px][64 ch] (dx = -1, 0, +1) ONCE and the nine taps address them with a row offset (dy * 1024 B), so A
// moves 55 KB per chunk instead of 9 x 16 KB.  Each slab is a standard 128B-swizzled K-major operand (8-pixel row =
// one 1024-B swizzle atom), so the UMMA descriptor is the same as in the plain variant.
constexpr int HALO_TW = 8, HALO_TH = 16;
constexpr int HALO_SLAB_BYTES = 18 * 8 * 128;
constexpr int HALO_A_BYTES = 3 * HALO_SLAB_BYTES;
constexpr int HALO_A_STAGES = 2;

struct Args {
  FFConvGemm p;
  int Ho, Wo;
  int Hc, Wc;   // geometry of the direct-store outputs (== Ho, Wo unless out_crop_h / out_crop_w crop the bottom / right edge)
  int tiles_x, tiles_per_img, m_tiles, n_tiles;
  int ntaps, cchunks;
  int cchunks1;   // k-blocks taken from x (the rest, up to cchunks, come from the second operand tensor x2: K-concatenated 1x1 layers)
  int vec_ok;   // every epilogue operand allows 16-byte fp32 / 8-byte bf16 vector access
};

// EPI selects the epilogue at compile time:
//   EPI_GENERIC    -- every fused operand, runtime flags, LSU loads/stores after an smem transpose
//   EPI_STORE      -- out_bf16 = acc + bias        } rows converted to bf16, staged in 64B-swizzled smem, one TMA store per
//   EPI_STORE_GELU -- out_bf16 = gelu(acc + bias)  } (warp, 32-column block)
//   EPI_RES        -- out_f32 = (acc + bias) * (alpha * col_scale) + res_f32 (+ bf16 copy): the fp32 residual block is
//                     TMA-loaded (prefetched one block ahead) into 128B-swizzled smem, updated in place and TMA-stored
//   EPI_RES_AUX    -- EPI_RES + aux_alpha * aux_bf16[p,n] * aux_chan[b,n] (HAT: x = shortcut + proj(attn) + 0.01 * cab * se); no bf16 copy
//   EPI_STORE_GATE -- SimpleGate folded: 64 accumulator columns -> 32 bf16 outputs, TMA store
//   EPI_NARROW     -- n_store <= 4 (image-space outputs): one accumulator row per thread straight from TMEM to global, fp32 / bf16
//                     residual prefetched before the accumulator wait (BN = 16 only)
//   EPI_OPS1..3    -- bf16 output with up to three bf16 operand tiles (res / mul / aux) of the same [pixels, n] shape:
//                     out = post(act(acc + bias) * alpha * col_scale * mul + aux_alpha * aux * aux_chan + res); the operand
//                     blocks are TMA-loaded one item ahead into 64B-swizzled smem, the result leaves through a TMA store
enum { EPI_GENERIC = 0, EPI_STORE = 1, EPI_STORE_GELU = 2, EPI_RES = 3, EPI_RES_AUX = 4, EPI_STORE_GATE = 5, EPI_NARROW = 6,
       EPI_OPS1 = 7, EPI_OPS2 = 8, EPI_OPS3 = 9 };

template <int BN, int EPI, bool HALO = false>
struct Cfg {
  static constexpr int TW = HALO ? HALO_TW : TILE_W;              // output tile geometry (TW x TH = 128 pixels)
  static constexpr int TH = HALO ? HALO_TH : TILE_H;
  static constexpr int QR = 32 / TW;                              // tile rows per TMEM lane quadrant
  static constexpr int B_STAGE_BYTES = BN * BLOCK_K * 2;
  // HALO: one ring stage holds the weights of a whole kernel ROW (three taps) of a 64-channel chunk, so the MMA warp issues 12
  // instructions per barrier wait instead of 4 (N <= 64 instructions are issue bound: the wait + fence + election between
  // taps cost as much as the four MMAs)
  static constexpr int HALO_TAPS = (EPI == 3 || EPI == 4) ? 1 : 3;      // (the residual epilogues' staging leaves room for single-tap stages only)
  static constexpr int STAGE_BYTES = HALO ? HALO_TAPS * B_STAGE_BYTES : A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int HALO_BYTES = HALO ? HALO_A_STAGES * HALO_A_BYTES : 0;
  static constexpr int CB = BN < 32 ? BN : 32;                    // epilogue column block
  static constexpr int STG_PITCH = CB + 4;                        // floats; +4 keeps float4 accesses conflict-free (generic path)
  static constexpr int STG_WARP_BYTES = EPI == EPI_RES ? 10240 : EPI == EPI_RES_AUX ? 12288 : EPI == EPI_NARROW ? 0 : EPI >= EPI_OPS1 ? (2 * (EPI - EPI_OPS1 + 1) + 2) * 2048 : (EPI == EPI_GENERIC ? 32 * STG_PITCH * 4 : 4096);
  // The plain TMA-store epilogues are latency bound per warp (tcgen05.ld -> cvt -> st.shared -> proxy fence -> TMA store is one
  // dependent chain per 32-column block), so they run four warps per TMEM lane quadrant instead of two.
  static constexpr int EPI_WARPS = ((EPI == 1 || EPI == 2) && !HALO) ? 16 : NUM_EPI_WARPS;   // HALO tiles hold <= 2 column blocks; the smaller staging area leaves room for a 9-stage weight ring (417 vs 463 us on the HR 64->64 layers)
  static constexpr int STG_BYTES = EPI_WARPS * STG_WARP_BYTES;
  static constexpr int STAGES_RAW = (225 * 1024 - STG_BYTES - HALO_BYTES) / STAGE_BYTES;
  static constexpr int STAGES_MAX = HALO ? (HALO_TAPS == 3 ? 4 : 9) : 6;
  static constexpr int STAGES = STAGES_RAW > STAGES_MAX ? STAGES_MAX : STAGES_RAW;
  static constexpr int RING_BYTES = HALO_BYTES + STAGES * STAGE_BYTES;        // [halo A stages][ring]; staging follows
  static constexpr int SMEM_BYTES = RING_BYTES + STG_BYTES + 1024;            // + alignment slack
  static constexpr int THREADS = 64 + 32 * EPI_WARPS + (HALO ? 32 : 0);       // HALO: the last warp = A-slab producer
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
};

// erf-form GELU with the Abramowitz-Stegun 7.1.28 approximation erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16 (|err| <= 3e-7,
// far below the bf16 rounding of the stored activations): 6 FMA + 4 FMUL + one MUFU.RCP, branch free.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = x * 0.70710678118654752440f;
  const float az = fabsf(z);
  float p = fmaf(az, 0.0000430638f, 0.0002765672f);
  p = fmaf(az, p, 0.0001520143f);
  p = fmaf(az, p, 0.0092705272f);
  p = fmaf(az, p, 0.0422820123f);
  p = fmaf(az, p, 0.0705230784f);
  p = fmaf(az, p, 1.0f);
  p = p * p; p = p * p; p = p * p; p = p * p;          // ^16 (saturates to +inf for |z| > ~9: erf -> 1, as it should)
  const float e = 1.0f - __fdividef(1.0f, p);
  return 0.5f * x * (1.0f + copysignf(e, z));
}

// Hot GELU epilogue (fc1 / conv + GELU -> bf16): tanh form on the hardware tanh unit, 5 FP32 ops + 1 MUFU.
// |gelu_tanh - gelu_erf| <= 4.8e-4 and tanh.approx adds <= 2.5e-4*|x|: both far below the bf16 rounding of the stored
// activation (half ulp = 3.9e-3*|y|).  Measured effect on the HAT-L output: 3.4e-5 max-abs (tolerance 2e-2).
__device__ __forceinline__ float gelu_tanh_hw(float x) {
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case FF_ACT_GELU: return gelu_fast(v);
    case FF_ACT_RELU: return fmaxf(v, 0.f);
    case FF_ACT_LRELU: return v > 0.f ? v : 0.01f * v;
    case FF_ACT_SIGMOID: return sigmoidf_(v);
    case FF_ACT_CLAMP01: return fminf(fmaxf(v, 0.f), 1.f);
    default: return v;
  }
}

// one uniform switch for a vector of values (the per-element switch of apply_act costs a branch per value)
template <int N>
__device__ __forceinline__ void apply_act_n(float (&v)[N], int act) {
  switch (act) {
    case FF_ACT_GELU:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = gelu_fast(v[i]);
      break;
    case FF_ACT_RELU:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = fmaxf(v[i], 0.f);
      break;
    case FF_ACT_LRELU:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = v[i] > 0.f ? v[i] : 0.01f * v[i];
      break;
    case FF_ACT_SIGMOID:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = sigmoidf_(v[i]);
      break;
    case FF_ACT_CLAMP01:
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = fminf(fmaxf(v[i], 0.f), 1.f);
      break;
    default: break;
  }
}

__device__ __forceinline__ void load_bf16x16(const bf16* ptr, float (&o)[16]) {
  const uint4* q = reinterpret_cast<const uint4*>(ptr);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    o[2 * i] = __uint_as_float(w[i] << 16);
    o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// ----------------------------------------------------------------------------------------------
// Fused epilogue.
//  * scalar flavour (`epilogue16`): 16 consecutive channels of one pixel -- used by the SIMT debug kernel and
//    as the semantic reference of the vector flavour;
//  * vector flavour (`epilogue_vec4`): 4 consecutive channels of one pixel, 16-byte fp32 / 8-byte bf16 accesses --
//    used by the tcgen05 kernel after the accumulator tile has been transposed through shared memory so that
//    consecutive lanes touch consecutive channels (coalesced global traffic).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void out_location(const Args& a, int b, int oy, int ox, int n0, long long& opix, int& oc) {
  const FFConvGemm& p = a.p;
  if (p.pixel_shuffle == 2) {
    const int cq = p.n_store >> 2;
    const int sub = n0 / cq;
    oc = n0 - sub * cq;
    opix = ((long long)(b * 2 * a.Ho + 2 * oy + (sub >> 1))) * (2 * a.Wo) + 2 * ox + (sub & 1);
  } else {
    oc = n0;
    opix = ((long long)(b * a.Ho + oy)) * a.Wo + ox;
  }
}

__device__ __forceinline__ void epilogue16(const Args& a, float (&v)[16], int b, int oy, int ox, int n0) {
  const FFConvGemm& p = a.p;
  if (n0 >= p.n_store) return;
  if (p.bias) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + n0 + i);
  }
  if (p.gate_pairs) {
    const long long gpix = ((long long)(b * a.Ho + oy)) * a.Wo + ox;
    bf16* q = reinterpret_cast<bf16*>(p.out_bf16) + gpix * p.out_ld + (n0 >> 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = __float2bfloat16_rn(v[i] * v[i + 8]);
    return;
  }
  long long opix;
  int oc;
  out_location(a, b, oy, ox, n0, opix, oc);
  const int width = (p.pixel_shuffle == 2) ? (p.n_store >> 2) : p.n_store;
  const int nvalid = min(16, width - oc);
  for (int i = 0; i < nvalid; ++i) {
    float x = apply_act(v[i], p.act) * p.alpha;
    if (p.col_scale) x *= __ldg(p.col_scale + n0 + i);
    if (p.mul) x *= __bfloat162float(reinterpret_cast<const bf16*>(p.mul)[opix * p.mul_ld + oc + i]);
    if (p.aux) {
      const float m = __bfloat162float(reinterpret_cast<const bf16*>(p.aux)[opix * p.aux_ld + oc + i]);
      x += p.aux_alpha * m * (p.aux_chan ? __ldg(p.aux_chan + (long long)b * p.aux_chan_ld + oc + i) : 1.f);
    }
    if (p.res) {
      x += p.res_is_f32 ? reinterpret_cast<const float*>(p.res)[opix * p.res_ld + oc + i]
                        : __bfloat162float(reinterpret_cast<const bf16*>(p.res)[opix * p.res_ld + oc + i]);
    }
    x = apply_act(x, p.post_act);
    if (p.out_f32) p.out_f32[opix * p.out_f32_ld + oc + i] = x;
    if (p.out_bf16) reinterpret_cast<bf16*>(p.out_bf16)[opix * p.out_ld + oc + i] = __float2bfloat16_rn(x);
  }
}

struct Vec4Operands {   // global operands of one (pixel, 4 channels) item, loaded ahead of the math
  float4 res;
  uint2 mul, aux;
};

__device__ __forceinline__ void bf16x4_to_float(const uint2& q, float (&f)[4]) {
  f[0] = __uint_as_float(q.x << 16); f[1] = __uint_as_float(q.x & 0xffff0000u);
  f[2] = __uint_as_float(q.y << 16); f[3] = __uint_as_float(q.y & 0xffff0000u);
}

__device__ __forceinline__ void vec4_load(const FFConvGemm& p, long long opix, int oc, Vec4Operands& o) {
  if (p.res) {
    if (p.res_is_f32) {
      o.res = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + opix * p.res_ld + oc));
    } else {
      const uint2 q = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(p.res) + opix * p.res_ld + oc));
      float f[4];
      bf16x4_to_float(q, f);
      o.res = make_float4(f[0], f[1], f[2], f[3]);
    }
  }
  if (p.mul) o.mul = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(p.mul) + opix * p.mul_ld + oc));
  if (p.aux) o.aux = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(p.aux) + opix * p.aux_ld + oc));
}

// bias / col_scale / aux_chan for the 4 channels are loop-invariant per column block and passed in registers
__device__ __forceinline__ void vec4_finish(const FFConvGemm& p, float4 acc, const Vec4Operands& o, const float (&bias)[4],
                                            const float (&cscale)[4], const float (&achan)[4], long long opix, int oc) {
  float v[4] = {acc.x + bias[0], acc.y + bias[1], acc.z + bias[2], acc.w + bias[3]};
  if (p.act) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = apply_act(v[i], p.act);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] *= cscale[i];    // alpha * col_scale folded by the caller
  if (p.mul) {
    float m[4];
    bf16x4_to_float(o.mul, m);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] *= m[i];
  }
  if (p.aux) {
    float m[4];
    bf16x4_to_float(o.aux, m);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] += p.aux_alpha * m[i] * achan[i];
  }
  if (p.res) { v[0] += o.res.x; v[1] += o.res.y; v[2] += o.res.z; v[3] += o.res.w; }
  if (p.post_act) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = apply_act(v[i], p.post_act);
  }
  if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + opix * p.out_f32_ld + oc) = make_float4(v[0], v[1], v[2], v[3]);
  if (p.out_bf16) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out_bf16) + opix * p.out_ld + oc) =
        make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

// ----------------------------------------------------------------------------------------------
// tcgen05 kernel
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// One [rows x pixels x 32 channels] epilogue block of the output-shaped tensors.  With PixelShuffle(2) folded into the layer the
// tensor map is 5-D (channel, sub-x, x, sub-y, image row): GEMM column n = sub * cq + c lands at pixel (2y + sub / 2, 2x + sub % 2).
__device__ __forceinline__ void tma_store_blk(const CUtensorMap* m, const void* src, const Args& a, int n_blk, int x, int y, int b) {
  if (a.p.pixel_shuffle) {
    const int cq = a.p.n_store >> 2, sub = n_blk / cq;
    tma_store_5d(m, src, n_blk - sub * cq, sub & 1, x, sub >> 1, b * a.Ho + y);
  } else {
    tma_store_4d(m, src, n_blk, x, y, b);
  }
}
__device__ __forceinline__ void tma_load_blk(void* dst, const CUtensorMap* m, uint64_t* bar, const Args& a, int n_blk, int x, int y, int b) {
  if (a.p.pixel_shuffle) {
    const int cq = a.p.n_store >> 2, sub = n_blk / cq;
    tma_load_5d(dst, m, bar, n_blk - sub * cq, sub & 1, x, sub >> 1, b * a.Ho + y);
  } else {
    tma_load_4d(dst, m, bar, n_blk, x, y, b);
  }
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int BN, int EPI, bool HALO>
__global__ void __launch_bounds__(Cfg<BN, EPI, HALO>::THREADS, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
                    const __grid_constant__ CUtensorMap tmO32, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ Args a) {
  using C = Cfg<BN, EPI, HALO>;
  constexpr int TW = C::TW, TH = C::TH, QR = C::QR;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[C::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[C::STAGES];
  __shared__ __align__(8) uint64_t halo_full[HALO_A_STAGES];
  __shared__ __align__(8) uint64_t halo_empty[HALO_A_STAGES];
  __shared__ __align__(8) uint64_t tmem_full[2];
  __shared__ __align__(8) uint64_t tmem_empty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(8) uint64_t res_bar[NUM_EPI_WARPS][2];
  __shared__ float2 ln_part[TILE_M];            // fused LayerNorm: per-row exchange slot between the two warps of a lane quadrant (1 KB:
                                                // the largest RES configuration leaves 2 KB of static shared memory)

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = a.m_tiles * a.n_tiles;
  const int kblocks = a.ntaps * a.cchunks;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int w = 0; w < NUM_EPI_WARPS; ++w) { mbar_init(&res_bar[w][0], 1); mbar_init(&res_bar[w][1], 1); }
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], C::EPI_WARPS);
    }
    for (int s = 0; s < HALO_A_STAGES; ++s) {
      mbar_init(&halo_full[s], 1);
      mbar_init(&halo_empty[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();      // everything above touches no tensor of the forward; from here on the predecessor's writes are visible
  const uint32_t tmem_base = tmem_base_smem;

  // The producer and MMA warps stay converged: every lane runs the loops and polls the barriers, one elected lane issues
  // the TMA / tcgen05 instructions.  (Issuing from a divergent `lane == 0` branch makes ptxas wrap every tcgen05.mma in a
  // per-thread election loop and recompute the descriptors through a long dependent chain: ~150 cycles per MMA instead of
  // the ~50 of back-to-back UTCHMMA, which capped every N <= 192 GEMM.)
  if (warp == 0) {
    // ================= TMA producer =================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
      const int b = m_tile / a.tiles_per_img;
      const int t = m_tile - b * a.tiles_per_img;
      const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
      const int y0 = ty * TH, x0 = tx * TW;
      const int brow = b * a.p.w_batch_rows + n_tile * BN;
      if constexpr (HALO) {
        // weights only, chunk-major k order; the A slabs come from the last warp
        uint8_t* ring = smem + C::HALO_BYTES;
        for (int cc = 0; cc < a.cchunks; ++cc) {
          for (int g = 0; g < 9 / C::HALO_TAPS; ++g) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
#pragma unroll
              for (int j = 0; j < C::HALO_TAPS; ++j)
                tma_load_2d(ring + stage * C::STAGE_BYTES + j * C::B_STAGE_BYTES, &tmB, &full_bar[stage], ((g * C::HALO_TAPS + j) * a.cchunks + cc) * BLOCK_K, brow);
            }
            __syncwarp();
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      } else {
        int tap = 0, cc = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          int cx, cy;
          if (a.p.kind == FF_CONV_3X3) {
            const int dy = (tap * 11) >> 5;       // tap / 3 for tap in [0, 9)
            cx = x0 + (tap - 3 * dy) - 1;
            cy = y0 + dy - 1;
          } else if (a.p.kind == FF_CONV_2X2S2) {
            cx = 2 * x0 + (tap & 1);
            cy = 2 * y0 + (tap >> 1);
          } else {
            cx = x0;
            cy = y0;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + stage * C::STAGE_BYTES;
            mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            if (cc < a.cchunks1) tma_load_4d(sa, &tmA, &full_bar[stage], cc * BLOCK_K, cx, cy, b);
            else tma_load_4d(sa, &tmA2, &full_bar[stage], (cc - a.cchunks1) * BLOCK_K, cx, cy, b);      // second K segment (x2)
            tma_load_2d(sa + A_STAGE_BYTES, &tmB, &full_bar[stage], kb * BLOCK_K, brow);
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          if (++cc == a.cchunks) { cc = 0; ++tap; }
        }
      }
    }
  } else if (HALO && warp == 2 + C::EPI_WARPS) {
    // ================= HALO A-slab producer =================
    int hs = 0;
    uint32_t hphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / a.n_tiles;
      const int b = m_tile / a.tiles_per_img;
      const int t = m_tile - b * a.tiles_per_img;
      const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
      for (int cc = 0; cc < a.cchunks; ++cc) {
        mbar_wait(&halo_empty[hs], hphase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + hs * HALO_A_BYTES;
          mbar_arrive_expect_tx(&halo_full[hs], HALO_A_BYTES);
#pragma unroll
          for (int j = 0; j < 3; ++j)
            tma_load_4d(sa + j * HALO_SLAB_BYTES, &tmA, &halo_full[hs], cc * BLOCK_K, tx * TW - 1 + j, ty * TH - 1, b);
        }
        __syncwarp();
        if (++hs == HALO_A_STAGES) { hs = 0; hphase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN);
    // descriptors advance by plain adds on the 14-bit (address >> 4) field: +2 per 16-element k step, +bytes/16 per stage / slab
    const uint64_t desc_ring = umma_desc_k_sw128(smem_u32(smem + C::HALO_BYTES));
    const uint64_t desc_halo = umma_desc_k_sw128(smem_u32(smem));
    (void)desc_halo;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int hs = 0;
    uint32_t hphase = 0;
    (void)hs; (void)hphase;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      if constexpr (HALO) {
        for (int cc = 0; cc < a.cchunks; ++cc) {
          mbar_wait(&halo_full[hs], hphase);
          const uint64_t dslab = desc_halo + (uint64_t)(hs * (HALO_A_BYTES >> 4));
#pragma unroll
          for (int g = 0; g < 9 / C::HALO_TAPS; ++g) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int j = 0; j < C::HALO_TAPS; ++j) {
                const int tap = g * C::HALO_TAPS + j;
                const uint64_t da = dslab + (uint64_t)((tap % 3) * (HALO_SLAB_BYTES >> 4) + (tap / 3) * (1024 >> 4));   // dx slab, dy row offset
                const uint64_t db = desc_ring + (uint64_t)((stage * C::STAGE_BYTES + j * C::B_STAGE_BYTES) >> 4);
#pragma unroll
                for (int k = 0; k < BLOCK_K / 16; ++k) tc_mma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (cc | tap | k) != 0 ? 1u : 0u);
              }
              tc_commit(&empty_bar[stage]);
              if (g == 9 / C::HALO_TAPS - 1) tc_commit(&halo_empty[hs]);
            }
            __syncwarp();
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
          if (++hs == HALO_A_STAGES) { hs = 0; hphase ^= 1; }
        }
      } else {
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = desc_ring + (uint64_t)(stage * (C::STAGE_BYTES >> 4));
            const uint64_t db = da + (A_STAGE_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) tc_mma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            tc_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (elect_one()) tc_commit(&tmem_full[acc]);
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================= epilogue warps =================
    if constexpr (EPI == EPI_RES || EPI == EPI_RES_AUX) {
      constexpr bool AUX = (EPI == EPI_RES_AUX);   // layout per warp: [R0 4K][R1 4K][O16 2K] or [R0 4K][R1 4K][X0 2K][X1 2K]
      // ---------- residual epilogue: TMA-load res block -> in-place update in 128B-swizzled smem -> TMA store ----------
      static_assert(C::CB == 32, "EPI_RES needs 32-column blocks");
      const FFConvGemm& p = a.p;
      const int ew = warp - 2;
      const int quad = warp & 3;
      const int half = ew >> 2;
      uint8_t* wbase = smem + C::RING_BYTES + ew * C::STG_WARP_BYTES;   // [R0 4K][R1 4K][O16 2K], 1024-B aligned
      const int ncb = BN / 32;
      int buf = 0;
      uint32_t ph[2] = {0, 0};
      int acc = 0;
      uint32_t acc_phase = 0;
      if (lane == 0) { tma_prefetch_desc(&tmR); tma_prefetch_desc(&tmO32); if (AUX || p.out_bf16) tma_prefetch_desc(&tmO); if (p.ln_out) tma_prefetch_desc(&tmX); }
      // items of this warp: (tile, cb) with cb = half, half+2, ... while the block starts below n_store.  The residual
      // block of the NEXT item (possibly in the next tile) is prefetched while the current one is processed.
      auto valid = [&](int tl, int c) { return c < ncb && (tl % a.n_tiles) * BN + c * 32 < p.n_store; };
      auto issue_load = [&](int tl, int c, int bsel) {
        const int m_tile = tl / a.n_tiles, n_tile = tl - m_tile * a.n_tiles;
        const int b = m_tile / a.tiles_per_img;
        const int t = m_tile - b * a.tiles_per_img;
        const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
        mbar_arrive_expect_tx(&res_bar[ew][bsel], AUX ? 6144 : 4096);
        tma_load_blk(wbase + bsel * 4096, &tmR, &res_bar[ew][bsel], a, n_tile * BN + c * 32, tx * TW, ty * TH + quad * QR, b);
        if constexpr (AUX) tma_load_4d(wbase + 8192 + bsel * 2048, &tmO, &res_bar[ew][bsel], n_tile * BN + c * 32, tx * TW, ty * TH + quad * QR, b);
      };
      {
        int ft = blockIdx.x;
        while (ft < num_tiles && !valid(ft, half)) ft += gridDim.x;
        if (lane == 0 && ft < num_tiles) issue_load(ft, half, 0);
      }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
        const int b = m_tile / a.tiles_per_img;
        const int t = m_tile - b * a.tiles_per_img;
        const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quad * 32) << 16);
        float ln_s1 = 0.f, ln_s2 = 0.f;      // fused LayerNorm: this thread's row, over this warp's column blocks
#pragma unroll 1
        for (int cb = half; valid(tile, cb); cb += 2) {
          const int n_blk = n_tile * BN + cb * 32;
          // locate the next item and prefetch its residual block into the other buffer
          int ntile = tile, ncb2 = cb + 2;
          if (!valid(ntile, ncb2)) {
            ncb2 = half;
            ntile = tile + gridDim.x;
            while (ntile < num_tiles && !valid(ntile, ncb2)) ntile += gridDim.x;
          }
          if (lane == 0) {
            tma_store_wait_read<0>();      // the stores that read R[buf^1] / O16 (issued one item ago) have drained
            if (ntile < num_tiles) issue_load(ntile, ncb2, buf ^ 1);
          }
          uint32_t raw[32];
          tmem_ld16(taddr + cb * 32, *reinterpret_cast<uint32_t(*)[16]>(&raw[0]));
          tmem_ld16(taddr + cb * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&raw[16]));
          tc_wait_ld();
          mbar_wait(&res_bar[ew][buf], ph[buf]);
          ph[buf] ^= 1;
          __syncwarp();
          uint8_t* rrow = wbase + buf * 4096 + lane * 128;
          uint8_t* orow = wbase + 8192 + lane * 64;
          const int sw7 = lane & 7, sw3 = (lane >> 1) & 3;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f), cs = make_float4(p.alpha, p.alpha, p.alpha, p.alpha);
            if (p.bias) bb = __ldg(reinterpret_cast<const float4*>(p.bias + n_blk + c * 4));
            if (p.col_scale) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(p.col_scale + n_blk + c * 4));
              cs.x *= q.x; cs.y *= q.y; cs.z *= q.z; cs.w *= q.w;
            }
            float4* rp = reinterpret_cast<float4*>(rrow + ((c ^ sw7) << 4));
            float4 r = *rp;
            r.x = fmaf(__uint_as_float(raw[4 * c]) + bb.x, cs.x, r.x);
            r.y = fmaf(__uint_as_float(raw[4 * c + 1]) + bb.y, cs.y, r.y);
            r.z = fmaf(__uint_as_float(raw[4 * c + 2]) + bb.z, cs.z, r.z);
            r.w = fmaf(__uint_as_float(raw[4 * c + 3]) + bb.w, cs.w, r.w);
            if constexpr (AUX) {
              const uint2 q = *reinterpret_cast<const uint2*>(wbase + 8192 + buf * 2048 + lane * 64 + (((c >> 1) ^ sw3) << 4) + ((c & 1) << 3));
              float4 ch = make_float4(p.aux_alpha, p.aux_alpha, p.aux_alpha, p.aux_alpha);
              if (p.aux_chan) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(p.aux_chan + (long long)b * p.aux_chan_ld + n_blk + c * 4));
                ch.x *= g.x; ch.y *= g.y; ch.z *= g.z; ch.w *= g.w;
              }
              r.x = fmaf(__uint_as_float(q.x << 16), ch.x, r.x);
              r.y = fmaf(__uint_as_float(q.x & 0xffff0000u), ch.y, r.y);
              r.z = fmaf(__uint_as_float(q.y << 16), ch.z, r.z);
              r.w = fmaf(__uint_as_float(q.y & 0xffff0000u), ch.w, r.w);
            }
            *rp = r;
            if (!AUX && p.out_bf16) {
              __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y), hi = __floats2bfloat162_rn(r.z, r.w);
              uint2* op = reinterpret_cast<uint2*>(orow + (((c >> 1) ^ sw3) << 4) + ((c & 1) << 3));
              *op = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            }
            if (p.ln_out) {      // keep the updated row for the normalisation pass (written back over the accumulator below)
              ln_s1 += (r.x + r.y) + (r.z + r.w);
              ln_s2 += (r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w);
              raw[4 * c] = __float_as_uint(r.x); raw[4 * c + 1] = __float_as_uint(r.y);
              raw[4 * c + 2] = __float_as_uint(r.z); raw[4 * c + 3] = __float_as_uint(r.w);
            }
          }
          if (p.ln_out) {
            tmem_st16(taddr + cb * 32, *reinterpret_cast<uint32_t(*)[16]>(&raw[0]));
            tmem_st16(taddr + cb * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&raw[16]));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_blk(&tmO32, wbase + buf * 4096, a, n_blk, tx * TW, ty * TH + quad * QR, b);
            if (!AUX && p.out_bf16) tma_store_blk(&tmO, wbase + 8192, a, n_blk, tx * TW, ty * TH + quad * QR, b);
            tma_store_commit();
          }
          buf ^= 1;
        }
        if (p.ln_out) {
          // ---------- fused LayerNorm of the updated rows: the n tile spans the whole row (host check), the two warps of a lane
          // quadrant hold its two halves.  Row statistics = both warps' partial sums: the second warp publishes its partial, the
          // first one combines and publishes (rstd, -mean*rstd) in the same smem slot (two 64-thread named barriers per tile);
          // the fp32 row is re-read from TMEM, normalised, and leaves as bf16 through a TMA store.
          float rstd, nmr;
          {
            float2* slot = &ln_part[quad * 32 + lane];
            if (half == 1) *slot = make_float2(ln_s1, ln_s2);
            named_bar_sync(1 + quad, 64);
            if (half == 0) {
              const float2 other = *slot;
              const float inv_c = 1.0f / (float)p.ln_cols;
              const float mean = (ln_s1 + other.x) * inv_c;
              const float var = fmaxf((ln_s2 + other.y) * inv_c - mean * mean, 0.f);
              rstd = rsqrtf(var + p.ln_eps);
              nmr = -mean * rstd;
              *slot = make_float2(rstd, nmr);
            }
            named_bar_sync(1 + quad, 64);
            if (half == 1) { const float2 st2 = *slot; rstd = st2.x; nmr = st2.y; }
          }
          if (lane == 0) tma_store_wait_read<0>();      // R[buf^1] / the bf16 staging of the last item have been read
          __syncwarp();
          tc_wait_st();
          const int sw3n = (lane >> 1) & 3;
          int sb = 0;
#pragma unroll 1
          for (int cb = half; valid(tile, cb); cb += 2) {
            const int n_blk = n_tile * BN + cb * 32;
            uint32_t xr[32];
            tmem_ld16(taddr + cb * 32, *reinterpret_cast<uint32_t(*)[16]>(&xr[0]));
            tmem_ld16(taddr + cb * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&xr[16]));
            uint8_t* stg = sb == 0 ? (wbase + (buf ^ 1) * 4096) : (AUX ? (wbase + 8192 + (buf ^ 1) * 2048) : (wbase + 8192));
            if (lane == 0) tma_store_wait_read<1>();     // the store that read this staging buffer two blocks ago
            __syncwarp();
            tc_wait_ld();
            uint8_t* srow = stg + lane * 64;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + n_blk + c * 4));
              const float4 be = __ldg(reinterpret_cast<const float4*>(p.ln_beta + n_blk + c * 4));
              const float y0 = fmaf(fmaf(__uint_as_float(xr[4 * c]), rstd, nmr), g.x, be.x);
              const float y1 = fmaf(fmaf(__uint_as_float(xr[4 * c + 1]), rstd, nmr), g.y, be.y);
              const float y2 = fmaf(fmaf(__uint_as_float(xr[4 * c + 2]), rstd, nmr), g.z, be.z);
              const float y3 = fmaf(fmaf(__uint_as_float(xr[4 * c + 3]), rstd, nmr), g.w, be.w);
              *reinterpret_cast<uint2*>(srow + (((c >> 1) ^ sw3n) << 4) + ((c & 1) << 3)) = make_uint2(pack_bf16(y0, y1), pack_bf16(y2, y3));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmX, stg, n_blk, tx * TW, ty * TH + quad * QR, b);
              tma_store_commit();
            }
            sb ^= 1;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (lane == 0) tma_store_wait_all();
    } else
    if constexpr (EPI == EPI_NARROW) {
      // ---------- narrow epilogue: <= 4 output channels, thread = pixel, no staging ----------
      static_assert(BN == 16, "EPI_NARROW is the BN = 16 tile");
      const FFConvGemm& p = a.p;
      const int quad = warp & 3;
      const bool worker = ((warp - 2) >> 2) == 0;      // the second warp of each quadrant only takes part in the TMEM handshake
      float bias_r[4], cs_r[4];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        bias_r[n] = (p.bias && n < p.n_store) ? __ldg(p.bias + n) : 0.f;
        cs_r[n] = p.alpha * ((p.col_scale && n < p.n_store) ? __ldg(p.col_scale + n) : 1.f);
      }
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / a.n_tiles;
        const int b = m_tile / a.tiles_per_img;
        const int t = m_tile - b * a.tiles_per_img;
        const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
        const int r = quad * 32 + lane;
        const int oy = ty * TH + r / TW, ox = tx * TW + r % TW;
        const bool inb = oy < a.Hc && ox < a.Wc;                  // partial edge tiles / cropped outputs
        const long long px = ((long long)(b * a.Ho + oy)) * a.Wo + ox;      // operand (full) geometry
        const long long pxo = ((long long)(b * a.Hc + oy)) * a.Wc + ox;     // output geometry
        float resv[4] = {0.f, 0.f, 0.f, 0.f};
        if (worker && inb && p.res) {      // requested before the accumulator wait: the round trip hides behind the main loop
#pragma unroll
          for (int n = 0; n < 4; ++n)
            if (n < p.n_store)
              resv[n] = p.res_is_f32 ? __ldg(reinterpret_cast<const float*>(p.res) + px * p.res_ld + n)
                                     : __bfloat162float(reinterpret_cast<const bf16*>(p.res)[px * p.res_ld + n]);
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        if (worker) {
          uint32_t raw[16];
          tmem_ld16(tmem_base + acc * BN + ((uint32_t)(quad * 32) << 16), raw);
          tc_wait_ld();
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            if (n < p.n_store) {
              float x = __uint_as_float(raw[n]) + bias_r[n];
              if (p.act) x = apply_act(x, p.act);
              x = fmaf(x, cs_r[n], resv[n]);
              if (p.post_act) x = apply_act(x, p.post_act);
              if (inb) {
                if (p.out_f32) p.out_f32[pxo * p.out_f32_ld + n] = x;
                if (p.out_bf16) reinterpret_cast<bf16*>(p.out_bf16)[pxo * p.out_ld + n] = __float2bfloat16_rn(x);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    } else
    if constexpr (EPI >= EPI_OPS1) {
      // ---------- bf16 operand epilogue: TMA-load res / mul / aux blocks (one item ahead) -> math -> bf16 -> TMA store ----------
      static_assert(C::CB == 32, "EPI_OPS needs 32-column blocks");
      constexpr int NOPS = EPI - EPI_OPS1 + 1;
      const FFConvGemm& p = a.p;
      const int ew = warp - 2;
      const int quad = warp & 3;
      const int half = ew >> 2;
      uint8_t* wbase = smem + C::RING_BYTES + ew * C::STG_WARP_BYTES;   // [slot0 b0|b1][slot1 b0|b1]..[out b0|b1], 2 KB each
      uint8_t* obase = wbase + NOPS * 4096;
      // operand slots in the order res, mul, aux (maps tmR, tmO32, tmX)
      const int s_res = 0, s_mul = p.res ? 1 : 0, s_aux = (p.res ? 1 : 0) + (p.mul ? 1 : 0);
      const uint32_t tx_bytes = 2048u * ((p.res ? 1 : 0) + (p.mul ? 1 : 0) + (p.aux ? 1 : 0));
      const int ncb = BN / 32;
      int buf = 0;
      uint32_t ph[2] = {0, 0};
      int acc = 0;
      uint32_t acc_phase = 0;
      if (lane == 0) { tma_prefetch_desc(&tmO); if (p.res) tma_prefetch_desc(&tmR); if (p.mul) tma_prefetch_desc(&tmO32); if (p.aux) tma_prefetch_desc(&tmX); }
      auto valid = [&](int tl, int c) { return c < ncb && (tl % a.n_tiles) * BN + c * 32 < p.n_store; };
      auto issue_load = [&](int tl, int c, int bsel) {
        const int m_tile = tl / a.n_tiles, n_tile = tl - m_tile * a.n_tiles;
        const int b = m_tile / a.tiles_per_img;
        const int t = m_tile - b * a.tiles_per_img;
        const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
        const int n0 = n_tile * BN + c * 32, x0 = tx * TW, y0 = ty * TH + quad * QR;
        mbar_arrive_expect_tx(&res_bar[ew][bsel], tx_bytes);
        if (p.res) tma_load_4d(wbase + s_res * 4096 + bsel * 2048, &tmR, &res_bar[ew][bsel], n0, x0, y0, b);
        if (p.mul) tma_load_4d(wbase + s_mul * 4096 + bsel * 2048, &tmO32, &res_bar[ew][bsel], n0, x0, y0, b);
        if (p.aux) tma_load_4d(wbase + s_aux * 4096 + bsel * 2048, &tmX, &res_bar[ew][bsel], n0, x0, y0, b);
      };
      {
        int ft = blockIdx.x;
        while (ft < num_tiles && !valid(ft, half)) ft += gridDim.x;
        if (lane == 0 && ft < num_tiles) issue_load(ft, half, 0);
      }
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
        const int b = m_tile / a.tiles_per_img;
        const int t = m_tile - b * a.tiles_per_img;
        const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
        for (int cb = half; valid(tile, cb); cb += 2) {
          const int n_blk = n_tile * BN + cb * 32;
          int ntile = tile, ncb2 = cb + 2;
          if (!valid(ntile, ncb2)) {
            ncb2 = half;
            ntile = tile + gridDim.x;
            while (ntile < num_tiles && !valid(ntile, ncb2)) ntile += gridDim.x;
          }
          if (lane == 0) {
            tma_store_wait_read<1>();      // the store that read out[buf] two items ago has drained
            if (ntile < num_tiles) issue_load(ntile, ncb2, buf ^ 1);   // operand buffers [buf^1] were consumed (generic reads) one item ago
          }
          uint32_t raw[32];
          tmem_ld16(taddr + cb * 32, *reinterpret_cast<uint32_t(*)[16]>(&raw[0]));
          tmem_ld16(taddr + cb * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&raw[16]));
          tc_wait_ld();
          mbar_wait(&res_bar[ew][buf], ph[buf]);
          ph[buf] ^= 1;
          __syncwarp();
          const int sw = (lane >> 1) & 3;
          const uint8_t* rrow = wbase + s_res * 4096 + buf * 2048 + lane * 64;
          const uint8_t* mrow = wbase + s_mul * 4096 + buf * 2048 + lane * 64;
          const uint8_t* xrow = wbase + s_aux * 4096 + buf * 2048 + lane * 64;
          uint8_t* orow = obase + buf * 2048 + lane * 64;
#pragma unroll
          for (int c = 0; c < 4; ++c) {          // 8 channels per 16-byte chunk
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[c * 8 + i]);
            if (p.bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n_blk + c * 8)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n_blk + c * 8) + 1);
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (p.act) apply_act_n(v, p.act);
            float cs[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) cs[i] = p.alpha;
            if (p.col_scale) {
#pragma unroll
              for (int i = 0; i < 8; ++i) cs[i] *= __ldg(p.col_scale + n_blk + c * 8 + i);
            }
            const int off = (c ^ sw) << 4;
            if (p.mul) {
              float m[8];
              const uint4 q = *reinterpret_cast<const uint4*>(mrow + off);
              m[0] = __uint_as_float(q.x << 16); m[1] = __uint_as_float(q.x & 0xffff0000u); m[2] = __uint_as_float(q.y << 16); m[3] = __uint_as_float(q.y & 0xffff0000u);
              m[4] = __uint_as_float(q.z << 16); m[5] = __uint_as_float(q.z & 0xffff0000u); m[6] = __uint_as_float(q.w << 16); m[7] = __uint_as_float(q.w & 0xffff0000u);
#pragma unroll
              for (int i = 0; i < 8; ++i) cs[i] *= m[i];
            }
            float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (p.res) {
              const uint4 q = *reinterpret_cast<const uint4*>(rrow + off);
              r[0] = __uint_as_float(q.x << 16); r[1] = __uint_as_float(q.x & 0xffff0000u); r[2] = __uint_as_float(q.y << 16); r[3] = __uint_as_float(q.y & 0xffff0000u);
              r[4] = __uint_as_float(q.z << 16); r[5] = __uint_as_float(q.z & 0xffff0000u); r[6] = __uint_as_float(q.w << 16); r[7] = __uint_as_float(q.w & 0xffff0000u);
            }
            if (p.aux) {
              float x[8];
              const uint4 q = *reinterpret_cast<const uint4*>(xrow + off);
              x[0] = __uint_as_float(q.x << 16); x[1] = __uint_as_float(q.x & 0xffff0000u); x[2] = __uint_as_float(q.y << 16); x[3] = __uint_as_float(q.y & 0xffff0000u);
              x[4] = __uint_as_float(q.z << 16); x[5] = __uint_as_float(q.z & 0xffff0000u); x[6] = __uint_as_float(q.w << 16); x[7] = __uint_as_float(q.w & 0xffff0000u);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float ch = p.aux_chan ? __ldg(p.aux_chan + (long long)b * p.aux_chan_ld + n_blk + c * 8 + i) : 1.f;
                r[i] = fmaf(p.aux_alpha * ch, x[i], r[i]);
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], cs[i], r[i]);
            if (p.post_act) apply_act_n(v, p.post_act);
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
              w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(orow + off) = make_uint4(w[0], w[1], w[2], w[3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmO, obase + buf * 2048, n_blk, tx * TW, ty * TH + quad * QR, b);
            tma_store_commit();
          }
          buf ^= 1;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (lane == 0) tma_store_wait_all();
    } else
    if constexpr (EPI == EPI_STORE_GATE) {
      // ---------- SimpleGate TMA-store epilogue: item = 64 accumulator columns (4 chunks of [8 x1 | 8 x2]) -> 32 bf16 outputs ----------
      static_assert(BN % 64 == 0, "EPI_STORE_GATE needs BN % 64 == 0");
      const FFConvGemm& p = a.p;
      const int ew = warp - 2;
      const int quad = warp & 3;
      const int half = ew >> 2;
      uint8_t* stg_base = smem + C::RING_BYTES + ew * C::STG_WARP_BYTES;
      int buf = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (lane == 0) tma_prefetch_desc(&tmO);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
        const int b = m_tile / a.tiles_per_img;
        const int t = m_tile - b * a.tiles_per_img;
        const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
        for (int it = half; it < BN / 64; it += 2) {
          const int n_blk = n_tile * BN + it * 64;
          if (n_blk >= p.n_store) break;
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          uint8_t* dst = stg_base + buf * 2048 + lane * 64;
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int c = 0; c < 4; ++c) {      // chunk c: accumulator columns [16c, 16c+16) -> output columns [8c, 8c+8)
            uint32_t raw[16];
            tmem_ld16(taddr + it * 64 + c * 16, raw);
            tc_wait_ld();
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = 2 * i;
              const float x1a = __uint_as_float(raw[j]) + __ldg(p.bias + n_blk + c * 16 + j), x2a = __uint_as_float(raw[8 + j]) + __ldg(p.bias + n_blk + c * 16 + 8 + j);
              const float x1b = __uint_as_float(raw[j + 1]) + __ldg(p.bias + n_blk + c * 16 + j + 1), x2b = __uint_as_float(raw[9 + j]) + __ldg(p.bias + n_blk + c * 16 + 9 + j);
              __nv_bfloat162 h = __floats2bfloat162_rn(x1a * x2a, x1b * x2b);
              w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(dst + ((c ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmO, stg_base + buf * 2048, n_blk >> 1, tx * TW, ty * TH + quad * QR, b);
            tma_store_commit();
          }
          buf ^= 1;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (lane == 0) tma_store_wait_all();
    } else if constexpr (EPI != EPI_GENERIC) {
      // ---------- TMA-store epilogue: bias (+GELU) -> bf16 -> swizzled smem -> cp.async.bulk.tensor store ----------
      static_assert(C::CB == 32, "TMA-store epilogue needs 32-column blocks");
      const FFConvGemm& p = a.p;
      const int ew = warp - 2;
      const int quad = warp & 3;
      const int half = ew >> 2;                 // position among the warps of this lane quadrant
      constexpr int CB_STEP = C::EPI_WARPS / 4;
      uint8_t* stg_base = smem + C::RING_BYTES + ew * C::STG_WARP_BYTES;   // 2 x 2 KB, 512-B aligned
      int buf = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (lane == 0) tma_prefetch_desc(&tmO);
      const bool plain_store = p.act == FF_ACT_NONE && p.col_sums == nullptr;
      (void)plain_store;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
        const int b = m_tile / a.tiles_per_img;
        const int t = m_tile - b * a.tiles_per_img;
        const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
        for (int cb = half; cb < BN / 32; cb += CB_STEP) {
          const int n_blk = n_tile * BN + cb * 32;
          if (n_blk >= p.n_store) break;
          uint32_t raw[32];
          tmem_ld16(taddr + cb * 32, *reinterpret_cast<uint32_t(*)[16]>(&raw[0]));
          tmem_ld16(taddr + cb * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&raw[16]));
          // the block's bias is requested while the TMEM load is in flight (tc_wait_ld is a compiler barrier for loads)
          float4 bq[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bq[j] = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n_blk) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          // the staging buffer about to be overwritten must have been read by the TMA store issued two blocks ago
          if (elect_one()) tma_store_wait_read<1>();      // elect.sync picks the same (lowest) lane every time: it owns the bulk groups
          __syncwarp();
          tc_wait_ld();
          uint8_t* dst = stg_base + buf * 2048 + lane * 64;
          const int sw = (lane >> 1) & 3;
          if (EPI == EPI_STORE && plain_store) {
            // lean path (bias only): 16 packed adds, 16 packs, four 16-byte stores
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t w[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 bv = bq[2 * c + (i >> 1)];
                const float2 b2 = (i & 1) ? make_float2(bv.z, bv.w) : make_float2(bv.x, bv.y);
                const float2 x2 = __fadd2_rn(make_float2(__uint_as_float(raw[c * 8 + 2 * i]), __uint_as_float(raw[c * 8 + 2 * i + 1])), b2);
                __nv_bfloat162 h = __floats2bfloat162_rn(x2.x, x2.y);
                w[i] = *reinterpret_cast<uint32_t*>(&h);
              }
              *reinterpret_cast<uint4*>(dst + ((c ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          } else {
          float cs_v[EPI == EPI_STORE ? 32 : 1];      // this lane's row of the block, kept for the column sums (EPI_STORE only)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t w[4];
            const float bb[8] = {bq[2 * c].x, bq[2 * c].y, bq[2 * c].z, bq[2 * c].w, bq[2 * c + 1].x, bq[2 * c + 1].y, bq[2 * c + 1].z, bq[2 * c + 1].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float x0 = __uint_as_float(raw[c * 8 + 2 * i]) + bb[2 * i], x1 = __uint_as_float(raw[c * 8 + 2 * i + 1]) + bb[2 * i + 1];
              if constexpr (EPI == EPI_STORE_GELU) { x0 = gelu_tanh_hw(x0); x1 = gelu_tanh_hw(x1); }
              else if (p.act) { x0 = apply_act(x0, p.act) * p.alpha; x1 = apply_act(x1, p.act) * p.alpha; }
              if constexpr (EPI == EPI_STORE) { cs_v[c * 8 + 2 * i] = x0; cs_v[c * 8 + 2 * i + 1] = x1; }
              __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
              w[i] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(dst + ((c ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
          if constexpr (EPI == EPI_STORE) {
            if (p.col_sums) {
              {   // rows of a partial edge tile that lie outside the image do not belong to the pool
                const int r_ = quad * 32 + lane;
                if (ty * TH + r_ / TW >= a.Ho || tx * TW + r_ % TW >= a.Wo) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) cs_v[j] = 0.f;
                }
              }
              // transpose-reduce: after the five halving steps lane c holds the sum of column c over the warp's 32 rows
#pragma unroll
              for (int off = 16; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int j = 0; j < off; ++j) {
                  const float send = up ? cs_v[j] : cs_v[j + off];
                  const float keep = up ? cs_v[j + off] : cs_v[j];
                  cs_v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
              }
              if (n_blk + lane < p.n_store) p.col_sums[((long long)m_tile * 4 + quad) * p.n_store + n_blk + lane] = cs_v[0];
            }
          }
          }  // !plain_store
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            tma_store_blk(&tmO, stg_base + buf * 2048, a, n_blk, tx * TW, ty * TH + quad * QR, b);
            tma_store_commit();
          }
          buf ^= 1;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (elect_one()) tma_store_wait_all();
    } else {
    // 8 warps: TMEM lane quadrant = warp % 4; the two warps of a quadrant alternate over CB-wide column blocks.
    // Per column block: TMEM -> registers (one accumulator row per thread) -> per-warp smem staging (transpose) ->
    // each lane owns 4 consecutive channels and walks the 32 rows, with all global operand loads issued up front.
    constexpr int CB = C::CB;                 // columns per block
    constexpr int LPR = CB / 4;               // lanes per row in the read-out phase
    constexpr int RPI = 32 / LPR;             // rows per warp instruction
    constexpr int ITERS = 32 / RPI;
    const FFConvGemm& p = a.p;
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int half = ew >> 2;
    float* stg = reinterpret_cast<float*>(smem + C::RING_BYTES) + ew * (32 * C::STG_PITCH);
    const int width = (p.pixel_shuffle == 2) ? (p.n_store >> 2) : (p.gate_pairs ? (p.n_store >> 1) : p.n_store);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
      const int b = m_tile / a.tiles_per_img;
      const int t = m_tile - b * a.tiles_per_img;
      const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int cb = half; cb < BN / CB; cb += 2) {
        const int n_blk = n_tile * BN + cb * CB;
        if (n_blk >= p.n_store) break;
        // ---- phase 1: TMEM -> smem staging, row `lane` of this quadrant
        int cw = CB;      // staged columns per row
        {
          float v[CB];
#pragma unroll
          for (int j = 0; j < CB / 16; ++j) {
            uint32_t raw[16];
            tmem_ld16(taddr + cb * CB + j * 16, raw);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[j * 16 + i] = __uint_as_float(raw[i]);
          }
          float* dst = stg + lane * C::STG_PITCH;
          if (p.gate_pairs) {
            // SimpleGate folded: (acc+bias)[i] * (acc+bias)[i+8] within each 16-column chunk -> CB/2 staged columns
            cw = CB / 2;
#pragma unroll
            for (int j = 0; j < CB / 16; ++j)
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float x1 = v[j * 16 + i] + __ldg(p.bias + n_blk + j * 16 + i);
                const float x2 = v[j * 16 + 8 + i] + __ldg(p.bias + n_blk + j * 16 + 8 + i);
                dst[j * 8 + i] = x1 * x2;
              }
          } else {
#pragma unroll
            for (int j = 0; j < CB / 4; ++j) *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
        __syncwarp();
        // ---- phase 2: coalesced read-out
        const int lpr = cw / 4;                       // lanes per row (LPR, or LPR/2 with gate_pairs)
        const int rpi = 32 / lpr;
        const int c4 = (lane % lpr) * 4;
        const int rsub = lane / lpr;
        const int ncol = p.gate_pairs ? ((n_blk >> 1) + c4) : (n_blk + c4);     // column in the (logical) output row
        long long opix0;
        int oc;
        {
          // column-dependent part of the output location (pixel part added per row below)
          if (p.pixel_shuffle == 2) {
            const int cq = p.n_store >> 2;
            oc = ncol % cq;
          } else {
            oc = ncol;
          }
          opix0 = 0;
        }
        const bool col_ok = (p.pixel_shuffle == 2) ? true : (ncol < width);
        if (a.vec_ok && (ncol + 4 <= ((p.pixel_shuffle == 2) ? p.n_store : width) || p.pixel_shuffle == 2)) {
          float bias[4] = {0.f, 0.f, 0.f, 0.f}, cscale[4], achan[4] = {1.f, 1.f, 1.f, 1.f};
          if (!p.gate_pairs && p.bias) {
#pragma unroll
            for (int i = 0; i < 4; ++i) bias[i] = __ldg(p.bias + ncol + i);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) cscale[i] = p.alpha * (p.col_scale ? __ldg(p.col_scale + ncol + i) : 1.f);
          if (p.aux && p.aux_chan) {
#pragma unroll
            for (int i = 0; i < 4; ++i) achan[i] = __ldg(p.aux_chan + (long long)b * p.aux_chan_ld + oc + i);
          }
          const int iters = 32 / rpi;
          // two passes of up to ITERS/…: keep the unroll bounded -- process rows in groups of 4 iterations
#pragma unroll 1
          for (int it0 = 0; it0 < iters; it0 += 4) {
            Vec4Operands ops[4];
            long long opx[4];
            float4 accv[4];
            bool inb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int row = (it0 + u) * rpi + rsub;           // row within the quadrant
              const int r = quad * 32 + row;
              const int oy = ty * TH + r / TW, ox = tx * TW + r % TW;
              inb[u] = oy < a.Hc && ox < a.Wc;                  // partial edge tiles / cropped outputs
              int dummy;
              long long px;
              if (p.pixel_shuffle == 2) out_location(a, b, oy, ox, ncol, px, dummy);
              else px = ((long long)(b * a.Hc + oy)) * a.Wc + ox;
              opx[u] = px;
              if (inb[u]) vec4_load(p, px, oc, ops[u]);
              accv[u] = *reinterpret_cast<const float4*>(stg + row * C::STG_PITCH + c4);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (!inb[u]) continue;
              if (p.gate_pairs) {
                // only the bf16 store applies
                __nv_bfloat162 lo = __floats2bfloat162_rn(accv[u].x, accv[u].y), hi = __floats2bfloat162_rn(accv[u].z, accv[u].w);
                *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out_bf16) + opx[u] * p.out_ld + oc) =
                    make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
              } else {
                vec4_finish(p, accv[u], ops[u], bias, cscale, achan, opx[u], oc);
              }
            }
          }
        } else if (col_ok) {
          // scalar fallback (narrow / unaligned outputs, e.g. 3-channel image writes)
          const int iters = 32 / rpi;
          for (int it = 0; it < iters; ++it) {
            const int row = it * rpi + rsub;
            const int r = quad * 32 + row;
            const int oy = ty * TH + r / TW, ox = tx * TW + r % TW;
            if (oy >= a.Hc || ox >= a.Wc) continue;
            const long long px = ((long long)(b * a.Hc + oy)) * a.Wc + ox;
            for (int i = 0; i < 4; ++i) {
              const int n = ncol + i;
              if (n >= width) break;
              float x = stg[row * C::STG_PITCH + c4 + i];
              if (p.bias) x += __ldg(p.bias + n);
              x = apply_act(x, p.act) * p.alpha;
              if (p.col_scale) x *= __ldg(p.col_scale + n);
              if (p.mul) x *= __bfloat162float(reinterpret_cast<const bf16*>(p.mul)[px * p.mul_ld + n]);
              if (p.aux) {
                const float m = __bfloat162float(reinterpret_cast<const bf16*>(p.aux)[px * p.aux_ld + n]);
                x += p.aux_alpha * m * (p.aux_chan ? __ldg(p.aux_chan + (long long)b * p.aux_chan_ld + n) : 1.f);
              }
              if (p.res) {
                x += p.res_is_f32 ? reinterpret_cast<const float*>(p.res)[px * p.res_ld + n]
                                  : __bfloat162float(reinterpret_cast<const bf16*>(p.res)[px * p.res_ld + n]);
              }
              x = apply_act(x, p.post_act);
              if (p.out_f32) p.out_f32[px * p.out_f32_ld + n] = x;
              if (p.out_bf16) reinterpret_cast<bf16*>(p.out_bf16)[px * p.out_ld + n] = __float2bfloat16_rn(x);
            }
          }
        }
        (void)opix0;
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    }  // EPI_GENERIC
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ----------------------------------------------------------------------------------------------
// SIMT reference main loop (testing only; same epilogue).  One thread per output pixel.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv_gemm_simt_kernel(const __grid_constant__ Args a) {
  const FFConvGemm& p = a.p;
  const int m_tile = blockIdx.x;
  const int r = threadIdx.x;
  const int b = m_tile / a.tiles_per_img;
  const int t = m_tile - b * a.tiles_per_img;
  const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
  const int oy = ty * TILE_H + (r >> 4), ox = tx * TILE_W + (r & 15);
  if (oy >= a.Ho || ox >= a.Wo) return;
  const bf16* x = reinterpret_cast<const bf16*>(p.x);
  const bf16* w = reinterpret_cast<const bf16*>(p.w);
  const int K = a.ntaps * p.cin;
  for (int n0 = blockIdx.y * 16; n0 < p.n_pad; n0 += gridDim.y * 16) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
    for (int tap = 0; tap < a.ntaps; ++tap) {
      int iy, ix;
      if (p.kind == FF_CONV_3X3) { iy = oy + tap / 3 - 1; ix = ox + tap % 3 - 1; }
      else if (p.kind == FF_CONV_2X2S2) { iy = 2 * oy + (tap >> 1); ix = 2 * ox + (tap & 1); }
      else { iy = oy; ix = ox; }
      if (iy < 0 || iy >= p.H || ix < 0 || ix >= p.W) continue;
      const bf16* xr = x + ((long long)(b * p.H + iy) * p.W + ix) * p.x_ld;
      for (int c = 0; c < p.cin; ++c) {
        const float xv = __bfloat162float(xr[c]);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += xv * __bfloat162float(w[(long long)(b * p.w_batch_rows + n0 + i) * K + tap * p.cin + c]);
      }
    }
    epilogue16(a, v, b, oy, ox, n0);
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

struct Maps { CUtensorMap A, B, O, R, O32, X, A2; };

template <int BN, int EPI, bool HALO = false>
int launch_tc(const Maps& m, const Args& a, cudaStream_t st) {
  using C = Cfg<BN, EPI, HALO>;
  static_assert(C::STAGES >= 2, "ff_conv_gemm: smem ring too shallow");
  static FFPerDeviceFlag configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, EPI, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) {
      ff_set_error("ff_conv_gemm: cudaFuncSetAttribute(%d) failed: %s", C::SMEM_BYTES, cudaGetErrorString(e));
      return FF_ERR_CUDA;
    }
    configured = true;
  }
  const int tiles = a.m_tiles * a.n_tiles;
  const int grid = tiles < ff_num_sms() ? tiles : ff_num_sms();
  const cudaError_t le = ff_launch_pdl(conv_gemm_tc_kernel<BN, EPI, HALO>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, m.A, m.B, m.O, m.R, m.O32, m.X, m.A2, a);
  if (le != cudaSuccess) { ff_set_error("ff_conv_gemm: launch failed: %s", cudaGetErrorString(le)); return FF_ERR_CUDA; }
  FF_CHECK_LAUNCH("ff_conv_gemm");
  return FF_OK;
}

constexpr int HALO_MAX_BN = 64;   // wider tiles are MMA-bound already and would not leave room for the slabs

template <int BN>
int launch_bn(int epi, const Maps& m, const Args& a, cudaStream_t st, bool halo) {
  if constexpr (BN <= HALO_MAX_BN) {
    if (halo) {
      if constexpr (BN >= 32) {
        if (epi == EPI_STORE) return launch_tc<BN, EPI_STORE, true>(m, a, st);
        if (epi == EPI_STORE_GELU) return launch_tc<BN, EPI_STORE_GELU, true>(m, a, st);
        if (epi == EPI_RES) return launch_tc<BN, EPI_RES, true>(m, a, st);
      }
      if constexpr (BN == 16) {
        if (epi == EPI_NARROW) return launch_tc<BN, EPI_NARROW, true>(m, a, st);
      }
      return launch_tc<BN, EPI_GENERIC, true>(m, a, st);
    }
  }
  if constexpr (BN == 16) {
    if (epi == EPI_NARROW) return launch_tc<BN, EPI_NARROW, false>(m, a, st);
  }
  if constexpr (BN == 64) {
    if (epi == EPI_OPS1) return launch_tc<BN, EPI_OPS1, false>(m, a, st);
    if (epi == EPI_OPS2) return launch_tc<BN, EPI_OPS2, false>(m, a, st);
    if (epi == EPI_OPS3) return launch_tc<BN, EPI_OPS3, false>(m, a, st);
  }
  if constexpr (BN >= 32) {
    if (epi == EPI_STORE) return launch_tc<BN, EPI_STORE>(m, a, st);
    if (epi == EPI_STORE_GELU) return launch_tc<BN, EPI_STORE_GELU>(m, a, st);
    if (epi == EPI_RES) return launch_tc<BN, EPI_RES>(m, a, st);
    if (epi == EPI_RES_AUX) return launch_tc<BN, EPI_RES_AUX>(m, a, st);
    if constexpr (BN % 64 == 0) {
      if (epi == EPI_STORE_GATE) return launch_tc<BN, EPI_STORE_GATE>(m, a, st);
    }
  }
  return launch_tc<BN, EPI_GENERIC>(m, a, st);
}

}  // namespace

extern long long g_ff_launches;

extern "C" int ff_conv_gemm(const FFConvGemm* pp, void* stream) {
  FF_CHECK_ARG(pp != nullptr, "ff_conv_gemm: null params");
  Args a;
  a.p = *pp;
  const FFConvGemm& p = a.p;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FF_CHECK_ARG(p.x && p.w, "ff_conv_gemm: null x/w");
  FF_CHECK_ARG(p.kind >= 0 && p.kind <= 2, "ff_conv_gemm: bad kind %d", p.kind);
  FF_CHECK_ARG(p.cin > 0 && p.cin % BLOCK_K == 0 && p.cin <= p.x_ld, "ff_conv_gemm: cin=%d must be a multiple of 64 and <= x_ld=%d", p.cin, p.x_ld);
  FF_CHECK_ARG(p.x_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0, "ff_conv_gemm: x must be 16B aligned with x_ld%%8==0");
  FF_CHECK_ARG(p.n_pad > 0 && p.n_pad % 16 == 0 && p.n_store > 0 && p.n_store <= p.n_pad, "ff_conv_gemm: bad n_pad=%d n_store=%d", p.n_pad, p.n_store);
  FF_CHECK_ARG(p.out_bf16 || p.out_f32, "ff_conv_gemm: no output buffer");
  a.Ho = (p.kind == FF_CONV_2X2S2) ? p.H / 2 : p.H;
  a.Wo = (p.kind == FF_CONV_2X2S2) ? p.W / 2 : p.W;
  // Any output size: edge tiles are partial -- TMA zero-fills loads and clips stores outside the tensor, the direct-store
  // epilogues mask by coordinates.
  FF_CHECK_ARG(a.Ho > 0 && a.Wo > 0 && (p.kind != FF_CONV_2X2S2 || (p.H % 2 == 0 && p.W % 2 == 0)), "ff_conv_gemm: bad spatial size %dx%d", p.H, p.W);
  a.Hc = p.out_crop_h > 0 ? p.out_crop_h : a.Ho;
  a.Wc = p.out_crop_w > 0 ? p.out_crop_w : a.Wo;
  FF_CHECK_ARG(a.Hc <= a.Ho && a.Wc <= a.Wo, "ff_conv_gemm: out_crop %dx%d exceeds the output %dx%d", a.Hc, a.Wc, a.Ho, a.Wo);
  static const bool halo_enabled = []() { const char* e = getenv("FFB200_CONV_HALO"); return !(e && e[0] == '0'); }();
  // bf16-operand TMA epilogue (EPI_OPS*): N tile 64, bf16 output, 1..3 bf16 operand tensors among res / mul / aux
  auto al16p = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  int n_ops = 0;
  if (p.n_pad % 64 == 0 && p.n_pad % 128 != 0 && p.n_pad % 192 != 0 && p.out_bf16 && !p.out_f32 && !p.pixel_shuffle && !p.gate_pairs && p.n_store % 8 == 0 &&
      (!p.bias || al16p(p.bias)) && p.out_ld % 8 == 0 && al16p(p.out_bf16) && (!p.res || (!p.res_is_f32 && p.res_ld % 8 == 0 && al16p(p.res))) &&
      (!p.mul || (p.mul_ld % 8 == 0 && al16p(p.mul))) && (!p.aux || (p.aux_ld % 8 == 0 && al16p(p.aux))))
    n_ops = (p.res ? 1 : 0) + (p.mul ? 1 : 0) + (p.aux ? 1 : 0);
  const bool halo = halo_enabled && !p.debug_simt && p.kind == FF_CONV_3X3 && p.n_pad % 128 != 0 && p.n_pad % 192 != 0 && p.n_pad <= 4 * HALO_MAX_BN &&
                    n_ops == 0;   // the operand rings and the halo slabs do not fit together (and measured slower at one ring)
  const int TW = halo ? HALO_TW : TILE_W, TH = halo ? HALO_TH : TILE_H;
  if (p.gate_pairs) {
    FF_CHECK_ARG(p.out_bf16 && !p.out_f32 && !p.act && !p.mul && !p.aux && !p.res && !p.pixel_shuffle && !p.col_scale && p.n_store % 16 == 0,
                 "ff_conv_gemm: gate_pairs supports bias + bf16 store only");
    FF_CHECK_ARG(p.out_ld >= p.n_store / 2 && p.out_ld % 8 == 0, "ff_conv_gemm: gate_pairs out_ld too small");
    FF_CHECK_ARG(p.bias != nullptr && p.n_store >= 32, "ff_conv_gemm: gate_pairs needs a bias vector and n_store >= 32");
  }
  if (p.w_batch_rows) FF_CHECK_ARG(p.w_batch_rows >= p.n_pad, "ff_conv_gemm: w_batch_rows < n_pad");
  if (p.pixel_shuffle && p.B > 1 && (a.Ho % TILE_H != 0 || a.Ho % HALO_TH != 0)) {
    // The 5-D PixelShuffle maps merge (batch, output row) into one dimension, so the out-of-range rows of a partial edge tile
    // would land in the next sample instead of being clipped: such shapes (deep UNet levels of un-aligned images) run one
    // sample per launch.
    FF_CHECK_ARG(!p.w_batch_rows && !p.aux_chan && !p.col_sums, "ff_conv_gemm: per-sample operands are not supported with an un-aligned PixelShuffle layer");
    for (int b = 0; b < p.B; ++b) {
      FFConvGemm q = p;
      q.B = 1;
      const long long in_px = (long long)b * p.H * p.W, out_px = (long long)b * 4 * a.Ho * a.Wo;
      q.x = reinterpret_cast<const bf16*>(p.x) + in_px * p.x_ld;
      if (p.res) q.res = p.res_is_f32 ? (const void*)(reinterpret_cast<const float*>(p.res) + out_px * p.res_ld) : (const void*)(reinterpret_cast<const bf16*>(p.res) + out_px * p.res_ld);
      if (p.out_bf16) q.out_bf16 = reinterpret_cast<bf16*>(p.out_bf16) + out_px * p.out_ld;
      if (p.out_f32) q.out_f32 = p.out_f32 + out_px * p.out_f32_ld;
      const int rc = ff_conv_gemm(&q, stream);
      if (rc != FF_OK) return rc;
    }
    return FF_OK;
  }
  if (p.pixel_shuffle) {
    FF_CHECK_ARG(p.pixel_shuffle == 2 && p.n_store % 64 == 0, "ff_conv_gemm: pixel_shuffle needs r=2 and n_store%%64==0");
    FF_CHECK_ARG(!p.mul && !p.aux, "ff_conv_gemm: pixel_shuffle supports bias/act/res only");
  }
  const int width_ok = p.pixel_shuffle ? (p.n_store >> 2) : p.gate_pairs ? (p.n_store >> 1) : p.n_store;
  const bool vec = width_ok >= 16;  // narrower outputs take the scalar store path: no alignment requirement
  a.vec_ok = vec ? 1 : 0;
  if (p.out_bf16) FF_CHECK_ARG(p.out_ld >= width_ok && (!vec || (p.out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out_bf16) & 15) == 0)), "ff_conv_gemm: bad out_ld=%d", p.out_ld);
  if (p.out_f32) FF_CHECK_ARG(p.out_f32_ld >= width_ok && (!vec || (p.out_f32_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(p.out_f32) & 15) == 0)), "ff_conv_gemm: bad out_f32_ld=%d", p.out_f32_ld);
  if (p.mul) FF_CHECK_ARG(p.mul_ld >= width_ok && (!vec || (p.mul_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(p.mul) & 15) == 0)), "ff_conv_gemm: bad mul_ld");
  if (p.aux) FF_CHECK_ARG(p.aux_ld >= width_ok && (!vec || (p.aux_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0)), "ff_conv_gemm: bad aux_ld");
  if (p.res) FF_CHECK_ARG(p.res_ld >= width_ok && (!vec || (p.res_ld % (p.res_is_f32 ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(p.res) & 15) == 0)), "ff_conv_gemm: bad res_ld");

  a.tiles_x = ff_cdiv(a.Wo, TW);
  a.tiles_per_img = a.tiles_x * ff_cdiv(a.Ho, TH);
  a.m_tiles = a.tiles_per_img * p.B;
  a.ntaps = (p.kind == FF_CONV_3X3) ? 9 : (p.kind == FF_CONV_2X2S2) ? 4 : 1;
  a.cchunks = p.cin / BLOCK_K;
  a.cchunks1 = a.cchunks;
  if (p.x2) {
    // K-concatenated 1x1 layer: out = [x | x2] . W^T with W [n_pad][cin + cin2]; x2 has the geometry of x.  Lets an epilogue add
    // (HAT: x = shortcut + proj(attn) + 0.01 * cab * se, hat_arch.py:306) run on the tensor pipe: x2 = cab, its weight block =
    // diag(0.01 * se_b) per sample (ff_build_concat_diag_weights), and the layer keeps the plain residual epilogue.
    FF_CHECK_ARG(p.kind == FF_CONV_1X1 && p.cin2 > 0 && p.cin2 % BLOCK_K == 0 && p.cin2 <= p.x2_ld && p.x2_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(p.x2) & 15) == 0 && !p.debug_simt,
                 "ff_conv_gemm: x2 needs a 1x1 layer, cin2 a multiple of 64 and 16-byte aligned rows");
    a.cchunks = (p.cin + p.cin2) / BLOCK_K;
  }

  if (p.debug_simt) {
    a.n_tiles = 1;
    dim3 grid(a.m_tiles, p.n_pad / 16 < 8 ? p.n_pad / 16 : 8);
    conv_gemm_simt_kernel<<<grid, 128, 0, st>>>(a);
    ++g_ff_launches;
    FF_CHECK_LAUNCH("ff_conv_gemm(simt)");
    return FF_OK;
  }

  int BN;
  if (p.n_pad % 256 == 0) BN = 256;
  else if (p.n_pad % 192 == 0) BN = 192;
  else if (p.n_pad % 128 == 0) BN = 128;
  else if (p.n_pad % 64 == 0) BN = 64;
  else if (p.n_pad % 32 == 0) BN = 32;
  else BN = 16;
  a.n_tiles = p.n_pad / BN;

  EncodeTiledFn enc = get_encode();
  if (!enc) {
    ff_set_error("ff_conv_gemm: cuTensorMapEncodeTiled entry point unavailable");
    return FF_ERR_DRIVER;
  }
  CUtensorMap tmA, tmB;
  {
    const int es = (p.kind == FF_CONV_2X2S2) ? 2 : 1;
    cuuint64_t dims[4] = {(cuuint64_t)p.cin, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    cuuint64_t strides[3] = {(cuuint64_t)p.x_ld * 2, (cuuint64_t)p.x_ld * 2 * p.W, (cuuint64_t)p.x_ld * 2 * p.W * p.H};
    cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(TILE_W * es), (cuuint32_t)(TILE_H * es), 1};
    if (halo) { box[1] = HALO_TW; box[2] = HALO_TH + 2; }
    cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      ff_set_error("ff_conv_gemm: cuTensorMapEncodeTiled(A) failed with %d (B=%d H=%d W=%d ld=%d cin=%d)", (int)r, p.B, p.H, p.W, p.x_ld, p.cin);
      return FF_ERR_DRIVER;
    }
  }
  {
    const cuuint64_t K = (cuuint64_t)a.ntaps * (p.cin + (p.x2 ? p.cin2 : 0));
    cuuint64_t dims[2] = {K, (cuuint64_t)(p.w_batch_rows ? (long long)p.w_batch_rows * p.B : p.n_pad)};
    cuuint64_t strides[1] = {K * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(p.w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      ff_set_error("ff_conv_gemm: cuTensorMapEncodeTiled(B) failed with %d (K=%llu n_pad=%d)", (int)r, (unsigned long long)K, p.n_pad);
      return FF_ERR_DRIVER;
    }
  }
  // epilogue selection: plain "bias (+GELU) -> bf16" layers and fp32-residual layers take the TMA epilogues
  int epi = EPI_GENERIC;
  Maps m;
  m.A = tmA; m.B = tmB; m.O = tmA; m.R = tmA; m.O32 = tmA; m.X = tmA; m.A2 = tmA;
  if (p.x2) {
    cuuint64_t dims[4] = {(cuuint64_t)p.cin2, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
    cuuint64_t strides[3] = {(cuuint64_t)p.x2_ld * 2, (cuuint64_t)p.x2_ld * 2 * p.W, (cuuint64_t)p.x2_ld * 2 * p.W * p.H};
    cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)TILE_W, (cuuint32_t)TILE_H, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&m.A2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.x2), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ff_set_error("ff_conv_gemm: cuTensorMapEncodeTiled(x2) failed with %d", (int)r); return FF_ERR_DRIVER; }
  }
  auto out_map = [&](CUtensorMap* tm, void* ptr, int ld, int esz, CUtensorMapDataType dt, CUtensorMapSwizzle sw) {
    if (p.pixel_shuffle) {
      const cuuint64_t e = (cuuint64_t)ld * esz;
      cuuint64_t dims[5] = {(cuuint64_t)(p.n_store >> 2), 2, (cuuint64_t)a.Wo, 2, (cuuint64_t)p.B * a.Ho};
      cuuint64_t strides[4] = {e, 2 * e, 2 * e * a.Wo, 4 * e * a.Wo};
      cuuint32_t box[5] = {32, 1, (cuuint32_t)TW, 1, (cuuint32_t)(32 / TW)};
      cuuint32_t estr[5] = {1, 1, 1, 1, 1};
      return enc(tm, dt, 5, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    cuuint64_t dims[4] = {(cuuint64_t)p.n_store, (cuuint64_t)a.Wo, (cuuint64_t)a.Ho, (cuuint64_t)p.B};
    cuuint64_t strides[3] = {(cuuint64_t)ld * esz, (cuuint64_t)ld * esz * a.Wo, (cuuint64_t)ld * esz * a.Wo * a.Ho};
    cuuint32_t box[4] = {32, (cuuint32_t)TW, (cuuint32_t)(32 / TW), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(tm, dt, 4, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  const bool al16 = (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) && (!p.col_scale || (reinterpret_cast<uintptr_t>(p.col_scale) & 15) == 0);
  const bool bf16_ok = !p.out_bf16 || (p.out_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out_bf16) & 15) == 0);
  const bool ps_ok = !p.pixel_shuffle || (p.n_store % 128 == 0);     // 32-column blocks must not straddle a sub-pixel group
  const bool base_ok = !p.mul && !p.post_act && !p.pixel_shuffle && p.n_store % 8 == 0 && al16 && bf16_ok && BN >= 32;
  const bool plain_ps = !p.mul && !p.post_act && ps_ok && p.n_store % 8 == 0 && al16 && bf16_ok && BN >= 32 && !p.aux && !p.gate_pairs;
  const bool plain = base_ok && !p.aux && !p.gate_pairs;
  auto map_n = [&](CUtensorMap* tm, void* ptr, int ld, int esz, CUtensorMapDataType dt, CUtensorMapSwizzle sw, int ncols) {
    cuuint64_t dims[4] = {(cuuint64_t)ncols, (cuuint64_t)a.Wo, (cuuint64_t)a.Ho, (cuuint64_t)p.B};
    cuuint64_t strides[3] = {(cuuint64_t)ld * esz, (cuuint64_t)ld * esz * a.Wo, (cuuint64_t)ld * esz * a.Wo * a.Ho};
    cuuint32_t box[4] = {32, (cuuint32_t)TW, (cuuint32_t)(32 / TW), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(tm, dt, 4, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  };
  const bool res_ok = p.res && p.res_is_f32 && p.out_f32 && p.act == FF_ACT_NONE && p.res_ld % 4 == 0 && p.out_f32_ld % 4 == 0 &&
                      (reinterpret_cast<uintptr_t>(p.res) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out_f32) & 15) == 0;
  if (n_ops > 0) {
    bool ok = out_map(&m.O, p.out_bf16, p.out_ld, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B);
    if (ok && p.res) ok = out_map(&m.R, const_cast<void*>(p.res), p.res_ld, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B);
    if (ok && p.mul) ok = out_map(&m.O32, const_cast<void*>(p.mul), p.mul_ld, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B);
    if (ok && p.aux) ok = out_map(&m.X, const_cast<void*>(p.aux), p.aux_ld, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B);
    FF_CHECK_ARG(ok, "ff_conv_gemm: cuTensorMapEncodeTiled failed for an epilogue operand");
    epi = EPI_OPS1 + n_ops - 1;
  } else if (plain_ps && p.out_bf16 && !p.out_f32 && !p.res && !p.col_scale && (p.alpha == 1.0f || (p.act != FF_ACT_NONE && p.act != FF_ACT_GELU))) {
    if (out_map(&m.O, p.out_bf16, p.out_ld, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B))
      epi = (p.act == FF_ACT_GELU) ? EPI_STORE_GELU : EPI_STORE;
  } else if (base_ok && p.gate_pairs && BN % 64 == 0 && p.n_store % 64 == 0 && p.bias && !p.aux && !p.res && !p.out_f32 && !p.col_scale && p.alpha == 1.0f && p.act == FF_ACT_NONE) {
    if (map_n(&m.O, p.out_bf16, p.out_ld, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B, p.n_store / 2)) epi = EPI_STORE_GATE;
  } else if (plain_ps && res_ok) {
    bool ok = out_map(&m.R, const_cast<void*>(p.res), p.res_ld, 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B) &&
              out_map(&m.O32, p.out_f32, p.out_f32_ld, 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B);
    if (ok && p.out_bf16) ok = out_map(&m.O, p.out_bf16, p.out_ld, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B);
    if (ok) epi = EPI_RES;
  } else if (base_ok && res_ok && p.aux && !p.gate_pairs && !p.out_bf16 && p.aux_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0 &&
             (!p.aux_chan || (p.aux_chan_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(p.aux_chan) & 15) == 0))) {
    bool ok = out_map(&m.R, const_cast<void*>(p.res), p.res_ld, 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B) &&
              out_map(&m.O32, p.out_f32, p.out_f32_ld, 4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B) &&
              out_map(&m.O, const_cast<void*>(p.aux), p.aux_ld, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B);
    if (ok) epi = EPI_RES_AUX;
  }
  if (BN == 16 && p.n_store <= 4 && !p.mul && !p.aux && !p.pixel_shuffle && !p.gate_pairs) epi = EPI_NARROW;
  if (a.Hc != a.Ho || a.Wc != a.Wo)
    FF_CHECK_ARG(epi == EPI_NARROW || (epi == EPI_GENERIC && !p.res && !p.mul && !p.aux && !p.pixel_shuffle),
                 "ff_conv_gemm: out_crop needs a direct-store epilogue (narrow outputs, or no per-pixel operands)");
  if (p.ln_out) {
    FF_CHECK_ARG((epi == EPI_RES || epi == EPI_RES_AUX) && !halo && a.n_tiles == 1 && !p.pixel_shuffle && p.n_store == p.n_pad,
                 "ff_conv_gemm: ln_out needs the fp32-residual epilogue with one n tile spanning the row (n_pad=%d n_store=%d epi=%d)", p.n_pad, p.n_store, epi);
    FF_CHECK_ARG(p.ln_gamma && p.ln_beta && al16p(p.ln_gamma) && al16p(p.ln_beta) && p.ln_cols > 0 && p.ln_cols <= p.n_store && p.ln_eps > 0.f,
                 "ff_conv_gemm: ln_gamma / ln_beta must be 16-byte aligned [n_store] vectors, 0 < ln_cols <= n_store");
    FF_CHECK_ARG(p.ln_out_ld >= p.n_store && p.ln_out_ld % 8 == 0 && al16p(p.ln_out), "ff_conv_gemm: bad ln_out_ld=%d", p.ln_out_ld);
    FF_CHECK_ARG(out_map(&m.X, p.ln_out, p.ln_out_ld, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B), "ff_conv_gemm: tensor map of ln_out failed");
  }
  FF_CHECK_ARG(!p.col_sums || (epi == EPI_STORE && !p.pixel_shuffle), "ff_conv_gemm: col_sums needs the plain bf16-store epilogue (bias, optional non-GELU act)");
  ++g_ff_launches;
  switch (BN) {
    case 256: return launch_bn<256>(epi, m, a, st, false);
    case 192: return launch_bn<192>(epi, m, a, st, false);
    case 128: return launch_bn<128>(epi, m, a, st, false);
    case 64: return launch_bn<64>(epi, m, a, st, halo);
    case 32: return launch_bn<32>(epi, m, a, st, halo);
    default: return launch_bn<16>(epi, m, a, st, halo);
  }
}
