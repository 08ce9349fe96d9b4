// Overlapped-tile stitching (reference models/team29_FreqFusion/io.py:82-121, _tiled_forward):
//   canvas += sr_tile * (wy (x) wx);  weight_map += wy (x) wx;  out = canvas / clamp(weight_map, 1e-8)
// evaluated as a gather per output pixel that visits the tiles in the reference's order (y-major, x-minor) with
// un-fused fp32 multiply / add, so the result is bit-identical to the reference's sequential accumulation.
// Optionally also emits the uint8 image of io._save_image (:71-76): round_half_even(clamp(x,0,1) * 255).
#include "ff_common.cuh"
#include "../../include/ffb200.h"

extern long long g_ff_launches;

namespace {

struct StitchArgs {
  const float* tiles;     // [T][3][ts][ts] fp32 NCHW SR tiles, T = ny*nx in (y-major, x-minor) order
  const int* ty; const int* tx;   // HR top-left of each tile row / column: [ny], [nx]
  const float* wy; const float* wx; // per tile-row / tile-column 1-D blend weights: [ny][ts], [nx][ts]
  int ny, nx, ts;
  int H, W;               // HR canvas size
  float* out;             // [3][H][W] fp32 or null
  unsigned char* out_u8;  // [H][W][3] uint8 (HWC, as PIL wants it) or null
};

__global__ void __launch_bounds__(256) stitch_kernel(const __grid_constant__ StitchArgs a) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)a.H * a.W) return;
  const int X = (int)(idx % a.W), Y = (int)(idx / a.W);
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, wsum = 0.f;
  const long long plane = (long long)a.ts * a.ts;
  for (int iy = 0; iy < a.ny; ++iy) {
    const int ly = Y - a.ty[iy];
    if (ly < 0 || ly >= a.ts) continue;
    const float wyv = a.wy[iy * a.ts + ly];
    for (int ix = 0; ix < a.nx; ++ix) {
      const int lx = X - a.tx[ix];
      if (lx < 0 || lx >= a.ts) continue;
      const float wgt = __fmul_rn(wyv, a.wx[ix * a.ts + lx]);
      const float* t = a.tiles + ((long long)(iy * a.nx + ix) * 3) * plane + (long long)ly * a.ts + lx;
      acc0 = __fadd_rn(acc0, __fmul_rn(t[0], wgt));
      acc1 = __fadd_rn(acc1, __fmul_rn(t[plane], wgt));
      acc2 = __fadd_rn(acc2, __fmul_rn(t[2 * plane], wgt));
      wsum = __fadd_rn(wsum, wgt);
    }
  }
  const float den = fmaxf(wsum, 1e-8f);
  const float v0 = __fdiv_rn(acc0, den), v1 = __fdiv_rn(acc1, den), v2 = __fdiv_rn(acc2, den);
  if (a.out) {
    const long long hw = (long long)a.H * a.W;
    a.out[idx] = v0; a.out[hw + idx] = v1; a.out[2 * hw + idx] = v2;
  }
  if (a.out_u8) {
    a.out_u8[idx * 3] = (unsigned char)__float2int_rn(__fmul_rn(fminf(fmaxf(v0, 0.f), 1.f), 255.f));
    a.out_u8[idx * 3 + 1] = (unsigned char)__float2int_rn(__fmul_rn(fminf(fmaxf(v1, 0.f), 1.f), 255.f));
    a.out_u8[idx * 3 + 2] = (unsigned char)__float2int_rn(__fmul_rn(fminf(fmaxf(v2, 0.f), 1.f), 255.f));
  }
}

}  // namespace

extern "C" int ff_stitch(const float* tiles, const int* ty, const int* tx, const float* wy, const float* wx, int ny, int nx, int ts,
                         int H, int W, float* out, unsigned char* out_u8, void* stream) {
  FF_CHECK_ARG(tiles && ty && tx && wy && wx && (out || out_u8), "ff_stitch: null buffer");
  FF_CHECK_ARG(ny > 0 && nx > 0 && ts > 0 && H > 0 && W > 0, "ff_stitch: bad sizes");
  StitchArgs a{tiles, ty, tx, wy, wx, ny, nx, ts, H, W, out, out_u8};
  stitch_kernel<<<ff_cdiv((long long)H * W, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_stitch");
  return FF_OK;
}

// ----------------------------------------------------------------------------------------------
// PSNR on the BT.601 luma channel with a border crop (reference src/utils/metrics.py:30-52, 76-126: rgb_to_y then
// crop_border=4, MSE on [0,1] data).  Two-phase deterministic reduction: per-block partial sums of squared Y differences.
// ----------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) sqdiff_y_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W, int crop,
                                                      double* __restrict__ partial) {
  __shared__ double red[256];
  const int Hc = H - 2 * crop, Wc = W - 2 * crop;
  const long long n = (long long)Hc * Wc;
  const long long hw = (long long)H * W;
  const float* pa = a + (long long)blockIdx.y * 3 * hw;
  const float* pb = b + (long long)blockIdx.y * 3 * hw;
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int y = (int)(i / Wc) + crop, x = (int)(i % Wc) + crop;
    const long long o = (long long)y * W + x;
    const float ya = (65.481f * pa[o] + 128.553f * pa[hw + o] + 24.966f * pa[2 * hw + o] + 16.0f) / 255.0f;
    const float yb = (65.481f * pb[o] + 128.553f * pb[hw + o] + 24.966f * pb[2 * hw + o] + 16.0f) / 255.0f;
    const double d = (double)ya - (double)yb;
    s += d * d;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(long long)blockIdx.y * gridDim.x + blockIdx.x] = red[0];
}
__global__ void psnr_final_kernel(const double* __restrict__ partial, int nblk, double n, float* __restrict__ out) {
  const int b = blockIdx.x;
  double s = 0.0;
  for (int i = 0; i < nblk; ++i) s += partial[(long long)b * nblk + i];
  const double mse = s / n;
  out[b] = mse <= 0.0 ? 100.0f : (float)(10.0 * log10(1.0 / mse));
}
}  // namespace

extern "C" int ff_psnr_y(const float* a, const float* b, int B, int H, int W, int crop, float* out, double* scratch, size_t scratch_bytes,
                         void* stream) {
  FF_CHECK_ARG(a && b && out && scratch && H > 2 * crop && W > 2 * crop && crop >= 0, "ff_psnr_y: bad args");
  const int nblk = 64;
  FF_CHECK_ARG(scratch_bytes >= (size_t)B * nblk * sizeof(double), "ff_psnr_y: scratch too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  sqdiff_y_kernel<<<dim3(nblk, B), 256, 0, st>>>(a, b, H, W, crop, scratch);
  psnr_final_kernel<<<B, 1, 0, st>>>(scratch, nblk, (double)(H - 2 * crop) * (W - 2 * crop), out);
  g_ff_launches += 2;
  FF_CHECK_LAUNCH("ff_psnr_y");
  return FF_OK;
}
