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
// Image I/O ends of the plugin (reference io.py:64-76): uint8 HWC image -> fp32 NCHW tiles in [0,1] (np.float32(arr) / 255.0, an
// IEEE fp32 division, fused with the tile extraction of _tiled_forward :99-103), and fp32 NCHW -> uint8 HWC
// (round_half_even(clamp(x, 0, 1) * 255), the quantisation of _save_image) for results that do not pass through ff_stitch.
// ----------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) u8_to_tiles_kernel(const unsigned char* __restrict__ img, int H, int W, const int* __restrict__ ys,
                                                         const int* __restrict__ xs, int nx, int th, int tw, float* __restrict__ tiles) {
  const int t = blockIdx.y;
  const int iy = t / nx, ix = t - iy * nx;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= th * tw) return;
  const int y = i / tw, x = i - y * tw;
  const unsigned char* p = img + ((long long)(ys[iy] + y) * W + xs[ix] + x) * 3;
  float* o = tiles + (long long)t * 3 * th * tw + i;
  o[0] = __fdiv_rn((float)p[0], 255.0f);
  o[(long long)th * tw] = __fdiv_rn((float)p[1], 255.0f);
  o[2LL * th * tw] = __fdiv_rn((float)p[2], 255.0f);
}
__global__ void __launch_bounds__(256) quantize_u8_kernel(const float* __restrict__ x, long long hw, unsigned char* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= hw) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) out[i * 3 + c] = (unsigned char)__float2int_rn(__fmul_rn(fminf(fmaxf(x[c * hw + i], 0.f), 1.f), 255.f));
}
}  // namespace

extern "C" int ff_u8_to_tiles(const unsigned char* img, int H, int W, const int* ys, const int* xs, int ny, int nx, int th, int tw, float* tiles,
                              void* stream) {
  FF_CHECK_ARG(img && ys && xs && tiles && H > 0 && W > 0 && ny > 0 && nx > 0 && th > 0 && tw > 0 && th <= H && tw <= W, "ff_u8_to_tiles: bad args");
  FF_CHECK_ARG(ny * nx <= 65535, "ff_u8_to_tiles: too many tiles");
  dim3 grid(ff_cdiv((long long)th * tw, 256), ny * nx);
  u8_to_tiles_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(img, H, W, ys, xs, nx, th, tw, tiles);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_u8_to_tiles");
  return FF_OK;
}

extern "C" int ff_quantize_u8(const float* x, int H, int W, unsigned char* out, void* stream) {
  FF_CHECK_ARG(x && out && H > 0 && W > 0, "ff_quantize_u8: bad args");
  quantize_u8_kernel<<<ff_cdiv((long long)H * W, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, (long long)H * W, out);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_quantize_u8");
  return FF_OK;
}

// ----------------------------------------------------------------------------------------------
// PSNR on the BT.601 luma channel with a border crop (reference src/utils/metrics.py:30-52, 76-126: rgb_to_y then
// crop_border=4, MSE on [0,1] data).  Two-phase deterministic reduction: per-block partial sums of squared Y differences.
// ----------------------------------------------------------------------------------------------
namespace {
// metrics.py:30-52 after the clamp(0, 1) of calculate_psnr / calculate_ssim (:104-105, :208-209)
__device__ __forceinline__ float luma601(const float* __restrict__ p, long long o, long long hw) {
  return (65.481f * __saturatef(p[o]) + 128.553f * __saturatef(p[hw + o]) + 24.966f * __saturatef(p[2 * hw + o]) + 16.0f) / 255.0f;
}
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
    const float ya = luma601(pa, o, hw), yb = luma601(pb, o, hw);
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

// ----------------------------------------------------------------------------------------------
// SSIM on the BT.601 luma channel with a border crop (reference src/utils/metrics.py:189-246 -> :129-186, the PyTorch branch
// calculate_ssim takes without scikit-image): 11x11 Gaussian window (sigma 1.5, outer product of the normalised 1-D
// window), zero padding of the cropped image, C1 = 0.01^2, C2 = 0.03^2, mean of the full map.  One CTA per 32x32 tile of
// the map: both luma tiles with a 5-pixel halo in smem, separable filter of the five moments (rows, then columns), two-phase
// deterministic fp64 reduction.
// ----------------------------------------------------------------------------------------------
namespace {
constexpr int SS_T = 32, SS_R = 5, SS_H = SS_T + 2 * SS_R;   // tile, window radius, tile + halo
__constant__ float c_ssim_win[11];

__global__ void __launch_bounds__(256) ssim_y_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W, int crop,
                                                    int tiles_x, double* __restrict__ partial) {
  __shared__ float ya[SS_H][SS_H + 1], yb[SS_H][SS_H + 1];
  __shared__ float hq[5][SS_H][SS_T];
  __shared__ double red[256];
  const int Hc = H - 2 * crop, Wc = W - 2 * crop;
  const long long hw = (long long)H * W;
  const float* pa = a + (long long)blockIdx.y * 3 * hw;
  const float* pb = b + (long long)blockIdx.y * 3 * hw;
  const int ty0 = (blockIdx.x / tiles_x) * SS_T, tx0 = (blockIdx.x % tiles_x) * SS_T;
  for (int i = threadIdx.x; i < SS_H * SS_H; i += 256) {
    const int r = i / SS_H, c = i - r * SS_H;
    const int y = ty0 + r - SS_R, x = tx0 + c - SS_R;      // coordinates in the cropped image
    float va = 0.f, vb = 0.f;                               // conv2d zero padding
    if (y >= 0 && y < Hc && x >= 0 && x < Wc) {
      const long long o = (long long)(y + crop) * W + (x + crop);
      va = luma601(pa, o, hw);
      vb = luma601(pb, o, hw);
    }
    ya[r][c] = va;
    yb[r][c] = vb;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < SS_H * SS_T; i += 256) {
    const int r = i / SS_T, c = i - r * SS_T;
    float s1 = 0.f, s2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float w = c_ssim_win[k], u = ya[r][c + k], v = yb[r][c + k];
      s1 += w * u; s2 += w * v; s11 += w * (u * u); s22 += w * (v * v); s12 += w * (u * v);
    }
    hq[0][r][c] = s1; hq[1][r][c] = s2; hq[2][r][c] = s11; hq[3][r][c] = s22; hq[4][r][c] = s12;
  }
  __syncthreads();
  double acc = 0.0;
  for (int i = threadIdx.x; i < SS_T * SS_T; i += 256) {
    const int r = i / SS_T, c = i - r * SS_T;
    if (ty0 + r >= Hc || tx0 + c >= Wc) continue;
    float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float w = c_ssim_win[k];
      m1 += w * hq[0][r + k][c]; m2 += w * hq[1][r + k][c]; e11 += w * hq[2][r + k][c]; e22 += w * hq[3][r + k][c];
      e12 += w * hq[4][r + k][c];
    }
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
    const float v1 = e11 - m11, v2 = e22 - m22, cov = e12 - m12;
    acc += (double)(((2.f * m12 + C1) * (2.f * cov + C2)) / ((m11 + m22 + C1) * (v1 + v2 + C2)));
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[(long long)blockIdx.y * gridDim.x + blockIdx.x] = red[0];
}
__global__ void ssim_final_kernel(const double* __restrict__ partial, int nblk, double n, float* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nblk; i += 256) s += partial[(long long)blockIdx.x * nblk + i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = (float)(red[0] / n);
}
}  // namespace

extern "C" size_t ff_ssim_y_scratch_bytes(int B, int H, int W, int crop) {
  if (B <= 0 || H <= 2 * crop || W <= 2 * crop || crop < 0) return 0;
  const size_t tiles = (size_t)((H - 2 * crop + SS_T - 1) / SS_T) * (size_t)((W - 2 * crop + SS_T - 1) / SS_T);
  return (size_t)B * tiles * sizeof(double);
}

extern "C" int ff_ssim_y(const float* a, const float* b, int B, int H, int W, int crop, float* out, double* scratch, size_t scratch_bytes,
                         void* stream) {
  FF_CHECK_ARG(a && b && out && scratch && B > 0 && H > 2 * crop && W > 2 * crop && crop >= 0, "ff_ssim_y: bad args");
  FF_CHECK_ARG(B <= 65535, "ff_ssim_y: batch too large");
  FF_CHECK_ARG(scratch_bytes >= ff_ssim_y_scratch_bytes(B, H, W, crop), "ff_ssim_y: scratch too small");
  static FFPerDeviceFlag have_window_dev;
  bool& have_window = have_window_dev.get();
  if (!have_window) {
    // metrics.py:150-154: exp(-(x-5)^2 / (2 sigma^2)) normalised to sum 1 (float32 tensor arithmetic in the reference)
    float g[11], sum = 0.f;
    for (int x = 0; x < 11; ++x) { g[x] = (float)exp(-(double)((x - 5) * (x - 5)) / (2.0 * 1.5 * 1.5)); sum += g[x]; }
    for (int x = 0; x < 11; ++x) g[x] /= sum;
    cudaError_t e = cudaMemcpyToSymbol(c_ssim_win, g, sizeof(g));
    if (e != cudaSuccess) {
      ff_set_error("ff_ssim_y: window upload: %s", cudaGetErrorString(e));
      return FF_ERR_CUDA;
    }
    have_window = true;
  }
  const int tiles_x = (W - 2 * crop + SS_T - 1) / SS_T, tiles_y = (H - 2 * crop + SS_T - 1) / SS_T;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ssim_y_kernel<<<dim3(tiles_x * tiles_y, B), 256, 0, st>>>(a, b, H, W, crop, tiles_x, scratch);
  ssim_final_kernel<<<B, 256, 0, st>>>(scratch, tiles_x * tiles_y, (double)(H - 2 * crop) * (W - 2 * crop), out);
  g_ff_launches += 2;
  FF_CHECK_LAUNCH("ff_ssim_y");
  return FF_OK;
}
