// ff_eval_psnr_ssim_u8: the PSNR / SSIM pair of the reference's evaluation harness (eval.py:157 -> utils/utils_image.py:287-312
// `cal_psnr_ssim`) on two uint8 RGB images that already sit in device memory:
//   crop `border` pixels; Y = OpenCV's 8-bit RGB2YCrCb luma (fixed point: (4899 R + 9617 G + 1868 B + 2^13) >> 14);
//   PSNR = 10 log10(255^2 / mean((Ya - Yb)^2)) in double precision (inf for identical images);
//   SSIM = scikit-image's structural_similarity defaults: 7x7 uniform window, sample covariance (49 / 48), K1 = 0.01, K2 = 0.03,
//          data_range 255, double precision, mean over the windows that lie inside the cropped image.
// Y is an integer, so every 7x7 window sum (of y, y^2, xy) is formed exactly in int32 and only the final per-window formula and the
// mean run in fp64; the squared luma differences of the PSNR are summed exactly in 64-bit integers.  Both reductions are two-phase
// and deterministic.  One CTA per 32x32 tile of the SSIM map (both luma tiles with a 3-pixel halo in shared memory, separable box sums).
#include "ff_common.cuh"
#include "../../include/ffb200.h"

extern long long g_ff_launches;

namespace {

constexpr int ET = 32, ER = 3, EH = ET + 2 * ER;      // map tile, window radius, luma tile with halo
constexpr int SQ_BLOCKS = 64;

__device__ __forceinline__ int luma_cv(const unsigned char* __restrict__ p) { return (4899 * (int)p[0] + 9617 * (int)p[1] + 1868 * (int)p[2] + 8192) >> 14; }

__global__ void __launch_bounds__(256) eval_ssim_kernel(const unsigned char* __restrict__ a, const unsigned char* __restrict__ b, int H, int W, int border,
                                                       int tiles_x, double* __restrict__ partial) {
  __shared__ int ya[EH][EH + 1], yb[EH][EH + 1];
  __shared__ int hq[5][EH][ET];
  __shared__ double red[256];
  const int Hc = H - 2 * border, Wc = W - 2 * border;
  const int Hm = Hc - 2 * ER, Wm = Wc - 2 * ER;      // SSIM map: one value per 7x7 window inside the cropped image
  const int ty0 = (blockIdx.x / tiles_x) * ET, tx0 = (blockIdx.x % tiles_x) * ET;
  for (int i = threadIdx.x; i < EH * EH; i += 256) {
    const int r = i / EH, c = i - r * EH;
    const int y = ty0 + r, x = tx0 + c;      // cropped-image coordinates of the window's top-left corner + (r, c)
    int va = 0, vb = 0;
    if (y < Hc && x < Wc) {
      const long long o = ((long long)(y + border) * W + (x + border)) * 3;
      va = luma_cv(a + o);
      vb = luma_cv(b + o);
    }
    ya[r][c] = va;
    yb[r][c] = vb;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < EH * ET; i += 256) {
    const int r = i / ET, c = i - r * ET;
    int s1 = 0, s2 = 0, s11 = 0, s22 = 0, s12 = 0;
#pragma unroll
    for (int k = 0; k < 2 * ER + 1; ++k) {
      const int u = ya[r][c + k], v = yb[r][c + k];
      s1 += u; s2 += v; s11 += u * u; s22 += v * v; s12 += u * v;
    }
    hq[0][r][c] = s1; hq[1][r][c] = s2; hq[2][r][c] = s11; hq[3][r][c] = s22; hq[4][r][c] = s12;
  }
  __syncthreads();
  double acc = 0.0;
  for (int i = threadIdx.x; i < ET * ET; i += 256) {
    const int r = i / ET, c = i - r * ET;
    if (ty0 + r >= Hm || tx0 + c >= Wm) continue;
    int s1 = 0, s2 = 0, s11 = 0, s22 = 0, s12 = 0;
#pragma unroll
    for (int k = 0; k < 2 * ER + 1; ++k) {
      s1 += hq[0][r + k][c]; s2 += hq[1][r + k][c]; s11 += hq[2][r + k][c]; s22 += hq[3][r + k][c]; s12 += hq[4][r + k][c];
    }
    const double np_ = 49.0, cov_norm = 49.0 / 48.0;
    const double ux = (double)s1 / np_, uy = (double)s2 / np_;
    const double vx = cov_norm * ((double)s11 / np_ - ux * ux), vy = cov_norm * ((double)s22 / np_ - uy * uy);
    const double vxy = cov_norm * ((double)s12 / np_ - ux * uy);
    const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
    acc += ((2.0 * ux * uy + C1) * (2.0 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) eval_sqdiff_kernel(const unsigned char* __restrict__ a, const unsigned char* __restrict__ b, int H, int W, int border,
                                                         unsigned long long* __restrict__ partial) {
  __shared__ unsigned long long red[256];
  const int Hc = H - 2 * border, Wc = W - 2 * border;
  const long long n = (long long)Hc * Wc;
  unsigned long long s = 0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int y = (int)(i / Wc) + border, x = (int)(i % Wc) + border;
    const long long o = ((long long)y * W + x) * 3;
    const int d = luma_cv(a + o) - luma_cv(b + o);
    s += (unsigned long long)(d * d);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) eval_final_kernel(const double* __restrict__ ssim_partial, int ntiles, const unsigned long long* __restrict__ sq_partial,
                                                        double n_px, double n_map, double* __restrict__ out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < ntiles; i += 256) s += ssim_partial[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    unsigned long long sq = 0;
    for (int i = 0; i < SQ_BLOCKS; ++i) sq += sq_partial[i];
    const double mse = (double)sq / n_px;
    out[0] = sq == 0 ? (double)INFINITY : 10.0 * log10(255.0 * 255.0 / mse);
    out[1] = red[0] / n_map;
  }
}

}  // namespace

extern "C" size_t ff_eval_scratch_bytes(int H, int W, int border) {
  const int Hm = H - 2 * border - 2 * ER, Wm = W - 2 * border - 2 * ER;
  if (border < 0 || Hm <= 0 || Wm <= 0) return 0;
  const size_t tiles = (size_t)((Hm + ET - 1) / ET) * (size_t)((Wm + ET - 1) / ET);
  return (tiles + SQ_BLOCKS) * sizeof(double);
}

extern "C" int ff_eval_psnr_ssim_u8(const unsigned char* a, const unsigned char* b, int H, int W, int border, double* out, void* scratch,
                                    size_t scratch_bytes, void* stream) {
  FF_CHECK_ARG(a && b && out && scratch, "ff_eval_psnr_ssim_u8: null buffer");
  FF_CHECK_ARG(border >= 0 && H - 2 * border >= 2 * ER + 1 && W - 2 * border >= 2 * ER + 1,
               "ff_eval_psnr_ssim_u8: the cropped image must hold a 7x7 window (H=%d W=%d border=%d)", H, W, border);
  FF_CHECK_ARG(scratch_bytes >= ff_eval_scratch_bytes(H, W, border) && (reinterpret_cast<uintptr_t>(scratch) & 7) == 0, "ff_eval_psnr_ssim_u8: scratch too small or misaligned");
  const int Hc = H - 2 * border, Wc = W - 2 * border, Hm = Hc - 2 * ER, Wm = Wc - 2 * ER;
  const int tiles_x = (Wm + ET - 1) / ET, tiles_y = (Hm + ET - 1) / ET;
  const int ntiles = tiles_x * tiles_y;
  double* ssim_partial = reinterpret_cast<double*>(scratch);
  unsigned long long* sq_partial = reinterpret_cast<unsigned long long*>(ssim_partial + ntiles);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  eval_ssim_kernel<<<ntiles, 256, 0, st>>>(a, b, H, W, border, tiles_x, ssim_partial);
  eval_sqdiff_kernel<<<SQ_BLOCKS, 256, 0, st>>>(a, b, H, W, border, sq_partial);
  eval_final_kernel<<<1, 256, 0, st>>>(ssim_partial, ntiles, sq_partial, (double)Hc * Wc, (double)Hm * Wm, out);
  g_ff_launches += 3;
  FF_CHECK_LAUNCH("ff_eval_psnr_ssim_u8");
  return FF_OK;
}
