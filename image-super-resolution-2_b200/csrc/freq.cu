// Multi-domain frequency decomposition of the LR tile into 9 bands (reference multi_domain_frequency.py:578-591):
//   DCT (3 bands, 8x8 block DCT-II + zigzag masks, :146-196), DWT (4 sub-bands, db4, reflect pad 7, stride 2, bilinear
//   up-sampling back to the tile size, :251-299), FFT (2 bands, learnable sigmoid mask on the rfft2 spectrum, :352-385).
// Output: fp32 NHWC rows [B*S*S][27], channel = band*3 + c with band order
//   [DCT_low, DCT_mid, DCT_high, DWT_LL, DWT_LH, DWT_HL, DWT_HH, FFT_low, FFT_high].
// The tile is treated as the whole image (block grid, reflect padding and the global FFT are per tile).
#include "ff_common.cuh"
#include "../../include/ffb200.h"

extern long long g_ff_launches;

namespace {

constexpr int NB = 27;

// ---------------------------------------------------------------- DCT
// one thread per (sample, channel, 8x8 block)
__global__ void __launch_bounds__(64) dct_bands_kernel(const float* __restrict__ x, int B, int H, int W,
                                                       const float* __restrict__ dmat,      // [8][8] D[k][n]
                                                       const int* __restrict__ band_of,     // [64] band id of coefficient (i*8+j)
                                                       const float* __restrict__ band_scale, float* __restrict__ out) {
  __shared__ float D[64];
  __shared__ int bo[64];
  D[threadIdx.x] = dmat[threadIdx.x];
  bo[threadIdx.x] = band_of[threadIdx.x];
  __syncthreads();
  // sizes that are not multiples of 8 are reflect-padded on the right / bottom (multi_domain_frequency.py:151-157) and the
  // result is cropped back (:190-193): edge blocks read mirrored pixels and only write the part inside the image
  const int nbx = (W + 7) / 8, nby = (H + 7) / 8;
  const long long idx = (long long)blockIdx.x * 64 + threadIdx.x;
  if (idx >= (long long)B * 3 * nbx * nby) return;
  const int bx = (int)(idx % nbx), by = (int)((idx / nbx) % nby), c = (int)((idx / (nbx * nby)) % 3), b = (int)(idx / (3LL * nbx * nby));
  const float* plane = x + (long long)(b * 3 + c) * H * W;
  float X[64], T[64], Y[64];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int yy = by * 8 + i;
    if (yy >= H) yy = 2 * (H - 1) - yy;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int xx = bx * 8 + j;
      if (xx >= W) xx = 2 * (W - 1) - xx;
      X[i * 8 + j] = plane[(long long)yy * W + xx];
    }
  }
  // T = X * D^T  (T[i][k] = sum_n X[i][n] D[k][n]);  Y = D * T
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float s = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) s += X[i * 8 + n] * D[k * 8 + n];
      T[i * 8 + k] = s;
    }
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) s += D[k * 8 + n] * T[n * 8 + j];
      Y[k * 8 + j] = s;
    }
  for (int band = 0; band < 3; ++band) {
    // X_band = D^T * (Y o M_band) * D
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) s += ((bo[k * 8 + n] == band) ? Y[k * 8 + n] : 0.f) * D[n * 8 + j];
        T[k * 8 + j] = s;
      }
    const float sc = band_scale[band];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) s += D[n * 8 + i] * T[n * 8 + j];
        if (by * 8 + i < H && bx * 8 + j < W) out[((long long)(b * H + by * 8 + i) * W + bx * 8 + j) * NB + band * 3 + c] = s * sc;
      }
  }
}

// ---------------------------------------------------------------- DWT
__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
// sub[b][c][k][Sh][Sw], k: 0 LL, 1 LH, 2 HL, 3 HH
__global__ void __launch_bounds__(128) dwt_subbands_kernel(const float* __restrict__ x, int B, int H, int W,
                                                          const float* __restrict__ lo, const float* __restrict__ hi,
                                                          float* __restrict__ sub) {
  const int Sh = (H + 6) / 2 + 1, Sw = (W + 6) / 2 + 1;      // conv of the 7+7 reflect-padded axis with 8 taps, stride 2
  const long long idx = (long long)blockIdx.x * 128 + threadIdx.x;
  if (idx >= (long long)B * 3 * Sh * Sw) return;
  const int xo = (int)(idx % Sw), yo = (int)((idx / Sw) % Sh);
  const long long plane = idx / ((long long)Sw * Sh);
  const float* p = x + plane * H * W;
  float ll = 0.f, lh = 0.f, hl = 0.f, hh = 0.f;
#pragma unroll
  for (int ty = 0; ty < 8; ++ty) {
    const int yy = reflect_idx(2 * yo + ty - 7, H);
    float rl = 0.f, rh = 0.f;
#pragma unroll
    for (int tx = 0; tx < 8; ++tx) {
      const float v = p[yy * W + reflect_idx(2 * xo + tx - 7, W)];
      rl += v * lo[tx];
      rh += v * hi[tx];
    }
    ll += rl * lo[ty]; lh += rl * hi[ty];
    hl += rh * lo[ty]; hh += rh * hi[ty];
  }
  float* o = sub + plane * 4 * Sh * Sw + (long long)yo * Sw + xo;
  o[0] = ll; o[(long long)Sh * Sw] = lh; o[2LL * Sh * Sw] = hl; o[3LL * Sh * Sw] = hh;
}
__global__ void __launch_bounds__(256) dwt_upsample_kernel(const float* __restrict__ sub, int B, int H, int W,
                                                          const float* __restrict__ subband_scale, float* __restrict__ out) {
  const int Sh = (H + 6) / 2 + 1, Sw = (W + 6) / 2 + 1;      // conv of the 7+7 reflect-padded axis with 8 taps, stride 2
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * H * W) return;
  const int xo = (int)(idx % W), yo = (int)((idx / W) % H), b = (int)(idx / ((long long)W * H));
  const float ry = (float)Sh / H, rx = (float)Sw / W;
  float sy = ry * (yo + 0.5f) - 0.5f, sx = rx * (xo + 0.5f) - 0.5f;
  sy = fmaxf(sy, 0.f); sx = fmaxf(sx, 0.f);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = min(y0 + 1, Sh - 1), x1 = min(x0 + 1, Sw - 1);
  const float ly = sy - y0, lx = sx - x0;
  for (int c = 0; c < 3; ++c)
    for (int k = 0; k < 4; ++k) {
      const float* p = sub + ((long long)(b * 3 + c) * 4 + k) * Sh * Sw;
      const float v = (1.f - ly) * ((1.f - lx) * p[y0 * Sw + x0] + lx * p[y0 * Sw + x1]) + ly * ((1.f - lx) * p[y1 * Sw + x0] + lx * p[y1 * Sw + x1]);
      out[idx * NB + (3 + k) * 3 + c] = v * subband_scale[k];
    }
}

// ---------------------------------------------------------------- FFT (as four DFT passes with a twiddle table)
// pass 1: rows, real -> complex:  X1[p][y][k] = sum_x x[y][x] e^{-2 pi i k x / W},  k = 0..W/2
__global__ void __launch_bounds__(128) fft_rows_fwd_kernel(const float* __restrict__ x, int H, int W, float2* __restrict__ X1) {
  extern __shared__ float sm[];
  float* row = sm;            // [W]
  float* cs = sm + W;         // [W] cos(2 pi n / W)
  float* sn = cs + W;         // [W] sin
  const int y = blockIdx.x, p = blockIdx.y, Wh = W / 2 + 1;
  for (int i = threadIdx.x; i < W; i += 128) {
    row[i] = x[((long long)p * H + y) * W + i];
    float s, c;
    sincospif(2.f * i / W, &s, &c);
    cs[i] = c; sn[i] = s;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < Wh; k += 128) {
    float re = 0.f, im = 0.f;
    int ph = 0;
    for (int n = 0; n < W; ++n) {
      re += row[n] * cs[ph];
      im -= row[n] * sn[ph];
      ph += k; if (ph >= W) ph -= W;
    }
    X1[((long long)p * H + y) * Wh + k] = make_float2(re, im);
  }
}
// pass 2/3: columns, complex -> complex (sign = -1 forward, +1 inverse), optional real mask multiply (forward only)
__global__ void __launch_bounds__(128) fft_cols_kernel(const float2* __restrict__ in, int H, int Wh, float sign,
                                                      const float* __restrict__ mask, float2* __restrict__ outp) {
  extern __shared__ float sm[];
  float* cs = sm;
  float* sn = sm + H;
  for (int i = threadIdx.x; i < H; i += 128) {
    float s, c;
    sincospif(2.f * i / H, &s, &c);
    cs[i] = c; sn[i] = s * sign;
  }
  __syncthreads();
  const int ky = blockIdx.x, p = blockIdx.y;
  for (int k = threadIdx.x; k < Wh; k += 128) {
    float re = 0.f, im = 0.f;
    int ph = 0;
    const float2* col = in + (long long)p * H * Wh + k;
    for (int y = 0; y < H; ++y) {
      const float2 v = col[(long long)y * Wh];
      const float c = cs[ph], s = sn[ph];
      re += v.x * c - v.y * s;
      im += v.x * s + v.y * c;
      ph += ky; if (ph >= H) ph -= H;
    }
    if (mask) { const float m = mask[ky * Wh + k]; re *= m; im *= m; }
    outp[((long long)p * H + ky) * Wh + k] = make_float2(re, im);
  }
}
// pass 4: rows, complex -> real (c2r: imaginary parts of the DC / Nyquist bins do not contribute), then
// low = scale0 * y, high = scale1 * (x - y)   (irfft2(X(1-m)) = x - irfft2(X m) by linearity)
__global__ void __launch_bounds__(128) fft_rows_inv_kernel(const float2* __restrict__ Y1, const float* __restrict__ x, int B, int H, int W,
                                                          float norm, const float* __restrict__ band_scale, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int Wh = W / 2 + 1;
  float* re = sm;             // [Wh]
  float* im = re + Wh;        // [Wh]
  float* cs = im + Wh;        // [W]
  float* sn = cs + W;
  const int y = blockIdx.x, p = blockIdx.y;   // p = b*3 + c
  for (int i = threadIdx.x; i < Wh; i += 128) {
    const float2 v = Y1[((long long)p * H + y) * Wh + i];
    const float w = (i == 0 || (2 * i == W)) ? 1.f : 2.f;       // DC and (even W only) Nyquist bins appear once in the full spectrum
    re[i] = v.x * w; im[i] = v.y * w;
  }
  for (int i = threadIdx.x; i < W; i += 128) {
    float s, c;
    sincospif(2.f * i / W, &s, &c);
    cs[i] = c; sn[i] = s;
  }
  __syncthreads();
  const int b = p / 3, c = p - b * 3;
  for (int n = threadIdx.x; n < W; n += 128) {
    float acc = 0.f;
    int ph = 0;
    for (int k = 0; k < Wh; ++k) {
      acc += re[k] * cs[ph] - im[k] * sn[ph];
      ph += n; if (ph >= W) ph -= W;
    }
    const float low = acc * norm;
    const float xv = x[((long long)p * H + y) * W + n];
    float* o = out + ((long long)(b * H + y) * W + n) * NB;
    o[7 * 3 + c] = low * band_scale[0];
    o[8 * 3 + c] = (xv - low) * band_scale[1];
  }
}

}  // namespace

extern "C" int ff_freq_decompose(const float* lr, int B, int H, int W, const float* dct_mat, const int* dct_band_of,
                                 const float* dct_scale, const float* dwt_lo, const float* dwt_hi, const float* dwt_scale,
                                 const float* fft_mask, const float* fft_scale, float* bands, float* scratch,
                                 size_t scratch_bytes, void* stream) {
  FF_CHECK_ARG(lr && dct_mat && dct_band_of && dct_scale && dwt_lo && dwt_hi && dwt_scale && fft_mask && fft_scale && bands && scratch,
               "ff_freq_decompose: null buffer");
  // any size the reference accepts: reflect padding needs 7 < H, W (DWT, :252) and W <= 4096 keeps the DFT twiddles in 48 KB of smem
  FF_CHECK_ARG(H >= 8 && W >= 8 && H <= 4096 && W <= 4096, "ff_freq_decompose: image %dx%d outside [8, 4096]", H, W);
  const int Wh = W / 2 + 1, Sh = (H + 6) / 2 + 1, Sw = (W + 6) / 2 + 1;
  const size_t need = ((size_t)B * 3 * H * Wh * 2 * 2 + (size_t)B * 3 * 4 * Sh * Sw) * sizeof(float);
  FF_CHECK_ARG(scratch_bytes >= need, "ff_freq_decompose: scratch %zu < %zu", scratch_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float2* bufA = reinterpret_cast<float2*>(scratch);
  float2* bufB = bufA + (size_t)B * 3 * H * Wh;
  float* sub = reinterpret_cast<float*>(bufB + (size_t)B * 3 * H * Wh);
  dct_bands_kernel<<<ff_cdiv((long long)B * 3 * ((H + 7) / 8) * ((W + 7) / 8), 64), 64, 0, st>>>(lr, B, H, W, dct_mat, dct_band_of, dct_scale, bands);
  dwt_subbands_kernel<<<ff_cdiv((long long)B * 3 * Sh * Sw, 128), 128, 0, st>>>(lr, B, H, W, dwt_lo, dwt_hi, sub);
  dwt_upsample_kernel<<<ff_cdiv((long long)B * H * W, 256), 256, 0, st>>>(sub, B, H, W, dwt_scale, bands);
  fft_rows_fwd_kernel<<<dim3(H, B * 3), 128, 3 * W * sizeof(float), st>>>(lr, H, W, bufA);
  fft_cols_kernel<<<dim3(H, B * 3), 128, 2 * H * sizeof(float), st>>>(bufA, H, Wh, -1.f, fft_mask, bufB);
  fft_cols_kernel<<<dim3(H, B * 3), 128, 2 * H * sizeof(float), st>>>(bufB, H, Wh, 1.f, nullptr, bufA);
  fft_rows_inv_kernel<<<dim3(H, B * 3), 128, (2 * Wh + 2 * W) * sizeof(float), st>>>(bufA, lr, B, H, W, 1.0f / ((float)H * W), fft_scale, bands);
  g_ff_launches += 7;
  FF_CHECK_LAUNCH("ff_freq_decompose");
  return FF_OK;
}
