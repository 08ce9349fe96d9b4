// Resampling kernels (PyTorch F.interpolate semantics, align_corners=False, no antialias).
#include "ff_common.cuh"
#include "../../include/ffb200.h"

extern long long g_ff_launches;

namespace {

__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  float x = t + 1.f;
  w[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
  x = t;
  w[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 1.f - t;
  w[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
  x = 2.f - t;
  w[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}

// NCHW fp32 [B,C,h,w] -> NHWC fp32 [B*(s*h)*(s*w)][ld], bicubic (a = -0.75, border-clamped taps, not clamped in value)
// The output image may be zero-padded on the right / bottom to Hp x Wp (NAFNet.check_image_size, nafnet_arch.py:219-225).
__global__ void bicubic_up_kernel(const float* __restrict__ x, int B, int C, int h, int w, int s, float* __restrict__ out, int ld, int Hp, int Wp) {
  const int H = Hp, W = Wp;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * H * W) return;
  const int ox = (int)(idx % W), oy = (int)((idx / W) % H), b = (int)(idx / ((long long)W * H));
  if (oy >= h * s || ox >= w * s) {
    for (int c = 0; c < ld; ++c) out[idx * ld + c] = 0.f;
    return;
  }
  const float sy = (oy + 0.5f) / s - 0.5f, sx = (ox + 0.5f) / s - 0.5f;
  const float fy = floorf(sy), fx = floorf(sx);
  float wy[4], wx[4];
  cubic_coeffs(sy - fy, wy);
  cubic_coeffs(sx - fx, wx);
  const int iy = (int)fy, ix = (int)fx;
  for (int c = 0; c < ld; ++c) {
    float acc = 0.f;
    if (c < C) {
      const float* p = x + ((long long)b * C + c) * h * w;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int yy = min(max(iy - 1 + i, 0), h - 1);
        float r = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int xx = min(max(ix - 1 + j, 0), w - 1);
          r += p[yy * w + xx] * wx[j];
        }
        acc += r * wy[i];
      }
    }
    out[idx * ld + c] = acc;
  }
}

}  // namespace

extern "C" int ff_bicubic_up_pad(const float* x, int B, int C, int h, int w, int scale, float* out, int ld, int Hp, int Wp, void* stream) {
  FF_CHECK_ARG(x && out && ld >= C && scale >= 1 && Hp >= h * scale && Wp >= w * scale, "ff_bicubic_up: bad args");
  const long long total = (long long)B * Hp * Wp;
  bicubic_up_kernel<<<ff_cdiv(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, B, C, h, w, scale, out, ld, Hp, Wp);
  ++g_ff_launches;
  FF_CHECK_LAUNCH("ff_bicubic_up");
  return FF_OK;
}
extern "C" int ff_bicubic_up(const float* x, int B, int C, int h, int w, int scale, float* out, int ld, void* stream) {
  return ff_bicubic_up_pad(x, B, C, h, w, scale, out, ld, h * scale, w * scale, stream);
}
