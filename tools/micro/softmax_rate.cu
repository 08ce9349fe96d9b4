// Micro-benchmark: how fast can W warps per SM sub-partition run the inner loop of the attention softmax
// (sub, bias LDS, add, ex2, bf16 pack) on registers?  Answers whether the exponential pass of window_attention_tc is
// MUFU-bound (8 cycles per warp-wide ex2 per sub-partition) or per-warp latency/issue bound.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_rate softmax_rate.cu && ./softmax_rate
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
// exp2 on the FMA/ALU pipes (Cody-Waite split + cubic, rel. error ~1e-4: below the bf16 rounding of P): 8-9 issue slots
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;          // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float n = t - 12582912.f;
  const float f = x - n;                   // [-0.5, 0.5]
  float p = fmaf(f, 0.0555041087f, 0.2402265070f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
template <int MODE>   // 0: ex2 only; 1: sub+ex2; 2: sub+lds+add+ex2; 3: mode 2 + bf16 pack; 4: mode 3 without ex2 (FADD instead);
                      // 5 / 6 / 7: mode 3 with every 8th / 4th / 2nd exponential on the FMA pipe (ex2_poly)
__global__ void k(float* out, long long* cyc, int iters) {
  __shared__ float tab[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) tab[i] = 0.001f * i;
  __syncthreads();
  const uint32_t tp = (uint32_t)__cvta_generic_to_shared(tab) + 4u * (threadIdx.x & 15) + 192u * ((threadIdx.x >> 4) & 1);
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = -0.01f * (i + threadIdx.x % 7);
  float sh = 0.5f;
  uint32_t accp = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float s = v[i];
      if (MODE >= 1) s = s - sh;
      if (MODE >= 2) s = s + lds_f32(tp + 4u * 48u * (i >> 4) + 4u * (i & 15));
      const bool poly = (MODE == 5 && (i & 7) == 7) || (MODE == 6 && (i & 3) == 3) || (MODE == 7 && (i & 1) == 1);
      v[i] = (MODE == 4) ? s + 1.0f : (poly ? ex2_poly(s) : ex2(s));
    }
    if (MODE >= 3 && MODE != 4 || MODE == 4) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        accp ^= *reinterpret_cast<uint32_t*>(&h);
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = v[i] * -0.25f;   // keep arguments bounded (1 FMUL per element, FMA pipe)
  }
  long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(accp);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 2000;
  for (int warps = 4; warps <= 32; warps *= 2) {      // warps per SM -> warps / 4 per sub-partition
    k<MODE><<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    k<MODE><<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += h[i];
    c /= 148;
    // cycles per warp-wide element-instruction group per sub-partition
    printf("%-34s warps/SMSP %d: %.2f cycles per element per warp, %.2f cycles per element per SMSP\n", name, warps / 4,
           c / (iters * 32.0), c / (iters * 32.0) / (warps / 4));
  }
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0>("ex2 (+fmul)", out, cyc);
  run<1>("sub, ex2", out, cyc);
  run<2>("sub, lds, add, ex2", out, cyc);
  run<3>("sub, lds, add, ex2, pack", out, cyc);
  run<4>("sub, lds, add, add, pack (no ex2)", out, cyc);
  run<5>("mode 3, 1/8 of ex2 on FMA pipe", out, cyc);
  run<6>("mode 3, 1/4 of ex2 on FMA pipe", out, cyc);
  run<7>("mode 3, 1/2 of ex2 on FMA pipe", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
