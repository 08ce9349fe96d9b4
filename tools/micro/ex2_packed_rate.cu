// Micro-benchmark: is the packed half-precision exponential (ex2.approx.f16x2 / ex2.approx.ftz.bf16x2: two elements per
// MUFU instruction) twice as fast per element as ex2.approx.ftz.f32, and what does the whole packed softmax inner loop
// (FADD2 shift, cvt to f16x2, LDS.32 of a packed bias pair, HADD2, ex2.f16x2 -> P pair ready for tcgen05.st) cost?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ex2_packed_rate ex2_packed_rate.cu && ./ex2_packed_rate
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t cvt_h2(float lo, float hi) { uint32_t y; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
__device__ __forceinline__ uint32_t cvt_b2(float lo, float hi) { uint32_t y; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) { uint32_t y; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(y) : "r"(a), "r"(b)); return y; }
__device__ __forceinline__ uint32_t badd2(uint32_t a, uint32_t b) { uint32_t y; asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(y) : "r"(a), "r"(b)); return y; }
// MODE 0: f32 ex2 per element (+ the rescaling FMUL); 1: f16x2 ex2 on packed registers (+ HMUL2); 2: bf16x2 ex2 (+ HMUL2.BF16)
// 3: full f32 loop: 2 LDS.32 + 2 FADD2 + 2 ex2 + pack per pair (the shipped pass 2)
// 4: packed loop: FADD2 (shift) + cvt f16x2 + LDS.32 (bias pair) + HADD2 + ex2.f16x2 per pair
// 5: mode 4 in bf16x2
// 6: mode 4 with the bias added in fp32 from a float2 LDS.64 (aligned pairs): FADD2 + FADD2(LDS.64) + cvt + ex2.f16x2
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  __shared__ __align__(16) float tab[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) tab[i] = MODE == 4 || MODE == 5 ? 0.f : 0.001f * i;
  __syncthreads();
  const uint32_t tp = (uint32_t)__cvta_generic_to_shared(tab) + 8u * (threadIdx.x & 15) + 192u * ((threadIdx.x >> 4) & 1);
  float v[32];
  uint32_t hv[16];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = -0.01f * (i + threadIdx.x % 7);
#pragma unroll
  for (int i = 0; i < 16; ++i) hv[i] = cvt_h2(v[2 * i], v[2 * i + 1]);
  const float2 nsh = make_float2(-0.5f, -0.5f);
  uint32_t accp = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ex2(v[i]) * -0.25f;
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { uint32_t e = ex2_h2(hv[i]); asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(hv[i]) : "r"(e), "r"(0xB400B400u)); }
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { uint32_t e = ex2_b2(hv[i]); asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(hv[i]) : "r"(e), "r"(0xBE80BE80u)); }
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float2 s = __fadd2_rn(make_float2(v[i], v[i + 1]), nsh);
        s = __fadd2_rn(s, make_float2(lds_f32(tp + 4u * 48u * (i >> 4) + 4u * (i & 15)), lds_f32(tp + 4u * 48u * (i >> 4) + 4u * (i & 15) + 4u)));
        const float e0 = ex2(s.x), e1 = ex2(s.y);
        accp ^= cvt_b2(e0, e1);
        v[i] = e0 * -0.25f; v[i + 1] = e1 * -0.25f;
      }
    } else if (MODE == 4 || MODE == 5) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float2 s = __fadd2_rn(make_float2(v[i], v[i + 1]), nsh);
        const uint32_t b = lds_u32(tp + 4u * 48u * (i >> 4) + 2u * (i & 15));
        uint32_t e;
        if (MODE == 4) e = ex2_h2(hadd2(cvt_h2(s.x, s.y), b)); else e = ex2_b2(badd2(cvt_b2(s.x, s.y), b));
        accp ^= e;
        v[i] = v[i] * -0.25f + __uint_as_float(e & 1u); v[i + 1] = v[i + 1] * -0.25f;      // keep a dependence on e
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float2 s = __fadd2_rn(make_float2(v[i], v[i + 1]), nsh);
        float2 b;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(b.x), "=f"(b.y) : "r"(tp + 4u * 48u * (i >> 4) + 4u * (i & 15)));
        s = __fadd2_rn(s, b);
        const uint32_t e = ex2_h2(cvt_h2(s.x, s.y));
        accp ^= e;
        v[i] = v[i] * -0.25f + __uint_as_float(e & 1u); v[i + 1] = v[i + 1] * -0.25f;
      }
    }
  }
  long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += v[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) accp ^= hv[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(accp);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 2000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    k<MODE><<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    k<MODE><<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += h[i];
    c /= 148;
    printf("%-44s warps/SMSP %d: %.2f cycles per element per warp, %.2f per element per SMSP\n", name, warps / 4, c / (iters * 32.0),
           c / (iters * 32.0) / (warps / 4));
  }
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0>("ex2.f32 (+fmul)", out, cyc);
  run<1>("ex2.f16x2 (+hmul2)", out, cyc);
  run<2>("ex2.bf16x2 (+hmul2.bf16)", out, cyc);
  run<3>("shipped loop (f32)", out, cyc);
  run<4>("packed loop f16x2 (lds.32 bias pair)", out, cyc);
  run<5>("packed loop bf16x2", out, cyc);
  run<6>("fp32 bias lds.64, cvt, ex2.f16x2", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
