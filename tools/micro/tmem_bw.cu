// TMEM -> register read bandwidth per SM for the tcgen05.ld shapes an epilogue can use (development micro-benchmark).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/tmem_bw tools/micro/tmem_bw.cu && tools/micro/tmem_bw
#include "../../image-super-resolution-2_b200/csrc/ff_common.cuh"
#include <cstdio>

template <int MODE>
__device__ __forceinline__ uint32_t ld_once(uint32_t taddr) {
  uint32_t acc = 0;
  if constexpr (MODE == 0) {        // 32x32b.x16: 32 lanes x 16 columns = 2 KB per warp instruction
    uint32_t v[16];
    tmem_ld16(taddr, v);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= v[i];
  } else if constexpr (MODE == 1) { // 32x32b.x32: 4 KB
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                   "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= v[i];
  } else {                          // 16x256b.x4: 16 lanes x 32 columns = 2 KB per warp instruction (fragment layout)
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= v[i];
  }
  return acc;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) tmem_bw_kernel(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&base_s, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = base_s + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) acc ^= ld_once<MODE>(base + ((it * 32 + (warp >> 2) * 64) & 255));
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(base_s, 512); }
}

template <int MODE>
void run(const char* name, int bytes_per_ld) {
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 1024 * 4);
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 2000;
    tmem_bw_kernel<MODE><<<148, warps * 32>>>(iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i] / 148.0;
    printf("%-14s warps=%2d: %8.0f cycles for %d loads/warp -> %.1f B/clk/SM (%.1f cycles per load per warp) %s\n", name, warps, avg, iters,
           (double)iters * warps * bytes_per_ld / avg, avg / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}

int main() {
  run<0>("32x32b.x16", 2048);
  run<1>("32x32b.x32", 4096);
  run<2>("16x256b.x4", 2048);
  return 0;
}
