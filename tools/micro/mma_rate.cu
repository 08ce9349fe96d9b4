// Micro-benchmark: issue rate of tcgen05.mma (M=128, K=16, bf16) for several N, operands static in smem.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I image-super-resolution-2_b200/csrc -o tools/micro/mma_rate tools/micro/mma_rate.cu
#include "ff_common.cuh"
#include <cstdlib>
void ff_set_error(const char*, ...) {}

template <int BN, int MODE>
__global__ void __launch_bounds__(320, 1) k(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[8];
  __shared__ uint32_t tbase;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < (16384 + BN * 128) * 4 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (MODE >= 4) {
    // producer/consumer handshake without TMA: warp 0 refills (arrives on full) as soon as a stage is released
    constexpr int S = 4;
    uint64_t* full = bar + 0; uint64_t* empty = bar + 4;   // bar[0..3] full, bar[4..7] empty
    __shared__ __align__(8) uint64_t done;
    if (threadIdx.x == 0) { mbar_init(&done, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x < 32) {
      int st = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&empty[st], ph ^ 1);
        if (elect_one()) mbar_arrive(&full[st]);
        __syncwarp();
        if (++st == S) { st = 0; ph ^= 1; }
      }
    } else if (threadIdx.x < 64) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      const uint32_t d = tbase;
      const uint64_t da0 = umma_desc_k_sw128(smem_u32(smem)), db0 = umma_desc_k_sw128(smem_u32(smem) + 16384);
      constexpr uint64_t STG = (16384 + BN * 128) >> 4;
      int st = 0; uint32_t ph = 0;
      long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (MODE == 6) { if (elect_one()) mbar_wait(&full[st], ph); __syncwarp(); } else mbar_wait(&full[st], ph);
        if (MODE != 5) tc_fence_after();
        const uint64_t da = da0 + st * STG, db = db0 + st * STG;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma_bf16(d, da + 2 * kk, db + 2 * kk, idesc, 1u);
          tc_commit(&empty[st]);
        }
        __syncwarp();
        if (++st == S) { st = 0; ph ^= 1; }
      }
      if (elect_one()) tc_commit(&done);
      __syncwarp();
      mbar_wait(&done, 0);
      long long t1 = clock64();
      if (blockIdx.x == 0 && threadIdx.x == 32) out[0] = t1 - t0;
    } else if (threadIdx.x >= 64) {
      mbar_wait(&done, 0);     // MODE 7: eight more warps polling a barrier, like idle epilogue warps
    }
  } else
  if (MODE == 3) {
    if (threadIdx.x < 32) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      const uint32_t d = tbase;
      const uint64_t da0 = umma_desc_k_sw128(smem_u32(smem)), db0 = umma_desc_k_sw128(smem_u32(smem) + 16384);
      constexpr uint64_t STG = (16384 + BN * 128) >> 4;
      long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const int st = it & 3;
        const uint64_t da = da0 + st * STG, db = db0 + st * STG;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) tc_mma_bf16(d, da + 2 * kk, db + 2 * kk, idesc, 1u);
          tc_commit(&bar[1 + (it & 3)]);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(&bar[0]);
      __syncwarp();
      mbar_wait(&bar[0], 0);
      long long t1 = clock64();
      if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    }
  } else
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
    const uint32_t d = tbase;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int st = it & 3;
      const uint32_t sa = smem_u32(smem) + st * (16384 + BN * 128);
      const uint32_t sb = sa + 16384;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) tc_mma_bf16(d + (MODE == 2 ? (it & 1) * BN : 0), umma_desc_k_sw128(sa + kk * 32), umma_desc_k_sw128(sb + kk * 32), idesc, 1u);
      if (MODE >= 1) tc_commit(&bar[1 + (it & 3)]);
    }
    tc_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

template <int BN, int MODE>
void run(int iters, long long* d_out) {
  const int smem = (16384 + BN * 128) * 4 + 1024;
  cudaFuncSetAttribute(k<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) k<BN, MODE><<<148, MODE == 7 ? 320 : 128, smem>>>(iters, d_out);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<BN, MODE><<<148, MODE == 7 ? 320 : 128, smem>>>(iters, d_out);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cyc; cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  const double mmas = 4.0 * iters;
  printf("BN=%3d mode=%d: %s  %.1f cycles/MMA  %.1f ns/MMA  chip %.0f TFLOP/s\n", BN, MODE, cudaGetErrorString(err), cyc / mmas, ms * 1e6 / mmas,
         148.0 * mmas * 2.0 * 128 * BN * 16 / (ms * 1e-3) / 1e12);
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 8);
  const int iters = 20000;
  run<16, 0>(iters, d_out); run<32, 0>(iters, d_out); run<64, 0>(iters, d_out); run<128, 0>(iters, d_out); run<192, 0>(iters, d_out); run<256, 0>(iters, d_out);
  run<64, 1>(iters, d_out); run<192, 1>(iters, d_out); run<256, 1>(iters, d_out);
  run<64, 2>(iters, d_out); run<192, 2>(iters, d_out); run<256, 2>(iters, d_out);
  run<16, 3>(iters, d_out); run<64, 3>(iters, d_out); run<128, 3>(iters, d_out); run<192, 3>(iters, d_out);
  run<64, 4>(iters, d_out); run<192, 4>(iters, d_out); run<256, 4>(iters, d_out);
  run<64, 5>(iters, d_out); run<192, 5>(iters, d_out);
  run<64, 6>(iters, d_out); run<192, 6>(iters, d_out);
  run<64, 7>(iters, d_out); run<192, 7>(iters, d_out); run<256, 7>(iters, d_out);
  return 0;
}
