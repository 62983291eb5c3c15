// Micro-benchmark (development aid): cycles per tcgen05.mma (M = 128, K = 16, bf16, SS mode, K-major SWIZZLE_128B operands)
// issued back to back by one thread, as a function of N and of how many independent accumulators the stream alternates
// between.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I kan-tts_b200/csrc -o gpurun_out/mma_rate scripts/micro/mma_rate.cu
#include <cstdio>
#include "tc_common.cuh"
using namespace kt::tc;

__global__ void __launch_bounds__(128, 1) k(int N, int nacc, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x >= 32 && threadIdx.x < 64) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
      const uint32_t a16 = (smem_u32(smem) >> 4) | 0x10000u, b16 = (smem_u32(smem + 32768) >> 4) | 0x10000u;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_bf16_lo(tm + (uint32_t)((it * 8 + j) % nacc) * (uint32_t)N, a16 + 2u * (j & 3), b16 + 2u * (j & 3), idesc, 1u);
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      out[0] = t1 - t0; out[1] = t2 - t0;
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 64;
  for (int N : {16, 32, 64, 128, 256})
    for (int nacc : {1, 2, 4}) {
      if (N * nacc > 512) continue;
      k<<<1, 128, 100 * 1024>>>(N, nacc, iters, d);
      k<<<1, 128, 100 * 1024>>>(N, nacc, iters, d);
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError();
      printf("N=%3d accumulators=%d: issue %.1f cycles/MMA, complete %.1f cycles/MMA (floor %d) %s\n", N, nacc, (double)h[0] / (iters * 8),
             (double)h[1] / (iters * 8), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
