// Microbenchmark: how fast can one CTA per SM add a [64 r x 256 d] fp32 block (64 KB) into global memory?
//  A: coalesced red.global.add.f32 (warp = 128 contiguous bytes), 64 per thread
//  B: cp.reduce.async.bulk from smem, 4 x 16 KB per block (no staging cost counted)
//  C: red.global.add.v4.f32, [d][r] layout (lane stride = Rtot*4 bytes), 16 per thread
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bench red_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int D = 256, ROWS = 256 * 304;   // 77824 rows x 1 KB = 80 MB
__global__ void __launch_bounds__(256, 1) kA(float* g, int iters, long long* clk) {
  const int d = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int row0 = ((blockIdx.x * 7919 + it * 131) % (ROWS / 64)) * 64;
    float* p = g + (size_t)row0 * D + d;
#pragma unroll 16
    for (int j = 0; j < 64; ++j) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + (size_t)j * D), "f"(1.0f) : "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}
__global__ void __launch_bounds__(256, 1) kB(float* g, int iters, long long* clk) {
  extern __shared__ __align__(1024) uint8_t sm[];
  for (int i = threadIdx.x; i < 32768 / 4; i += 256) reinterpret_cast<float*>(sm)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
    for (int it = 0; it < iters; ++it) {
      const int row0 = ((blockIdx.x * 7919 + it * 131) % (ROWS / 64)) * 64;
      for (int qt = 0; qt < 4; ++qt) {
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                     ::"l"(g + (size_t)(row0 + qt * 16) * D), "r"(s + (qt & 1) * 16384), "r"(16384) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}
__global__ void __launch_bounds__(256, 1) kC(float* g, int iters, long long* clk) {
  const int d = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int img = (blockIdx.x * 7919 + it * 131) % 256, c = it % 4;
    float* p = g + ((size_t)img * D + d) * 304 + c * 64;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p + 4 * j), "f"(1.0f) : "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}
int main() {
  float* g; long long* clk; long long h[148];
  cudaMalloc(&g, (size_t)ROWS * D * 4); cudaMemset(g, 0, (size_t)ROWS * D * 4);
  cudaMalloc(&clk, 148 * 8);
  const int iters = 300;
  cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int v = 0; v < 3; ++v) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a);
      if (v == 0) kA<<<148, 256>>>(g, iters, clk);
      if (v == 1) kB<<<148, 256, 34000>>>(g, iters, clk);
      if (v == 2) kC<<<148, 256>>>(g, iters, clk);
      cudaEventRecord(b); cudaError_t e = cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, a, b);
      cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("variant %c rep %d: %s  %.3f ms, max %lld clk per CTA = %.0f clk per 64 KB block, %.1f GB/s aggregate\n", "ABC"[v], rep,
             cudaGetErrorString(e), ms, mx, (double)mx / iters, 148.0 * iters * 65536 / (ms * 1e-3) / 1e9);
    }
  }
  return 0;
}
