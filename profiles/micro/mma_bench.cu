// Microbenchmark: issue rate of back-to-back tcgen05.mma (kind::f16, bf16 -> fp32, M = 128, K = 16) on one CTA per SM, as a
// function of N, of the operand sources (A from shared or tensor memory) and of the shared-memory layouts (K-major / MN-major).
// The question it answers: when is an MMA bound by its shared-memory operand reads instead of the tensor pipe (M*N/256 clk)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../xmc_gan_b200/csrc -o mma_bench mma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace xmc::tc;

template <int N, int A_TMEM, int A_MN, int B_MN, int CHAINS>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ int abort_flag;
  // operands: A [128 x 64] bf16 (16 KB, 4 k-steps), B [256 x 64] bf16 (32 KB); contents are zeros (timing only)
  for (int i = threadIdx.x; i < 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) { abort_flag = 0; mbar_init(&bar, 1); fence_barrier_init(); }
  const int warp = warp_index();
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const WaitCtx wc{&abort_flag, nullptr};
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t sa = smem_u32(smem), sb = sa + 16384;
      constexpr uint32_t idesc = idesc_bf16(128, N, A_MN != 0, B_MN != 0);
      const Desc da0 = A_MN ? make_desc(sa, 8192, 1024) : make_desc(sa, 16, 1024);
      const Desc db0 = B_MN ? make_desc(sb, 8192, 1024) : make_desc(sb, 16, 1024);
      constexpr uint32_t astep = (A_MN ? 2048 : 32) >> 4, bstep = (B_MN ? 2048 : 32) >> 4;
      // accumulators: chain j at columns j * N (CHAINS * N <= 256); A in tensor memory at column 256 (8 columns per k-step)
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
          for (int j = 0; j < CHAINS; ++j) {
            if (A_TMEM) mma_ts(tmem + j * N, tmem + 256 + ks * 8, db0 + ks * bstep, idesc, true);
            else mma_ss(tmem + j * N, da0 + ks * astep, db0 + ks * bstep, idesc, true);
          }
        }
      }
      mma_commit(&bar);
      mbar_wait(&bar, 0, wc, 1);
      out[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}


// MMA streams of one 128 x 64 x 256 block of the word-region backward (operands are zeros: timing only), issued back to back:
//  STREAM 0 — the kernel as built (per 64-region chunk): S, W = 2 x 16 SS N=64; dQ = 4 SS N=256; dK^T = 2 x 2 x 8 SS N=64
//  STREAM 1 — the transposed formulation of DESIGN 4.6 (per 128 regions x 64 words): S^T|W^T = 16 SS N=128;
//             dK += X^T Q + Y^T Chat = 2 x 4 TS N=256; dQ^T = 2 x 8 SS N=64
template <int STREAM>
__global__ void __launch_bounds__(128, 1) ks(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ int abort_flag;
  for (int i = threadIdx.x; i < 131072 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) { abort_flag = 0; mbar_init(&bar, 1); fence_barrier_init(); }
  const int warp = warp_index();
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const WaitCtx wc{&abort_flag, nullptr};
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t sa = smem_u32(smem), sb = sa + 65536;          // A region 64 KB, B region 64 KB
      const Desc ak = make_desc(sa, 16, 1024), amn = make_desc(sa, 16384, 1024);
      const Desc bk = make_desc(sb, 16, 1024), bmn = make_desc(sb, 8192, 1024);
      constexpr uint32_t i64 = idesc_bf16(128, 64, false, false), i64mn = idesc_bf16(128, 64, true, true);
      constexpr uint32_t i128 = idesc_bf16(128, 128, false, false), i256 = idesc_bf16(128, 256, false, true);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (STREAM == 0) {
#pragma unroll
          for (int k = 0; k < 16; ++k) mma_ss(tmem + 256, ak + (((k >> 2) * 16384 + (k & 3) * 32) >> 4), bk + (((k >> 2) * 8192 + (k & 3) * 32) >> 4), i64, true);
#pragma unroll
          for (int k = 0; k < 16; ++k) mma_ss(tmem + 320, ak + (((k >> 2) * 16384 + (k & 3) * 32) >> 4), bk + (((k >> 2) * 8192 + (k & 3) * 32) >> 4), i64, true);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ss(tmem + 0, ak + ((k * 32) >> 4), bmn + ((k * 2048) >> 4), i256, true);
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k = 0; k < 16; ++k) mma_ss(tmem + 384 + h * 64, amn + (((k & 7) * 2048) >> 4), bmn + (((k & 7) * 2048) >> 4), i64mn, true);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) mma_ss(tmem + 256, ak + (((k >> 2) * 16384 + (k & 3) * 32) >> 4), bk + (((k >> 2) * 16384 + (k & 3) * 32) >> 4), i128, true);
#pragma unroll
          for (int k = 0; k < 8; ++k) mma_ts(tmem + 0, tmem + 256 + (k & 3) * 8, bmn + (((k & 3) * 2048) >> 4), i256, true);
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_ss(tmem + 384 + h * 64, amn + ((k * 2048) >> 4), bmn + ((k * 2048) >> 4), i64mn, true);
        }
      }
      mma_commit(&bar);
      mbar_wait(&bar, 0, wc, 1);
      out[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}


// Operand reuse between CONSECUTIVE MMAs (N = 64, SS): pairs of MMAs into two accumulators that share A (different B), share B
// (different A) or share nothing.  MODE 0: nothing shared, 1: same A, 2: same B, 3: the hi/lo triple of the fp32 path (ah*bh, ah*bl, al*bh)
template <int MODE>
__global__ void __launch_bounds__(128, 1) kr(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ int abort_flag;
  for (int i = threadIdx.x; i < 131072 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) { abort_flag = 0; mbar_init(&bar, 1); fence_barrier_init(); }
  const int warp = warp_index();
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const WaitCtx wc{&abort_flag, nullptr};
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t sa = smem_u32(smem), sb = sa + 65536;
      const Desc a0 = make_desc(sa, 16, 1024), a1 = make_desc(sa + 32768, 16, 1024);
      const Desc b0 = make_desc(sb, 16, 1024), b1 = make_desc(sb + 32768, 16, 1024);
      constexpr uint32_t i64 = idesc_bf16(128, 64, false, false);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t o = (k * 32) >> 4;
          if (MODE == 0) { mma_ss(tmem, a0 + o, b0 + o, i64, true); mma_ss(tmem + 64, a1 + o, b1 + o, i64, true); }
          if (MODE == 1) { mma_ss(tmem, a0 + o, b0 + o, i64, true); mma_ss(tmem + 64, a0 + o, b1 + o, i64, true); }
          if (MODE == 2) { mma_ss(tmem, a0 + o, b0 + o, i64, true); mma_ss(tmem + 64, a1 + o, b0 + o, i64, true); }
          if (MODE == 3) { mma_ss(tmem, a0 + o, b0 + o, i64, true); mma_ss(tmem, a0 + o, b1 + o, i64, true); mma_ss(tmem, a1 + o, b0 + o, i64, true); }
        }
      }
      mma_commit(&bar);
      mbar_wait(&bar, 0, wc, 1);
      out[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// weight-stationary form: tcgen05.mma.ws keeps the B tile in a collector buffer; a pair of MMAs that share B marks it fill / lastuse
__device__ __forceinline__ void mma_ws(uint32_t d, Desc a, Desc b, uint32_t idesc, int mode) {
  if (mode == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
  else if (mode == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(128, 1) kw(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar; __shared__ uint32_t slot; __shared__ int abort_flag;
  for (int i = threadIdx.x; i < 131072 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) { abort_flag = 0; mbar_init(&bar, 1); fence_barrier_init(); }
  const int warp = warp_index();
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const WaitCtx wc{&abort_flag, nullptr};
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t sa = smem_u32(smem), sb = sa + 65536;
      const Desc a0 = make_desc(sa, 16, 1024), a1 = make_desc(sa + 32768, 16, 1024), b0 = make_desc(sb, 16, 1024);
      constexpr uint32_t i64 = idesc_bf16(128, 64, false, false);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t o = (k * 32) >> 4;
          if (MODE == 0) { mma_ws(tmem, a0 + o, b0 + o, i64, 2); mma_ws(tmem + 64, a1 + o, b0 + o, i64, 2); }      // ws, no hints
          if (MODE == 1) { mma_ws(tmem, a0 + o, b0 + o, i64, 0); mma_ws(tmem + 64, a1 + o, b0 + o, i64, 1); }      // ws, fill / lastuse
        }
      }
      mma_commit(&bar);
      mbar_wait(&bar, 0, wc, 1);
      out[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

long long* d;
template <int N, int A_TMEM, int A_MN, int B_MN, int CHAINS>
void run() {
  long long h[148];
  const int iters = 256;
  cudaFuncSetAttribute(k<N, A_TMEM, A_MN, B_MN, CHAINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 52000);
  for (int rep = 0; rep < 2; ++rep) {
    k<N, A_TMEM, A_MN, B_MN, CHAINS><<<148, 128, 52000>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1LL << 60;
  for (int i = 0; i < 148; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
  const double n = (double)iters * 4 * CHAINS;
  const double bytes = (A_TMEM ? 0 : 4096) + N * 32.0;
  printf("N=%3d A=%s/%s B=%s chains=%d: %6.1f clk per MMA (fastest CTA %6.1f), tensor-pipe nominal %3.0f, smem operand bytes %5.0f -> %5.1f B/clk\n", N,
         A_TMEM ? "tmem" : "smem", A_MN ? "MN" : "K ", B_MN ? "MN" : "K ", CHAINS, mx / n, mn / n, 128.0 * N / 256, bytes, bytes / (mx / n));
}

int main() {
  cudaMalloc(&d, 148 * 8);
  run<64, 0, 0, 0, 1>(); run<64, 0, 0, 0, 2>(); run<64, 0, 1, 1, 1>(); run<64, 0, 1, 1, 2>(); run<64, 0, 0, 1, 1>(); run<64, 1, 0, 0, 1>(); run<64, 1, 0, 1, 1>();
  run<128, 0, 0, 0, 1>(); run<128, 0, 1, 1, 1>(); run<128, 0, 0, 0, 2>(); run<128, 1, 0, 0, 1>();
  run<256, 0, 0, 0, 1>(); run<256, 0, 0, 1, 1>(); run<256, 0, 1, 0, 1>(); run<256, 1, 0, 1, 1>();
  run<48, 0, 0, 0, 1>(); run<32, 0, 0, 0, 1>(); run<16, 0, 0, 0, 1>();
  for (int st = 0; st < 2; ++st) {
    long long h[148];
    const int iters = 64;
    auto kern = st == 0 ? ks<0> : ks<1>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 133000);
    for (int rep = 0; rep < 2; ++rep) {
      kern<<<148, 128, 133000>>>(iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%s: %.0f clk per 128 x 64 x 256 block (5 GEMM units; tensor-pipe nominal 2560)\n",
           st == 0 ? "MMA stream of the backward as built   (32 SS N=64 | 4 SS N=256 | 32 SS N=64)" :
                     "MMA stream of the transposed backward (16 SS N=128 | 8 TS N=256 | 16 SS N=64)", (double)mx / iters);
  }
  for (int md = 0; md < 4; ++md) {
    long long h[148];
    const int iters = 256;
    auto kern = md == 0 ? kr<0> : md == 1 ? kr<1> : md == 2 ? kr<2> : kr<3>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 133000);
    for (int rep = 0; rep < 2; ++rep) {
      kern<<<148, 128, 133000>>>(iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const char* names[4] = {"consecutive N=64 SS MMAs sharing nothing", "consecutive N=64 SS MMAs sharing A      ", "consecutive N=64 SS MMAs sharing B      ",
                            "hi/lo triple (ah*bh, ah*bl, al*bh)      "};
    printf("%s: %.1f clk per MMA\n", names[md], (double)mx / (iters * 4 * (md == 3 ? 3 : 2)));
  }
  for (int md = 0; md < 2; ++md) {
    long long h[148];
    auto kern = md == 0 ? kw<0> : kw<1>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 133000);
    for (int rep = 0; rep < 2; ++rep) {
      kern<<<148, 128, 133000>>>(256, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%s: %.1f clk per MMA\n", md == 0 ? "tcgen05.mma.ws N=64 SS pairs sharing B, no collector hints      " :
                                               "tcgen05.mma.ws N=64 SS pairs sharing B, collector b0 fill/lastuse", (double)mx / (256 * 8));
  }
  return 0;
}
