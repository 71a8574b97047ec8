// gradpen.cu — the reduction of the matching-aware gradient penalty (MA-GP), xmc_gan/train_gan.py:244-249:
//   grad = cat(grad_img.view(B,-1), grad_sent.view(B,-1), 1);  d_loss = 2 * mean(sqrt(sum(grad^2, 1))^6)
// One HBM pass: per-sample sums of squares straight from the two gradient tensors (no cat, no grad**2
// temporary), then loss = weight * mean(sumsq^h) with h = p/2 (p = 6 -> sumsq^3, the sqrt never happens).
// Backward (first order; the double backward through netD stays autograd's):
//   d loss / d g[b, :] = grad_out * weight / B * 2h * sumsq[b]^(h-1) * g[b, :]
// Algorithmic bytes: forward B*(n0+n1)*s_in read; backward the same read + the same written.
#include "common.cuh"

namespace xmc {

constexpr int kGpThreads = 256;

__device__ __forceinline__ float sq_sum(float4 v) { return dot4(v, v); }

// partial[b*S + s] = sum of squares of slice s of row b (slice 0 also takes the second tensor's row)
template <typename T>
__global__ void __launch_bounds__(kGpThreads) gradpen_sumsq_kernel(const T* __restrict__ g0, long long n0,
                                                                   const T* __restrict__ g1, long long n1,
                                                                   int S, float* __restrict__ partial) {
  const int b = blockIdx.y, s = blockIdx.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  auto row_sum = [&](const T* row, long long lo, long long hi, bool vec) {
    if (vec) {                                          // 4 elements per load, 4 loads in flight
      long long i = lo / 4 + threadIdx.x;
      const long long e = hi / 4;
      for (; i + 3 * kGpThreads < e; i += 4 * kGpThreads) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld4_nc(row + 4 * (i + u * kGpThreads));
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += sq_sum(v[u]);
      }
      for (; i < e; i += kGpThreads) acc[0] += sq_sum(ld4_nc(row + 4 * i));
    } else {
      for (long long i = lo + threadIdx.x; i < hi; i += kGpThreads) { const float v = ld1(row + i); acc[0] = fmaf(v, v, acc[0]); }
    }
  };
  if (n0 > 0) {
    const T* row = g0 + (size_t)b * n0;
    const bool vec = (n0 % 4 == 0) && ((reinterpret_cast<uintptr_t>(g0) & 15) == 0);
    long long per = (n0 + S - 1) / S;
    per = (per + 3) / 4 * 4;                            // slices start on a vector boundary
    const long long lo = min((long long)s * per, n0), hi = min(lo + per, n0);
    row_sum(row, lo, hi, vec);
  }
  if (n1 > 0 && s == 0) {
    const T* row = g1 + (size_t)b * n1;
    const bool vec = (n1 % 4 == 0) && ((reinterpret_cast<uintptr_t>(g1) & 15) == 0);
    row_sum(row, 0, n1, vec);
  }
  float t = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  t = warp_sum(t);
  __shared__ float sh[kGpThreads / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < kGpThreads / 32; ++w) r += sh[w];
    partial[(size_t)b * S + s] = r;
  }
}

// sumsq[b] = sum_s partial[b][s] (fixed order);  loss = weight * mean_b sumsq[b]^h.  Single CTA.
__global__ void __launch_bounds__(kGpThreads) gradpen_loss_kernel(const float* __restrict__ partial, int B, int S,
                                                                  float h, float weight, float* __restrict__ sumsq,
                                                                  float* __restrict__ loss) {
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += kGpThreads) {
    float s = 0.f;
    for (int k = 0; k < S; ++k) s += partial[(size_t)b * S + k];
    sumsq[b] = s;
    acc += (h == 3.f) ? s * s * s : powf(s, h);
  }
  acc = warp_sum(acc);
  __shared__ float sh[kGpThreads / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int w = 0; w < kGpThreads / 32; ++w) r += sh[w];
    *loss = weight * r / (float)B;
  }
}

// d[b, :] = coef_b * g[b, :],  coef_b = grad_out * weight / B * 2h * sumsq[b]^(h-1)
template <typename T>
__global__ void __launch_bounds__(kGpThreads) gradpen_bwd_kernel(const T* __restrict__ g0, long long n0,
                                                                 const T* __restrict__ g1, long long n1,
                                                                 const float* __restrict__ sumsq,
                                                                 const float* __restrict__ grad_out, float scale, float h,
                                                                 int S, T* __restrict__ d0, T* __restrict__ d1) {
  const int b = blockIdx.y, s = blockIdx.x;
  const float sq = sumsq[b];
  const float coef = __ldg(grad_out) * scale * ((h == 3.f) ? sq * sq : powf(sq, h - 1.f));
  auto row_scale = [&](const T* src, T* dst, long long lo, long long hi, bool vec) {
    if (vec) {
      const long long e = hi / 4;
      for (long long i = lo / 4 + threadIdx.x; i < e; i += kGpThreads) {
        float4 v = ld4_nc(src + 4 * i);
        v.x *= coef; v.y *= coef; v.z *= coef; v.w *= coef;
        st4(dst + 4 * i, v);
      }
    } else {
      for (long long i = lo + threadIdx.x; i < hi; i += kGpThreads) st1(dst + i, ld1(src + i) * coef);
    }
  };
  if (n0 > 0 && d0) {
    const bool vec = (n0 % 4 == 0) && (((reinterpret_cast<uintptr_t>(g0) | reinterpret_cast<uintptr_t>(d0)) & 15) == 0);
    long long per = (n0 + S - 1) / S;
    per = (per + 3) / 4 * 4;
    const long long lo = min((long long)s * per, n0), hi = min(lo + per, n0);
    row_scale(g0 + (size_t)b * n0, d0 + (size_t)b * n0, lo, hi, vec);
  }
  if (n1 > 0 && d1 && s == 0) {
    const bool vec = (n1 % 4 == 0) && (((reinterpret_cast<uintptr_t>(g1) | reinterpret_cast<uintptr_t>(d1)) & 15) == 0);
    row_scale(g1 + (size_t)b * n1, d1 + (size_t)b * n1, 0, n1, vec);
  }
}

static int check_gp(const void* g0, long long n0, const void* g1, long long n1, int B, int dtype, float p, int S) {
  XMC_REQUIRE(B > 0 && n0 >= 0 && n1 >= 0 && n0 + n1 > 0, XMC_ERR_INVALID_ARG, "bad shape B=%d n0=%lld n1=%lld", B, n0, n1);
  XMC_REQUIRE((n0 == 0 || g0) && (n1 == 0 || g1), XMC_ERR_INVALID_ARG, "null gradient pointer");
  XMC_REQUIRE(dtype == XMC_F32 || dtype == XMC_BF16, XMC_ERR_UNSUPPORTED, "dtype");
  XMC_REQUIRE(p >= 2.f, XMC_ERR_INVALID_ARG, "power must be >= 2");
  XMC_REQUIRE(S >= 1 && S <= 65535, XMC_ERR_INVALID_ARG, "bad slice count %d", S);
  return XMC_OK;
}

}  // namespace xmc

using namespace xmc;

extern "C" int xmc_gradnorm_penalty_forward(const void* g0, long long n0, const void* g1, long long n1, int B, int dtype,
                                            float power, float weight, int slices, float* partial, float* sumsq,
                                            float* loss, void* stream) {
  if (int rc = check_gp(g0, n0, g1, n1, B, dtype, power, slices)) return rc;
  XMC_REQUIRE(partial && sumsq && loss, XMC_ERR_INVALID_ARG, "null output pointer");
  dim3 grid(slices, B);
  if (dtype == XMC_F32)
    gradpen_sumsq_kernel<float><<<grid, kGpThreads, 0, as_stream(stream)>>>(static_cast<const float*>(g0), n0, static_cast<const float*>(g1), n1, slices, partial);
  else
    gradpen_sumsq_kernel<__nv_bfloat16><<<grid, kGpThreads, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(g0), n0, static_cast<const __nv_bfloat16*>(g1), n1, slices, partial);
  XMC_RETURN_IF_CUDA(cudaGetLastError());
  gradpen_loss_kernel<<<1, kGpThreads, 0, as_stream(stream)>>>(partial, B, slices, 0.5f * power, weight, sumsq, loss);
  return cuda_fail(cudaGetLastError(), "gradpen_loss_kernel launch");
}

extern "C" int xmc_gradnorm_penalty_backward(const void* g0, long long n0, const void* g1, long long n1, int B, int dtype,
                                             float power, float weight, int slices, const float* sumsq,
                                             const float* grad_out, void* d0, void* d1, void* stream) {
  if (int rc = check_gp(g0, n0, g1, n1, B, dtype, power, slices)) return rc;
  XMC_REQUIRE(sumsq && grad_out, XMC_ERR_INVALID_ARG, "null pointer");
  const float h = 0.5f * power, scale = weight / (float)B * 2.f * h;
  dim3 grid(slices, B);
  if (dtype == XMC_F32)
    gradpen_bwd_kernel<float><<<grid, kGpThreads, 0, as_stream(stream)>>>(static_cast<const float*>(g0), n0, static_cast<const float*>(g1), n1, sumsq, grad_out, scale, h, slices, static_cast<float*>(d0), static_cast<float*>(d1));
  else
    gradpen_bwd_kernel<__nv_bfloat16><<<grid, kGpThreads, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(g0), n0, static_cast<const __nv_bfloat16*>(g1), n1, sumsq, grad_out, scale, h, slices, static_cast<__nv_bfloat16*>(d0), static_cast<__nv_bfloat16*>(d1));
  return cuda_fail(cudaGetLastError(), "gradpen_bwd_kernel launch");
}
