// common.cuh — shared helpers for libxmcloss (sm_100a).  Error plumbing for the C ABI,
// vector loads/stores for fp32 / bf16 storage, warp reductions and the online
// log-sum-exp accumulator used by every InfoNCE statistic.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/xmc_loss.h"

namespace xmc {

constexpr float kEps = 1e-12f;  // F.normalize eps, xmc_gan/train_gan.py:88-89
constexpr int kWarp = 32;

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define XMC_RETURN_IF_CUDA(expr)                              \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::xmc::cuda_fail(_e, #expr); \
  } while (0)

#define XMC_REQUIRE(cond, status, ...) \
  do {                                 \
    if (!(cond)) {                     \
      ::xmc::set_error(__VA_ARGS__);   \
      return (status);                 \
    }                                  \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- storage-type helpers -------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  uint2 raw = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 lo = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
  __nv_bfloat162 hi = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
  return make_float4(a.x, a.y, b.x, b.y);
}
// read-only (non-coherent) variants: the compiler may hoist them over stores to other buffers
__device__ __forceinline__ float4 ld4_nc(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4_nc(const __nv_bfloat16* p) {
  uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
  __nv_bfloat162 lo = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
  __nv_bfloat162 hi = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&lo);
  raw.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = raw;
}
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float dot4(float4 a, float4 b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ void fma4(float4& acc, float s, float4 v) {
  acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y);
  acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}

// ---- warp reductions ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the 16 lanes that share (lane >> 4)
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Reduce ROWS per-lane partial values across the warp with ROWS-1 + log2(32/ROWS) shuffles
// instead of 5*ROWS.  On return every lane holds the full sum of value index owner_row<ROWS>(lane).
template <int ROWS>
__device__ __forceinline__ int owner_row(int lane);
template <>
__device__ __forceinline__ int owner_row<8>(int lane) {
  return ((lane >> 4) & 1) | (((lane >> 3) & 1) << 1) | (((lane >> 2) & 1) << 2);
}
template <>
__device__ __forceinline__ int owner_row<4>(int lane) {
  return ((lane >> 4) & 1) | (((lane >> 3) & 1) << 1);
}

template <>
__device__ __forceinline__ int owner_row<2>(int lane) {
  return (lane >> 4) & 1;
}

template <int N>
__device__ __forceinline__ void fold_step(float (&v)[8], int lane, int bit) {
  // pairs (v[2k], v[2k+1]) -> v[k]; lanes with `bit` clear keep the even member.
  const bool hi = (lane & bit) != 0;
#pragma unroll
  for (int k = 0; k < N / 2; ++k) {
    float keep = hi ? v[2 * k + 1] : v[2 * k];
    float send = hi ? v[2 * k] : v[2 * k + 1];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
  }
}

template <int ROWS>
__device__ __forceinline__ float warp_multi_sum(float (&v)[8], int lane) {
  static_assert(ROWS == 8 || ROWS == 4 || ROWS == 2, "ROWS");
  if (ROWS == 2) {
    fold_step<2>(v, lane, 16);
    float t = v[0];
    t += __shfl_xor_sync(0xffffffffu, t, 8);
    t += __shfl_xor_sync(0xffffffffu, t, 4);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    return t;
  } else if (ROWS == 8) {
    fold_step<8>(v, lane, 16);
    fold_step<4>(v, lane, 8);
    fold_step<2>(v, lane, 4);
    float t = v[0];
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    return t;
  } else {
    fold_step<4>(v, lane, 16);
    fold_step<2>(v, lane, 8);
    float t = v[0];
    t += __shfl_xor_sync(0xffffffffu, t, 4);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    return t;
  }
}

// ---- online statistics of one row / column of an InfoNCE problem ------------------------------
struct Stat {
  float m;    // running max of z
  float s;    // sum exp(z - m)
  float sl;   // sum labels
  float slz;  // sum labels * z
  __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; sl = 0.f; slz = 0.f; }
  __device__ __forceinline__ void add(float z, float l) {
    if (z > m) { s = s * __expf(m - z) + 1.f; m = z; }   // exp(-inf)=0 on first element
    else       { s += __expf(z - m); }
    sl += l;
    slz = fmaf(l, z, slz);
  }
  __device__ __forceinline__ void merge(const Stat& o) {
    if (o.m == -INFINITY) return;
    if (m == -INFINITY) { *this = o; return; }
    float mm = fmaxf(m, o.m);
    s = s * __expf(m - mm) + o.s * __expf(o.m - mm);
    m = mm; sl += o.sl; slz += o.slz;
  }
  __device__ __forceinline__ float lse() const { return m + __logf(s); }
};

__device__ __forceinline__ Stat shfl_xor_stat(const Stat& a, int o) {
  Stat b;
  b.m = __shfl_xor_sync(0xffffffffu, a.m, o);
  b.s = __shfl_xor_sync(0xffffffffu, a.s, o);
  b.sl = __shfl_xor_sync(0xffffffffu, a.sl, o);
  b.slz = __shfl_xor_sync(0xffffffffu, a.slz, o);
  return b;
}

}  // namespace xmc
