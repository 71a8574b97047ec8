// prep.cu — layout/normalisation prologue and epilogue of the word-region loss, and the
// per-caption masked log-sum-exp that turns word relevances into the score matrix.
//
// The reference's encoders hand over words as [B, D, T] and the discriminator feature map as
// [B, D, H, W] (xmc_gan/model/encoder.py:68,140): the contraction dimension D is the SLOW one.
// The tensor-core kernels want unit rows with D contiguous, so one HBM-bound pass does
// F.normalize (train_gan.py:88-89 convention) + transpose + optional bf16 cast through a padded
// shared-memory tile (reads coalesced along L, writes coalesced along D).
#include <type_traits>

#include "common.cuh"

namespace xmc {

constexpr int kLT = 32;  // (b, 32 consecutive l) per CTA

// Compaction of the word rows: captions are padded to T words (mask non-zero = padding,
// encoder.py:61,149); padded words never contribute (their relevance is excluded from the
// log-sum-exp and their gradient is zero), so the tensor-core kernels work on the valid rows only.
// row_of[c*T+t] = index of the word among the valid ones (caption-major order) or -1;
// cap_ptr[c] = first compact row of caption c, cap_ptr[Bc] = number of valid rows.  One CTA.
__global__ void __launch_bounds__(1024) word_rows_compact_kernel(const uint8_t* __restrict__ mask, int n, int T,
                                                                  int* __restrict__ row_of, int* __restrict__ cap_ptr) {
  __shared__ int warp_tot[32];
  __shared__ int base_sh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_sh = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    const int v = (i < n && !mask[i]) ? 1 : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;                          // inclusive over warps
    }
    __syncthreads();
    const int base = base_sh;
    const int excl = base + (warp ? warp_tot[warp - 1] : 0) + incl - v;
    if (i < n) {
      row_of[i] = v ? excl : -1;
      if (i % T == 0) cap_ptr[i / T] = excl;
    }
    __syncthreads();
    if (threadIdx.x == 0) base_sh = base + warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) cap_ptr[n / T] = base_sh;
}

// x[B,D,L] -> xn[B,Lpad,D], norm[B,Lpad].  256 threads; dynamic smem D*(kLT+1) floats.
// row_of (nullable, needs Lpad == L): row (b,l) goes to xn[row_of[b*L+l]] and is skipped when negative.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) norm_tr_kernel(const TI* __restrict__ x, int D, int L, int Lpad,
                                                       const int* __restrict__ row_of,
                                                       TO* __restrict__ xn, float* __restrict__ norm) {
  extern __shared__ float tile[];                 // [D][kLT+1]
  __shared__ float part[8][kLT];
  __shared__ float inv_sh[kLT];
  const int b = blockIdx.y, l0 = blockIdx.x * kLT;
  const int lx = threadIdx.x & 31, dy = threadIdx.x >> 5;
  const int l = l0 + lx;
  float ss = 0.f;
  for (int d = dy; d < D; d += 8) {
    float v = (l < L) ? ld1(x + ((size_t)b * D + d) * L + l) : 0.f;
    tile[d * (kLT + 1) + lx] = v;
    ss = fmaf(v, v, ss);
  }
  part[dy][lx] = ss;
  __syncthreads();
  if (dy == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][lx];
    float n = fmaxf(sqrtf(t), kEps);
    inv_sh[lx] = 1.f / n;
    if (l < Lpad) norm[(size_t)b * Lpad + l] = (l < L) ? n : 0.f;
  }
  __syncthreads();
  // write rows: warp dy handles rows dy, dy+8, ...; lanes run along d
  for (int r = dy; r < kLT; r += 8) {
    const int lr = l0 + r;
    if (lr >= Lpad) break;
    const float inv = inv_sh[r];
    long long orow = (long long)b * Lpad + lr;
    if (row_of) {
      orow = row_of[orow];
      if (orow < 0) continue;
    }
    TO* dst = xn + (size_t)orow * D;
    for (int d = lx; d < D; d += 32) st1(dst + d, tile[d * (kLT + 1) + r] * inv);
  }
}

// dx[b,d,l] = (dxn[l,d] - xh[l,d] * <xh_l, dxn_l>) / norm_l + dnorm_l * xh[l,d]
template <typename TX, typename TO>
__global__ void __launch_bounds__(256) norm_tr_bwd_kernel(const TX* __restrict__ xn, const float* __restrict__ norm,
                                                           const float* __restrict__ dxn, const float* __restrict__ dnorm,
                                                           int D, int L, int Lpad, const int* __restrict__ row_of,
                                                           const int* __restrict__ error_word, TO* __restrict__ dx) {
  extern __shared__ float tile[];                 // [D][kLT+1] holds the finished dx tile
  // a kernel upstream (tcgen05 backward) timed out: its accumulators are garbage, so the gradient is NaN
  const float poison = (error_word && __ldg(error_word) != 0) ? __int_as_float(0x7fc00000) : 0.f;
  const int b = blockIdx.y, l0 = blockIdx.x * kLT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < kLT; r += 8) {
    const int l = l0 + r;
    long long irow = (long long)b * Lpad + l;
    if (l < L && row_of) irow = row_of[irow];          // compact rows: a dropped (padding) word has no gradient
    if (l >= L || irow < 0) {
      for (int d = lane; d < D; d += 32) tile[d * (kLT + 1) + r] = 0.f;
      continue;
    }
    const size_t row = (size_t)irow * D;
    const float n = norm[(size_t)b * Lpad + l];
    const bool clamped = n <= kEps;               // x/eps is linear: no projection, no norm path
    const float inv = 1.f / n;
    const float dn = (dnorm && !clamped) ? dnorm[(size_t)b * Lpad + l] : 0.f;
    if (D % 128 == 0 && D <= 512) {
      // one pass: each lane keeps its 4-element groups of the row in registers (128-bit loads)
      float4 xh[4], g[4];
      float proj = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int d = c * 128 + lane * 4;
        if (d < D) {
          xh[c] = ld4_nc(xn + row + d);
          g[c] = __ldg(reinterpret_cast<const float4*>(dxn + row + d));
          proj += dot4(xh[c], g[c]);
        }
      }
      proj = clamped ? 0.f : warp_sum(proj);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int d = c * 128 + lane * 4;
        if (d < D) {
          tile[(d + 0) * (kLT + 1) + r] = (g[c].x - xh[c].x * proj) * inv + dn * xh[c].x;
          tile[(d + 1) * (kLT + 1) + r] = (g[c].y - xh[c].y * proj) * inv + dn * xh[c].y;
          tile[(d + 2) * (kLT + 1) + r] = (g[c].z - xh[c].z * proj) * inv + dn * xh[c].z;
          tile[(d + 3) * (kLT + 1) + r] = (g[c].w - xh[c].w * proj) * inv + dn * xh[c].w;
        }
      }
      continue;
    }
    float proj = 0.f;
    for (int d = lane; d < D; d += 32) proj = fmaf(ld1(xn + row + d), dxn[row + d], proj);
    proj = warp_sum(proj);
    if (clamped) proj = 0.f;
    for (int d = lane; d < D; d += 32) {
      float xh = ld1(xn + row + d);
      tile[d * (kLT + 1) + r] = (dxn[row + d] - xh * proj) * inv + dn * xh;
    }
  }
  __syncthreads();
  const int l = l0 + lane;
  if (l < L)
    for (int d = warp; d < D; d += 8) st1(dx + ((size_t)b * D + d) * L + l, tile[d * (kLT + 1) + lane] + poison);
}

// ---- bf16 fast paths (D = 128 or 256: the shapes of the tcgen05 word-region kernels) ----------------------
// Same arithmetic, in the same order, as the generic kernels above (results are bit-identical); what changes
// is how the bytes move: every global load of a thread is issued before the first use (16 two-byte loads /
// four rows of 128-bit loads in flight), the staging tile holds bf16 with an odd 32-bit-word pitch so both
// the l-major and the d-major side are bank-conflict-free, and rows leave as 4-byte (bf16x2) stores, one
// full 128-byte line per warp instruction.

__device__ __forceinline__ float bf16_bits_to_float(unsigned short h) { return __uint_as_float((uint32_t)h << 16); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int D>
__global__ void __launch_bounds__(256) norm_tr_bf16_kernel(const __nv_bfloat16* __restrict__ x, int L, int Lpad,
                                                            const int* __restrict__ row_of,
                                                            __nv_bfloat16* __restrict__ xn, float* __restrict__ norm) {
  constexpr int kW = D / 2 + 1;                   // tile row pitch in 32-bit words (odd)
  constexpr int kPer = D / 8;                     // d values per thread
  constexpr int kBatch = 16;
  __shared__ uint32_t tile[kLT * kW];             // [l][d] raw bf16 inputs
  __shared__ float part[8][kLT];
  __shared__ float inv_sh[kLT];
  unsigned short* tile16 = reinterpret_cast<unsigned short*>(tile);
  const int b = blockIdx.y, l0 = blockIdx.x * kLT;
  const int lx = threadIdx.x & 31, dy = threadIdx.x >> 5;
  const int l = l0 + lx;
  const bool in = l < L;
  // lanes beyond L read the last valid column (no predicated loads); their tile rows are never stored
  const unsigned short* src = reinterpret_cast<const unsigned short*>(x) + ((size_t)b * D + dy) * L + (in ? l : L - 1);
  unsigned long long addr = reinterpret_cast<unsigned long long>(src);
  const unsigned long long step = 16ull * (unsigned)L;          // bytes between the d values of a thread
  float ss = 0.f;
#pragma unroll
  for (int i0 = 0; i0 < kPer; i0 += kBatch) {
    unsigned short v[kBatch];
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {          // opaque pointer bump: two integer ops per load, not five
      asm volatile("ld.global.nc.u16 %0, [%1];\n\tadd.u64 %1, %1, %2;" : "=h"(v[i]), "+l"(addr) : "l"(step));
    }
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const float f = bf16_bits_to_float(v[i]);
      tile16[lx * (2 * kW) + dy + 8 * (i0 + i)] = v[i];
      ss = fmaf(f, f, ss);
    }
  }
  part[dy][lx] = ss;
  __syncthreads();
  if (dy == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][lx];
    float n = fmaxf(sqrtf(t), kEps);
    inv_sh[lx] = 1.f / n;
    if (l < Lpad) norm[(size_t)b * Lpad + l] = in ? n : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kLT / 8; ++k) {
    const int r = dy + 8 * k;
    const int lr = l0 + r;
    if (lr >= Lpad) break;
    const float inv = inv_sh[r];
    long long orow = (long long)b * Lpad + lr;
    if (row_of) {
      orow = row_of[orow];
      if (orow < 0) continue;
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(xn + (size_t)orow * D);
    const bool pad_row = lr >= L;                 // zero rows up to Lpad
#pragma unroll
    for (int j = 0; j < D / 64; ++j) {
      const uint32_t w = tile[r * kW + lx + 32 * j];
      dst[lx + 32 * j] = pad_row ? 0u : pack_bf16x2(__uint_as_float(w << 16) * inv, __uint_as_float(w & 0xffff0000u) * inv);
    }
  }
}

template <typename TO> struct BwdTile;
template <> struct BwdTile<__nv_bfloat16> {      // bf16 tile, odd word pitch
  static constexpr int kWordsPerD2 = 1;           // tile words = kLT * (D * kWordsPerD2 / 2 + 1)
  template <int D> static __device__ __forceinline__ void put4(uint32_t* t, int r, int d, float a, float b, float c, float e) {
    t[r * (D / 2 + 1) + d / 2] = pack_bf16x2(a, b);
    t[r * (D / 2 + 1) + d / 2 + 1] = pack_bf16x2(c, e);
  }
  template <int D> static __device__ __forceinline__ __nv_bfloat16 get(const uint32_t* t, int r, int d) {
    return reinterpret_cast<const __nv_bfloat16*>(t)[r * (D + 2) + d];
  }
};
template <> struct BwdTile<float> {              // fp32 tile, odd word pitch
  static constexpr int kWordsPerD2 = 2;
  template <int D> static __device__ __forceinline__ void put4(uint32_t* t, int r, int d, float a, float b, float c, float e) {
    float* f = reinterpret_cast<float*>(t) + r * (D + 1) + d;
    f[0] = a; f[1] = b; f[2] = c; f[3] = e;
  }
  template <int D> static __device__ __forceinline__ float get(const uint32_t* t, int r, int d) {
    return reinterpret_cast<const float*>(t)[r * (D + 1) + d];
  }
};

template <int D, typename TO>
__global__ void __launch_bounds__(256, 3) norm_tr_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ xn, const float* __restrict__ norm,
                                                                const float* __restrict__ dxn, const float* __restrict__ dnorm,
                                                                int L, int Lpad, const int* __restrict__ row_of,
                                                                const int* __restrict__ error_word, TO* __restrict__ dx) {
  constexpr int kC = D / 128;                     // 128-bit groups per lane and row
  const float poison = (error_word && __ldg(error_word) != 0) ? __int_as_float(0x7fc00000) : 0.f;
  constexpr int kRows = kLT / 8;                  // rows per warp
  using Tile = BwdTile<TO>;
  __shared__ uint32_t tile[kLT * (D * Tile::kWordsPerD2 / 2 + 1)];   // [l][d] finished dx values
  const int b = blockIdx.y, l0 = blockIdx.x * kLT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 xh[kRows][kC], g[kRows][kC];
  float n[kRows], dn[kRows];
  bool live[kRows];
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    const int l = l0 + warp + 8 * k;
    long long irow = (long long)b * Lpad + l;
    if (l < L && row_of) irow = row_of[irow];
    live[k] = l < L && irow >= 0;
    const size_t row = live[k] ? (size_t)irow * D : 0;
    n[k] = live[k] ? norm[(size_t)b * Lpad + l] : 1.f;
    dn[k] = (live[k] && dnorm) ? dnorm[(size_t)b * Lpad + l] : 0.f;
#pragma unroll
    for (int c = 0; c < kC; ++c) {
      const int d = c * 128 + lane * 4;
      if (live[k]) {
        xh[k][c] = ld4_nc(xn + row + d);
        g[k][c] = __ldg(reinterpret_cast<const float4*>(dxn + row + d));
      } else {
        xh[k][c] = make_float4(0.f, 0.f, 0.f, 0.f);
        g[k][c] = xh[k][c];
      }
    }
  }
  float proj[kRows];
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    proj[k] = 0.f;
#pragma unroll
    for (int c = 0; c < kC; ++c) proj[k] += dot4(xh[k][c], g[k][c]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < kRows; ++k) proj[k] += __shfl_xor_sync(0xffffffffu, proj[k], o);
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    const int r = warp + 8 * k;
    const bool clamped = n[k] <= kEps;            // x/eps is linear: no projection, no norm path
    const float inv = 1.f / n[k];
    const float pj = clamped ? 0.f : proj[k];
    const float dnk = clamped ? 0.f : dn[k];
#pragma unroll
    for (int c = 0; c < kC; ++c) {
      const float4 X = xh[k][c], G = g[k][c];
      Tile::template put4<D>(tile, r, c * 128 + lane * 4,
                             (G.x - X.x * pj) * inv + dnk * X.x + poison, (G.y - X.y * pj) * inv + dnk * X.y + poison,
                             (G.z - X.z * pj) * inv + dnk * X.z + poison, (G.w - X.w * pj) * inv + dnk * X.w + poison);
    }
  }
  __syncthreads();
  const int l = l0 + lane;
  if (l < L) {
    TO* out = dx + ((size_t)b * D + warp) * L + l;
#pragma unroll 8
    for (int i = 0; i < D / 8; ++i) out[(size_t)i * 8 * L] = Tile::template get<D>(tile, lane, warp + 8 * i);
  }
}

// ---- rows that already are rows: x[B, L, D] with D contiguous (a channels-last feature map, the natural output of a
// 1x1 region head run as a GEMM) -> unit rows xn[B, Lpad, D] + norms.  No transpose: one warp per row, 16-byte accesses,
// the row stays in registers between the reduction and the store.  SURVEY section 8(f) N2: the producer emits the
// tensor-core kernels' layout directly, so the smem-transpose kernels above drop out of the word loss.
template <typename TI, typename TO, int kMaxV>
__global__ void __launch_bounds__(256) norm_rows_kernel(const TI* __restrict__ x, long long nrows, int D, int L, int Lpad,
                                                         TO* __restrict__ xn, float* __restrict__ norm) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);       // over B * Lpad
  if (row >= nrows) return;
  const int b = (int)(row / Lpad), l = (int)(row % Lpad);
  TO* dst = xn + (size_t)row * D;
  if (l >= L) {                                                              // padding rows: zeros, norm 0
    for (int d = lane * 4; d < D; d += 128) st4(dst + d, make_float4(0.f, 0.f, 0.f, 0.f));
    if (lane == 0) norm[row] = 0.f;
    return;
  }
  const TI* src = x + ((size_t)b * L + l) * D;
  float4 v[kMaxV];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxV; ++k) {
    const int d = k * 128 + lane * 4;
    v[k] = d < D ? ld4_nc(src + d) : make_float4(0.f, 0.f, 0.f, 0.f);
    ss += dot4(v[k], v[k]);
  }
  ss = warp_sum(ss);
  const float n = fmaxf(sqrtf(ss), kEps), inv = 1.f / n;
#pragma unroll
  for (int k = 0; k < kMaxV; ++k) {
    const int d = k * 128 + lane * 4;
    if (d < D) st4(dst + d, make_float4(v[k].x * inv, v[k].y * inv, v[k].z * inv, v[k].w * inv));
  }
  if (lane == 0) norm[row] = n;
}

// dx[b,l,:] = (dxn - xh <xh, dxn>) / norm + dnorm * xh    (rows l < L; same arithmetic as norm_tr_bwd_kernel)
template <typename TX, typename TO, int kMaxV>
__global__ void __launch_bounds__(256) norm_rows_bwd_kernel(const TX* __restrict__ xn, const float* __restrict__ norm,
                                                             const float* __restrict__ dxn, const float* __restrict__ dnorm,
                                                             long long nrows, int D, int L, int Lpad,
                                                             const int* __restrict__ error_word, TO* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const long long orow = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);      // over B * L
  if (orow >= nrows) return;
  const int b = (int)(orow / L), l = (int)(orow % L);
  const float poison = (error_word && __ldg(error_word) != 0) ? __int_as_float(0x7fc00000) : 0.f;
  const size_t irow = (size_t)b * Lpad + l;
  const float n = norm[irow];
  const bool clamped = n <= kEps;
  const float inv = 1.f / n;
  const float dn = (dnorm && !clamped) ? dnorm[irow] : 0.f;
  float4 xh[kMaxV], g[kMaxV];
  float proj = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxV; ++k) {
    const int d = k * 128 + lane * 4;
    if (d < D) {
      xh[k] = ld4_nc(xn + irow * D + d);
      g[k] = __ldg(reinterpret_cast<const float4*>(dxn + irow * D + d));
      proj += dot4(xh[k], g[k]);
    }
  }
  proj = clamped ? 0.f : warp_sum(proj);
  TO* dst = dx + (size_t)orow * D;
#pragma unroll
  for (int k = 0; k < kMaxV; ++k) {
    const int d = k * 128 + lane * 4;
    if (d < D)
      st4(dst + d, make_float4((g[k].x - xh[k].x * proj) * inv + dn * xh[k].x + poison, (g[k].y - xh[k].y * proj) * inv + dn * xh[k].y + poison,
                               (g[k].z - xh[k].z * proj) * inv + dn * xh[k].z + poison, (g[k].w - xh[k].w * proj) * inv + dn * xh[k].w + poison));
  }
}

// scores[i,c] = (1/rho2) log sum_{t unmasked} exp(rho2 rel[i, row(c,t)]);   one thread per (i,c).
// Dense rows: row(c,t) = c*T+t with the padding mask; compact rows: caption c owns rows [cap_ptr[c], cap_ptr[c+1]).
__global__ void __launch_bounds__(256) word_scores_kernel(const float* __restrict__ rel, const uint8_t* __restrict__ mask,
                                                           const int* __restrict__ cap_ptr, int Bi, int Bc, int T, int NQs,
                                                           float rho2, float* __restrict__ scores) {
  const size_t k = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (k >= (size_t)Bi * Bc) return;
  const int c = (int)(k % Bc);
  const size_t i = k / Bc;
  const int lo = cap_ptr ? cap_ptr[c] : c * T, hi = cap_ptr ? cap_ptr[c + 1] : c * T + T;
  const float* r = rel + i * NQs;
  const uint8_t* mk = (mask && !cap_ptr) ? mask : nullptr;
  float m = -INFINITY;
  for (int q0 = lo; q0 < hi; q0 += 8) {            // eight loads in flight
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (q0 + u < hi && (!mk || !mk[q0 + u])) ? rho2 * __ldg(r + q0 + u) : -INFINITY;
#pragma unroll
    for (int u = 0; u < 8; ++u) m = fmaxf(m, v[u]);
  }
  if (m == -INFINITY) { scores[k] = 0.f; return; }  // fully padded caption
  float s = 0.f;
  for (int q = lo; q < hi; ++q)
    if (!mk || !mk[q]) s += __expf(rho2 * __ldg(r + q) - m);
  scores[k] = (m + logf(s)) / rho2;
}

// grel[i, row(c,t)] = dscores[i,c] * softmax_t(rho2 rel)[t]  (0 for padding); one thread per (i,c)
__global__ void __launch_bounds__(256) word_scores_bwd_kernel(const float* __restrict__ rel, const uint8_t* __restrict__ mask,
                                                               const int* __restrict__ cap_ptr,
                                                               const float* __restrict__ scores, const float* __restrict__ dscores,
                                                               int Bi, int Bc, int T, int NQs, float rho2, float* __restrict__ grel) {
  const size_t k = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (k >= (size_t)Bi * Bc) return;
  const int c = (int)(k % Bc);
  const size_t i = k / Bc;
  const int lo = cap_ptr ? cap_ptr[c] : c * T, hi = cap_ptr ? cap_ptr[c + 1] : c * T + T;
  const uint8_t* mk = (mask && !cap_ptr) ? mask : nullptr;
  const float sc = scores[k], ds = dscores[k];
  for (int q0 = lo; q0 < hi; q0 += 8) {            // eight loads in flight
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (q0 + u < hi) ? __ldg(rel + i * NQs + q0 + u) : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (q0 + u >= hi) break;
      const bool pad = mk && mk[q0 + u];
      grel[i * NQs + q0 + u] = pad ? 0.f : ds * __expf(rho2 * (v[u] - sc));
    }
  }
}

#ifdef XMC_TEST_HOOKS
static int g_prep_generic = 0;   // tests / A-B timing only (libxmcloss_hooks.so): 1 = always take the generic kernels
#else
static constexpr int g_prep_generic = 0;
#endif

template <typename TI>
static int launch_norm_tr(const void* x, int B, int D, int L, int Lpad, int out_dtype, const int* row_of, void* xn, float* norm, cudaStream_t st) {
  dim3 grid((Lpad + kLT - 1) / kLT, B);
  if (std::is_same<TI, __nv_bfloat16>::value && out_dtype == XMC_BF16 && (D == 128 || D == 256) && !g_prep_generic) {
    auto* xi = static_cast<const __nv_bfloat16*>(x);
    auto* xo = static_cast<__nv_bfloat16*>(xn);
    if (D == 128) norm_tr_bf16_kernel<128><<<grid, 256, 0, st>>>(xi, L, Lpad, row_of, xo, norm);
    else norm_tr_bf16_kernel<256><<<grid, 256, 0, st>>>(xi, L, Lpad, row_of, xo, norm);
    return cuda_fail(cudaGetLastError(), "norm_tr_bf16_kernel launch");
  }
  size_t smem = (size_t)D * (kLT + 1) * sizeof(float);
  if (smem > 48 * 1024) {   // D > 360: opt in to the large dynamic shared-memory carve-out
    XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(norm_tr_kernel<TI, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(norm_tr_kernel<TI, __nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (out_dtype == XMC_F32)
    norm_tr_kernel<TI, float><<<grid, 256, smem, st>>>(static_cast<const TI*>(x), D, L, Lpad, row_of, static_cast<float*>(xn), norm);
  else
    norm_tr_kernel<TI, __nv_bfloat16><<<grid, 256, smem, st>>>(static_cast<const TI*>(x), D, L, Lpad, row_of, static_cast<__nv_bfloat16*>(xn), norm);
  return cuda_fail(cudaGetLastError(), "norm_tr_kernel launch");
}

template <typename TX>
static int launch_norm_tr_bwd(const void* xn, const float* norm, const float* dxn, const float* dnorm, int B, int D, int L,
                              int Lpad, int out_dtype, const int* row_of, const int* err, void* dx, cudaStream_t st) {
  dim3 grid((L + kLT - 1) / kLT, B);
  if (std::is_same<TX, __nv_bfloat16>::value && (D == 128 || D == 256) && !g_prep_generic) {
    auto* xi = static_cast<const __nv_bfloat16*>(xn);
    if (out_dtype == XMC_F32) {
      auto* o = static_cast<float*>(dx);
      if (D == 128) norm_tr_bwd_bf16_kernel<128, float><<<grid, 256, 0, st>>>(xi, norm, dxn, dnorm, L, Lpad, row_of, err, o);
      else norm_tr_bwd_bf16_kernel<256, float><<<grid, 256, 0, st>>>(xi, norm, dxn, dnorm, L, Lpad, row_of, err, o);
    } else {
      auto* o = static_cast<__nv_bfloat16*>(dx);
      if (D == 128) norm_tr_bwd_bf16_kernel<128, __nv_bfloat16><<<grid, 256, 0, st>>>(xi, norm, dxn, dnorm, L, Lpad, row_of, err, o);
      else norm_tr_bwd_bf16_kernel<256, __nv_bfloat16><<<grid, 256, 0, st>>>(xi, norm, dxn, dnorm, L, Lpad, row_of, err, o);
    }
    return cuda_fail(cudaGetLastError(), "norm_tr_bwd_bf16_kernel launch");
  }
  size_t smem = (size_t)D * (kLT + 1) * sizeof(float);
  if (smem > 48 * 1024) {
    XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(norm_tr_bwd_kernel<TX, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(norm_tr_bwd_kernel<TX, __nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (out_dtype == XMC_F32)
    norm_tr_bwd_kernel<TX, float><<<grid, 256, smem, st>>>(static_cast<const TX*>(xn), norm, dxn, dnorm, D, L, Lpad, row_of, err, static_cast<float*>(dx));
  else
    norm_tr_bwd_kernel<TX, __nv_bfloat16><<<grid, 256, smem, st>>>(static_cast<const TX*>(xn), norm, dxn, dnorm, D, L, Lpad, row_of, err, static_cast<__nv_bfloat16*>(dx));
  return cuda_fail(cudaGetLastError(), "norm_tr_bwd_kernel launch");
}

static int check_nt(const void* a, const void* b, int B, int D, int L, int Lpad, int t0, int t1) {
  XMC_REQUIRE(a && b, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(B > 0 && D > 0 && L > 0 && Lpad >= L, XMC_ERR_INVALID_ARG, "bad shape B=%d D=%d L=%d Lpad=%d", B, D, L, Lpad);
  XMC_REQUIRE(D <= 1024, XMC_ERR_UNSUPPORTED, "D=%d > 1024", D);
  XMC_REQUIRE((t0 == XMC_F32 || t0 == XMC_BF16) && (t1 == XMC_F32 || t1 == XMC_BF16), XMC_ERR_UNSUPPORTED, "dtype");
  return XMC_OK;
}

}  // namespace xmc

using namespace xmc;

#ifdef XMC_TEST_HOOKS
extern "C" void xmc_internal_set_prep_generic(int on) { xmc::g_prep_generic = on; }
#endif

extern "C" int xmc_word_rows_compact(const uint8_t* mask, int Bc, int T, int* row_of, int* cap_ptr, void* stream) {
  XMC_REQUIRE(mask && row_of && cap_ptr, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(Bc > 0 && T > 0, XMC_ERR_INVALID_ARG, "bad shape Bc=%d T=%d", Bc, T);
  word_rows_compact_kernel<<<1, 1024, 0, as_stream(stream)>>>(mask, Bc * T, T, row_of, cap_ptr);
  return cuda_fail(cudaGetLastError(), "word_rows_compact_kernel launch");
}

extern "C" int xmc_normalize_transpose(const void* x, int B, int D, int L, int Lpad, int in_dtype,
                                       int out_dtype, const int* row_of, void* xn, float* norm, void* stream) {
  if (int rc = check_nt(x, xn, B, D, L, Lpad, in_dtype, out_dtype)) return rc;
  XMC_REQUIRE(norm, XMC_ERR_INVALID_ARG, "null norm");
  XMC_REQUIRE(!row_of || Lpad == L, XMC_ERR_INVALID_ARG, "row_of needs Lpad == L");
  return in_dtype == XMC_F32 ? launch_norm_tr<float>(x, B, D, L, Lpad, out_dtype, row_of, xn, norm, as_stream(stream))
                             : launch_norm_tr<__nv_bfloat16>(x, B, D, L, Lpad, out_dtype, row_of, xn, norm, as_stream(stream));
}

extern "C" int xmc_normalize_transpose_backward(const void* xn, const float* norm, const float* dxn,
                                                const float* dnorm, int B, int D, int L, int Lpad,
                                                int xn_dtype, int out_dtype, const int* row_of,
                                                const int* error_word, void* dx, void* stream) {
  if (int rc = check_nt(xn, dx, B, D, L, Lpad, xn_dtype, out_dtype)) return rc;
  XMC_REQUIRE(norm && dxn, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(!row_of || (Lpad == L && !dnorm), XMC_ERR_INVALID_ARG, "row_of needs Lpad == L and no dnorm");
  return xn_dtype == XMC_F32
             ? launch_norm_tr_bwd<float>(xn, norm, dxn, dnorm, B, D, L, Lpad, out_dtype, row_of, error_word, dx, as_stream(stream))
             : launch_norm_tr_bwd<__nv_bfloat16>(xn, norm, dxn, dnorm, B, D, L, Lpad, out_dtype, row_of, error_word, dx, as_stream(stream));
}

template <typename TI, typename TO>
static int launch_norm_rows(const void* x, int B, int D, int L, int Lpad, void* xn, float* norm, cudaStream_t st) {
  const long long rows = (long long)B * Lpad;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  auto* xi = static_cast<const TI*>(x);
  auto* xo = static_cast<TO*>(xn);
  if (D <= 256) norm_rows_kernel<TI, TO, 2><<<grid, 256, 0, st>>>(xi, rows, D, L, Lpad, xo, norm);
  else norm_rows_kernel<TI, TO, 8><<<grid, 256, 0, st>>>(xi, rows, D, L, Lpad, xo, norm);
  return cuda_fail(cudaGetLastError(), "norm_rows_kernel launch");
}

template <typename TX, typename TO>
static int launch_norm_rows_bwd(const void* xn, const float* norm, const float* dxn, const float* dnorm, int B, int D, int L,
                                int Lpad, const int* err, void* dx, cudaStream_t st) {
  const long long rows = (long long)B * L;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  auto* xi = static_cast<const TX*>(xn);
  auto* o = static_cast<TO*>(dx);
  if (D <= 256) norm_rows_bwd_kernel<TX, TO, 2><<<grid, 256, 0, st>>>(xi, norm, dxn, dnorm, rows, D, L, Lpad, err, o);
  else norm_rows_bwd_kernel<TX, TO, 8><<<grid, 256, 0, st>>>(xi, norm, dxn, dnorm, rows, D, L, Lpad, err, o);
  return cuda_fail(cudaGetLastError(), "norm_rows_bwd_kernel launch");
}

extern "C" int xmc_normalize_rows(const void* x, int B, int D, int L, int Lpad, int in_dtype, int out_dtype, void* xn,
                                  float* norm, void* stream) {
  if (int rc = check_nt(x, xn, B, D, L, Lpad, in_dtype, out_dtype)) return rc;
  XMC_REQUIRE(norm, XMC_ERR_INVALID_ARG, "null norm");
  XMC_REQUIRE(D % 4 == 0 && aligned16(x) && aligned16(xn), XMC_ERR_ALIGNMENT, "rows need D %% 4 == 0 and 16-byte aligned pointers");
  cudaStream_t st = as_stream(stream);
  if (in_dtype == XMC_F32)
    return out_dtype == XMC_F32 ? launch_norm_rows<float, float>(x, B, D, L, Lpad, xn, norm, st)
                                : launch_norm_rows<float, __nv_bfloat16>(x, B, D, L, Lpad, xn, norm, st);
  return out_dtype == XMC_F32 ? launch_norm_rows<__nv_bfloat16, float>(x, B, D, L, Lpad, xn, norm, st)
                              : launch_norm_rows<__nv_bfloat16, __nv_bfloat16>(x, B, D, L, Lpad, xn, norm, st);
}

extern "C" int xmc_normalize_rows_backward(const void* xn, const float* norm, const float* dxn, const float* dnorm, int B, int D,
                                           int L, int Lpad, int xn_dtype, int out_dtype, const int* error_word, void* dx,
                                           void* stream) {
  if (int rc = check_nt(xn, dx, B, D, L, Lpad, xn_dtype, out_dtype)) return rc;
  XMC_REQUIRE(norm && dxn, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(D % 4 == 0 && aligned16(xn) && aligned16(dx) && aligned16(dxn), XMC_ERR_ALIGNMENT,
              "rows need D %% 4 == 0 and 16-byte aligned pointers");
  cudaStream_t st = as_stream(stream);
  if (xn_dtype == XMC_F32)
    return out_dtype == XMC_F32 ? launch_norm_rows_bwd<float, float>(xn, norm, dxn, dnorm, B, D, L, Lpad, error_word, dx, st)
                                : launch_norm_rows_bwd<float, __nv_bfloat16>(xn, norm, dxn, dnorm, B, D, L, Lpad, error_word, dx, st);
  return out_dtype == XMC_F32 ? launch_norm_rows_bwd<__nv_bfloat16, float>(xn, norm, dxn, dnorm, B, D, L, Lpad, error_word, dx, st)
                              : launch_norm_rows_bwd<__nv_bfloat16, __nv_bfloat16>(xn, norm, dxn, dnorm, B, D, L, Lpad, error_word, dx, st);
}

extern "C" int xmc_word_scores(const float* rel, const uint8_t* mask, const int* cap_ptr, int Bi, int Bc, int T,
                               int NQs, float rho2, float* scores, void* stream) {
  XMC_REQUIRE(rel && scores, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(Bi > 0 && Bc > 0 && T > 0 && NQs >= Bc && rho2 > 0.f, XMC_ERR_INVALID_ARG, "bad shape / rho2");
  size_t n = (size_t)Bi * Bc;
  word_scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(rel, mask, cap_ptr, Bi, Bc, T, NQs, rho2, scores);
  return cuda_fail(cudaGetLastError(), "word_scores_kernel launch");
}

extern "C" int xmc_word_scores_backward(const float* rel, const uint8_t* mask, const int* cap_ptr, const float* scores,
                                        const float* dscores, int Bi, int Bc, int T, int NQs, float rho2,
                                        float* grel, void* stream) {
  XMC_REQUIRE(rel && scores && dscores && grel, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(Bi > 0 && Bc > 0 && T > 0 && NQs >= Bc && rho2 > 0.f, XMC_ERR_INVALID_ARG, "bad shape / rho2");
  size_t n = (size_t)Bi * Bc;
  word_scores_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(rel, mask, cap_ptr, scores, dscores, Bi, Bc, T, NQs, rho2, grel);
  return cuda_fail(cudaGetLastError(), "word_scores_bwd_kernel launch");
}
