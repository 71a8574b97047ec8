// prep.cu — layout/normalisation prologue and epilogue of the word-region loss, and the
// per-caption masked log-sum-exp that turns word relevances into the score matrix.
//
// The reference's encoders hand over words as [B, D, T] and the discriminator feature map as
// [B, D, H, W] (xmc_gan/model/encoder.py:68,140): the contraction dimension D is the SLOW one.
// The tensor-core kernels want unit rows with D contiguous, so one HBM-bound pass does
// F.normalize (train_gan.py:88-89 convention) + transpose + optional bf16 cast through a padded
// shared-memory tile (reads coalesced along L, writes coalesced along D).
#include "common.cuh"

namespace xmc {

constexpr int kLT = 32;  // (b, 32 consecutive l) per CTA

// x[B,D,L] -> xn[B,Lpad,D], norm[B,Lpad].  256 threads; dynamic smem D*(kLT+1) floats.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) norm_tr_kernel(const TI* __restrict__ x, int D, int L, int Lpad,
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
    TO* dst = xn + ((size_t)b * Lpad + lr) * D;
    for (int d = lx; d < D; d += 32) st1(dst + d, tile[d * (kLT + 1) + r] * inv);
  }
}

// dx[b,d,l] = (dxn[l,d] - xh[l,d] * <xh_l, dxn_l>) / norm_l + dnorm_l * xh[l,d]
template <typename TX, typename TO>
__global__ void __launch_bounds__(256) norm_tr_bwd_kernel(const TX* __restrict__ xn, const float* __restrict__ norm,
                                                           const float* __restrict__ dxn, const float* __restrict__ dnorm,
                                                           int D, int L, int Lpad, TO* __restrict__ dx) {
  extern __shared__ float tile[];                 // [D][kLT+1] holds the finished dx tile
  const int b = blockIdx.y, l0 = blockIdx.x * kLT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < kLT; r += 8) {
    const int l = l0 + r;
    if (l >= L) {
      for (int d = lane; d < D; d += 32) tile[d * (kLT + 1) + r] = 0.f;
      continue;
    }
    const size_t row = ((size_t)b * Lpad + l) * D;
    float proj = 0.f;
    for (int d = lane; d < D; d += 32) proj = fmaf(ld1(xn + row + d), dxn[row + d], proj);
    proj = warp_sum(proj);
    const float n = norm[(size_t)b * Lpad + l];
    const bool clamped = n <= kEps;               // x/eps is linear: no projection, no norm path
    const float inv = 1.f / n;
    const float dn = (dnorm && !clamped) ? dnorm[(size_t)b * Lpad + l] : 0.f;
    if (clamped) proj = 0.f;
    for (int d = lane; d < D; d += 32) {
      float xh = ld1(xn + row + d);
      tile[d * (kLT + 1) + r] = (dxn[row + d] - xh * proj) * inv + dn * xh;
    }
  }
  __syncthreads();
  const int l = l0 + lane;
  if (l < L)
    for (int d = warp; d < D; d += 8) st1(dx + ((size_t)b * D + d) * L + l, tile[d * (kLT + 1) + lane]);
}

// scores[i,c] = (1/rho2) log sum_{t unmasked} exp(rho2 rel[i, c*T+t]);   one thread per (i,c)
__global__ void __launch_bounds__(256) word_scores_kernel(const float* __restrict__ rel, const uint8_t* __restrict__ mask,
                                                           int Bi, int Bc, int T, float rho2, float* __restrict__ scores) {
  const size_t k = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (k >= (size_t)Bi * Bc) return;
  const int c = (int)(k % Bc);
  const float* r = rel + k * T;                     // (i*Bc + c)*T
  float m = -INFINITY;
  for (int t = 0; t < T; ++t)
    if (!mask || !mask[(size_t)c * T + t]) m = fmaxf(m, rho2 * r[t]);
  if (m == -INFINITY) { scores[k] = 0.f; return; }  // fully padded caption
  float s = 0.f;
  for (int t = 0; t < T; ++t)
    if (!mask || !mask[(size_t)c * T + t]) s += __expf(rho2 * r[t] - m);
  scores[k] = (m + logf(s)) / rho2;
}

// grel[i, c*T+t] = dscores[i,c] * softmax_t(rho2 rel)[t]  (0 for padding)
__global__ void __launch_bounds__(256) word_scores_bwd_kernel(const float* __restrict__ rel, const uint8_t* __restrict__ mask,
                                                               const float* __restrict__ scores, const float* __restrict__ dscores,
                                                               int Bi, int Bc, int T, float rho2, float* __restrict__ grel) {
  const size_t n = (size_t)Bi * Bc * T;
  for (size_t k = (size_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (size_t)gridDim.x * 256) {
    const size_t ic = k / T;
    const int t = (int)(k % T), c = (int)(ic % Bc);
    const bool pad = mask && mask[(size_t)c * T + t];
    grel[k] = pad ? 0.f : dscores[ic] * __expf(rho2 * (rel[k] - scores[ic]));
  }
}

template <typename TI>
static int launch_norm_tr(const void* x, int B, int D, int L, int Lpad, int out_dtype, void* xn, float* norm, cudaStream_t st) {
  dim3 grid((Lpad + kLT - 1) / kLT, B);
  size_t smem = (size_t)D * (kLT + 1) * sizeof(float);
  if (out_dtype == XMC_F32)
    norm_tr_kernel<TI, float><<<grid, 256, smem, st>>>(static_cast<const TI*>(x), D, L, Lpad, static_cast<float*>(xn), norm);
  else
    norm_tr_kernel<TI, __nv_bfloat16><<<grid, 256, smem, st>>>(static_cast<const TI*>(x), D, L, Lpad, static_cast<__nv_bfloat16*>(xn), norm);
  return cuda_fail(cudaGetLastError(), "norm_tr_kernel launch");
}

template <typename TX>
static int launch_norm_tr_bwd(const void* xn, const float* norm, const float* dxn, const float* dnorm, int B, int D, int L,
                              int Lpad, int out_dtype, void* dx, cudaStream_t st) {
  dim3 grid((L + kLT - 1) / kLT, B);
  size_t smem = (size_t)D * (kLT + 1) * sizeof(float);
  if (out_dtype == XMC_F32)
    norm_tr_bwd_kernel<TX, float><<<grid, 256, smem, st>>>(static_cast<const TX*>(xn), norm, dxn, dnorm, D, L, Lpad, static_cast<float*>(dx));
  else
    norm_tr_bwd_kernel<TX, __nv_bfloat16><<<grid, 256, smem, st>>>(static_cast<const TX*>(xn), norm, dxn, dnorm, D, L, Lpad, static_cast<__nv_bfloat16*>(dx));
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

extern "C" int xmc_normalize_transpose(const void* x, int B, int D, int L, int Lpad, int in_dtype,
                                       int out_dtype, void* xn, float* norm, void* stream) {
  if (int rc = check_nt(x, xn, B, D, L, Lpad, in_dtype, out_dtype)) return rc;
  XMC_REQUIRE(norm, XMC_ERR_INVALID_ARG, "null norm");
  return in_dtype == XMC_F32 ? launch_norm_tr<float>(x, B, D, L, Lpad, out_dtype, xn, norm, as_stream(stream))
                             : launch_norm_tr<__nv_bfloat16>(x, B, D, L, Lpad, out_dtype, xn, norm, as_stream(stream));
}

extern "C" int xmc_normalize_transpose_backward(const void* xn, const float* norm, const float* dxn,
                                                const float* dnorm, int B, int D, int L, int Lpad,
                                                int xn_dtype, int out_dtype, void* dx, void* stream) {
  if (int rc = check_nt(xn, dx, B, D, L, Lpad, xn_dtype, out_dtype)) return rc;
  XMC_REQUIRE(norm && dxn, XMC_ERR_INVALID_ARG, "null pointer");
  return xn_dtype == XMC_F32
             ? launch_norm_tr_bwd<float>(xn, norm, dxn, dnorm, B, D, L, Lpad, out_dtype, dx, as_stream(stream))
             : launch_norm_tr_bwd<__nv_bfloat16>(xn, norm, dxn, dnorm, B, D, L, Lpad, out_dtype, dx, as_stream(stream));
}

extern "C" int xmc_word_scores(const float* rel, const uint8_t* mask, int Bi, int Bc, int T, float rho2,
                               float* scores, void* stream) {
  XMC_REQUIRE(rel && scores, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(Bi > 0 && Bc > 0 && T > 0 && rho2 > 0.f, XMC_ERR_INVALID_ARG, "bad shape / rho2");
  size_t n = (size_t)Bi * Bc;
  word_scores_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(rel, mask, Bi, Bc, T, rho2, scores);
  return cuda_fail(cudaGetLastError(), "word_scores_kernel launch");
}

extern "C" int xmc_word_scores_backward(const float* rel, const uint8_t* mask, const float* scores,
                                        const float* dscores, int Bi, int Bc, int T, float rho2,
                                        float* grel, void* stream) {
  XMC_REQUIRE(rel && scores && dscores && grel, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(Bi > 0 && Bc > 0 && T > 0 && rho2 > 0.f, XMC_ERR_INVALID_ARG, "bad shape / rho2");
  size_t n = (size_t)Bi * Bc * T;
  size_t g = (n + 255) / 256; if (g > 148 * 16) g = 148 * 16;
  word_scores_bwd_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>(rel, mask, scores, dscores, Bi, Bc, T, rho2, grel);
  return cuda_fail(cudaGetLastError(), "word_scores_bwd_kernel launch");
}
