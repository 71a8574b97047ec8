// wordregion_f32.cu — word–region attention statistics, fp32 CUDA-core path (sm_100a).
//
// The precise path of `word_loss` (the loss the reference names at xmc_gan/train_gan.py:220-222,
// 267-269 but never implements; spec in oracle/word_region.py).  fp32 operands, fp32 FMA, meets
// the 1e-4 tolerance that TF32/bf16 tensor-core inputs cannot.  The bf16 tcgen05 path lives in
// wordregion_tc.cu; both implement the same C ABI and the same tiling idea:
//
//   tile = (image i) x (64 consecutive word rows of the flattened [Bc*T, D] word matrix)
//   for each chunk of 64 regions:  S = Q Khat^T  ->  P = exp(rho1 (S-1)) (* ||v_r||)
//                                  C += P Khat   (flash-attention style, constant shift rho1:
//                                                 cosines are bounded so no running max)
//   epilogue: l = sum P, a = sum P' S, ||C||  ->  lsum, cnorm, rel  per (image, word row)
//
// so the [Bi, Bc, T, R] score tensor only ever exists as 64x64 register tiles.  Backward
// recomputes S / P / C per tile, forms dS in registers and accumulates dQ in registers across
// the images a CTA visits and dKhat per chunk (fp32 atomics to the per-image gradient).
#include "common.cuh"
#include "wordregion.h"

namespace xmc {

constexpr int BM = 64;        // word rows per tile
constexpr int RC = 64;        // regions per chunk
constexpr int KB = 32;        // contraction block
constexpr int LDA = BM + 4;   // padded leading dim of transposed panels (float4-aligned)
constexpr float kLog2e = 1.4426950408889634f;

// Load a [64 rows x 32 k] block of a row-major matrix (leading dim ld) into smem transposed,
// T[k][row] with leading dim LDA.  Each warp covers 8 k x 4 rows per pass: bank = 4k+row = lane.
__device__ __forceinline__ void load_block_T(float* T, const float* __restrict__ M, int ld, int row0,
                                             int rows_valid, int k0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int pass = 0; pass < 8; ++pass) {
    const int wt = pass * 8 + warp;
    const int k = (wt & 3) * 8 + (lane >> 2);
    const int row = (wt >> 2) * 4 + (lane & 3);
    const int gr = row0 + row;
    T[k * LDA + row] = (gr < rows_valid) ? __ldg(M + (size_t)gr * ld + k0 + k) : 0.f;
  }
}

// Copy 32 rows x D of a row-major matrix into smem [32][D] (float4, coalesced).
template <int D>
__device__ __forceinline__ void load_rows(float* dst, const float* __restrict__ M, int row0, int rows_valid) {
  constexpr int V = D / 4;
#pragma unroll
  for (int e = threadIdx.x; e < 32 * V; e += 256) {
    const int r = e / V, c = e % V;
    const int gr = row0 + r;
    float4 v = (gr < rows_valid) ? *reinterpret_cast<const float4*>(M + (size_t)gr * D + c * 4) : make_float4(0, 0, 0, 0);
    *reinterpret_cast<float4*>(dst + r * D + c * 4) = v;
  }
}

// acc[i][j] += sum_kk A[kk][ty*4+i] * B[kk][tx*4+j]   over one 32-deep block
__device__ __forceinline__ void mma_4x4(float (&acc)[4][4], const float* A, const float* B, int ty, int tx) {
#pragma unroll 8
  for (int kk = 0; kk < KB; ++kk) {
    const float4 a = *reinterpret_cast<const float4*>(A + kk * LDA + ty * 4);
    const float4 b = *reinterpret_cast<const float4*>(B + kk * LDA + tx * 4);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// acc[i][jj] += sum_kk A[(kb+kk)*LDA + ty*4+i] * Bp[kk*ldb + jj*64 + tx*4 ..]   (32-deep)
template <int D>
__device__ __forceinline__ void mma_4xD(float4 (&acc)[4][D / 64], const float* A, const float* Bp, int ldb, int ty, int tx) {
#pragma unroll 4
  for (int kk = 0; kk < KB; ++kk) {
    const float4 a = *reinterpret_cast<const float4*>(A + kk * LDA + ty * 4);
    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int jj = 0; jj < D / 64; ++jj) {
      const float4 b = *reinterpret_cast<const float4*>(Bp + kk * ldb + jj * 64 + tx * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) fma4(acc[i][jj], av[i], b);
    }
  }
}

// One pass over all region chunks of image `K`: unnormalised context C, and (optionally) the
// softmax denominator and the numerator of <q, c>.
template <int D, bool kStats>
__device__ __forceinline__ void context_pass(const WrParams& p, const float* __restrict__ K, const float* __restrict__ rn,
                                             int m0, float* As, float* Bs1, float* Ps, float* Bs2,
                                             float4 (&C)[4][D / 64], float (&lpart)[4], float (&apart)[4]) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const float c1 = p.rho1 * kLog2e;
  for (int rc0 = 0; rc0 < p.R; rc0 += RC) {
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
    for (int k0 = 0; k0 < D; k0 += KB) {
      __syncthreads();
      load_block_T(As, static_cast<const float*>(p.qn), D, m0, p.NQ, k0);
      load_block_T(Bs1, K, D, rc0, p.Rpad, k0);
      __syncthreads();
      mma_4x4(s, As, Bs1, ty, tx);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = rc0 + tx * 4 + j;
      const bool valid = r < p.R;
      const float mr = rn ? (valid ? __ldg(rn + r) : 0.f) : 1.f;
      float pw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float pv = valid ? exp2f(c1 * (s[i][j] - 1.f)) : 0.f;
        pw[i] = pv * mr;
        if (kStats) { lpart[i] += pv; apart[i] = fmaf(pw[i], s[i][j], apart[i]); }
      }
      *reinterpret_cast<float4*>(Ps + (tx * 4 + j) * LDA + ty * 4) = make_float4(pw[0], pw[1], pw[2], pw[3]);
    }
    for (int kb = 0; kb < RC; kb += KB) {
      __syncthreads();
      load_rows<D>(Bs2, K, rc0 + kb, p.Rpad);
      __syncthreads();
      mma_4xD<D>(C, Ps + kb * LDA, Bs2, D, ty, tx);
    }
  }
}

// Compacted word rows: the number of valid rows lives on the device (nq_dev); NQ stays the row stride of the
// [Bi, NQ] statistics.  Tiles that start past the count have nothing to do.
__device__ __forceinline__ int valid_rows(const WrParams& p) { return p.nq_dev ? min(p.NQ, __ldg(p.nq_dev)) : p.NQ; }

template <int D>
__global__ void __launch_bounds__(256) wr_fwd_f32_kernel(WrParams p) {
  extern __shared__ __align__(16) float smem[];
  const int NQs = p.NQ;                 // row stride of lsum / cnorm / rel
  p.NQ = valid_rows(p);                 // rows at or past the count are neither read nor written
  if ((int)blockIdx.x * BM >= p.NQ) return;
  float* As = smem;
  float* Bs1 = As + KB * LDA;
  float* Ps = Bs1 + KB * LDA;
  float* Bs2 = Ps + RC * LDA;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int m0 = blockIdx.x * BM, img = blockIdx.y;
  const float* K = static_cast<const float*>(p.kn) + (size_t)img * p.Rpad * D;
  const float* rn = p.rnorm ? p.rnorm + (size_t)img * p.Rpad : nullptr;

  float4 C[4][D / 64];
  float lpart[4], apart[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lpart[i] = 0.f; apart[i] = 0.f;
#pragma unroll
    for (int jj = 0; jj < D / 64; ++jj) C[i][jj] = make_float4(0, 0, 0, 0);
  }
  context_pass<D, true>(p, K, rn, m0, As, Bs1, Ps, Bs2, C, lpart, apart);

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float c2 = 0.f;
#pragma unroll
    for (int jj = 0; jj < D / 64; ++jj) c2 += dot4(C[i][jj], C[i][jj]);
    c2 = half_warp_sum(c2);
    const float l = half_warp_sum(lpart[i]);
    const float a = half_warp_sum(apart[i]);
    const int row = m0 + ty * 4 + i;
    if (tx == 0 && row < p.NQ) {
      const float cn = sqrtf(c2) / l;
      const size_t o = (size_t)img * NQs + row;
      p.lsum[o] = l;
      p.cnorm[o] = cn;
      p.rel[o] = (a / l) / fmaxf(cn, kEps);
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256, 1) wr_bwd_f32_kernel(WrParams p) {
  constexpr int LDC = D + 4;
  extern __shared__ __align__(16) float smem[];
  const int NQs = p.NQ;
  p.NQ = valid_rows(p);
  if ((int)blockIdx.x * BM >= p.NQ) return;
  float* As = smem;
  float* Bs1 = As + KB * LDA;
  float* Xs = Bs1 + KB * LDA;      // [r][t]  (k = r for dQ = X Khat)        also P in pass 1
  float* XTs = Xs + RC * LDA;      // [t][r]  (k = t for dKhat = X^T Q - Y^T Chat)
  float* YTs = XTs + BM * LDA;
  float* Bs2 = YTs + BM * LDA;     // [32][D]
  float* Cst = Bs2 + KB * D;       // [64 t][LDC]  unit context rows
  float* colacc = Cst + BM * LDC;  // [64] d rnorm partials of the current chunk
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int m0 = blockIdx.x * BM;
  const float c1 = p.rho1 * kLog2e;

  float4 dQ[4][D / 64];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jj = 0; jj < D / 64; ++jj) dQ[i][jj] = make_float4(0, 0, 0, 0);
  if (threadIdx.x < RC) colacc[threadIdx.x] = 0.f;

  for (int img = blockIdx.y; img < p.Bi; img += gridDim.y) {
    const float* K = static_cast<const float*>(p.kn) + (size_t)img * p.Rpad * D;
    const float* rn = p.rnorm ? p.rnorm + (size_t)img * p.Rpad : nullptr;

    // per-row saved statistics of my 4 word rows
    float inv_l[4], gam[4], relv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = m0 + ty * 4 + i;
      float l = 1.f, b = 1.f, rl = 0.f, g = 0.f;
      if (row < p.NQ) {
        const size_t o = (size_t)img * NQs + row;
        l = p.lsum[o]; b = fmaxf(p.cnorm[o], kEps); rl = p.rel[o]; g = p.grel[o];
      }
      inv_l[i] = 1.f / l; gam[i] = g / b; relv[i] = rl;
    }
    // ---- pass 1: context rows -> Cst --------------------------------------------------------
    {
      float4 C[4][D / 64];
      float dummy0[4], dummy1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < D / 64; ++jj) C[i][jj] = make_float4(0, 0, 0, 0);
      context_pass<D, false>(p, K, rn, m0, As, Bs1, Xs, Bs2, C, dummy0, dummy1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = m0 + ty * 4 + i;
        float cs = 0.f;
        if (row < p.NQ) {
          const size_t o = (size_t)img * NQs + row;
          cs = inv_l[i] / fmaxf(p.cnorm[o], kEps);
        }
#pragma unroll
        for (int jj = 0; jj < D / 64; ++jj) {
          float4 v = C[i][jj];
          v.x *= cs; v.y *= cs; v.z *= cs; v.w *= cs;
          *reinterpret_cast<float4*>(Cst + (ty * 4 + i) * LDC + jj * 64 + tx * 4) = v;
        }
      }
    }
    __syncthreads();

    // ---- pass 2: per chunk dS, dQ, dKhat ----------------------------------------------------
    for (int rc0 = 0; rc0 < p.R; rc0 += RC) {
      float s[4][4], w[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[i][j] = 0.f; w[i][j] = 0.f; }
      for (int k0 = 0; k0 < D; k0 += KB) {
        __syncthreads();
        load_block_T(As, static_cast<const float*>(p.qn), D, m0, p.NQ, k0);
        load_block_T(Bs1, K, D, rc0, p.Rpad, k0);
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < KB; ++kk) {
          const float4 a = *reinterpret_cast<const float4*>(As + kk * LDA + ty * 4);
          const float4 b = *reinterpret_cast<const float4*>(Bs1 + kk * LDA + tx * 4);
          const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float ch = Cst[(ty * 4 + i) * LDC + k0 + kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              s[i][j] = fmaf(av[i], bv[j], s[i][j]);
              w[i][j] = fmaf(ch, bv[j], w[i][j]);
            }
          }
        }
      }
      // elementwise: X = dS + gamma*alpha', Y = -gamma*rel*alpha'
      float xr[4][4], yr[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = rc0 + tx * 4 + j;
        const bool valid = r < p.R;
        const float mr = rn ? (valid ? __ldg(rn + r) : 0.f) : 1.f;
        float dm = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float al = valid ? exp2f(c1 * (s[i][j] - 1.f)) * inv_l[i] : 0.f;   // alpha
          const float alp = al * mr;                                               // alpha'
          const float dap = gam[i] * (s[i][j] - relv[i] * w[i][j]);                // d alpha'
          const float ds = p.rho1 * alp * dap;
          xr[i][j] = ds + gam[i] * alp;
          yr[i][j] = -gam[i] * relv[i] * alp;
          dm = fmaf(al, dap, dm);
        }
        if (rn && valid) atomicAdd(colacc + tx * 4 + j, dm);
        *reinterpret_cast<float4*>(Xs + (tx * 4 + j) * LDA + ty * 4) = make_float4(xr[0][j], xr[1][j], xr[2][j], xr[3][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        *reinterpret_cast<float4*>(XTs + (ty * 4 + i) * LDA + tx * 4) = make_float4(xr[i][0], xr[i][1], xr[i][2], xr[i][3]);
        *reinterpret_cast<float4*>(YTs + (ty * 4 + i) * LDA + tx * 4) = make_float4(yr[i][0], yr[i][1], yr[i][2], yr[i][3]);
      }
      // dQ += X Khat_chunk
      for (int kb = 0; kb < RC; kb += KB) {
        __syncthreads();
        load_rows<D>(Bs2, K, rc0 + kb, p.Rpad);
        __syncthreads();
        mma_4xD<D>(dQ, Xs + kb * LDA, Bs2, D, ty, tx);
      }
      // dKhat_chunk[r][d] = sum_t X[t][r] Q[t][d] + Y[t][r] Chat[t][d]
      float4 dK[4][D / 64];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < D / 64; ++jj) dK[i][jj] = make_float4(0, 0, 0, 0);
      for (int kb = 0; kb < BM; kb += KB) {
        __syncthreads();
        load_rows<D>(Bs2, static_cast<const float*>(p.qn), m0 + kb, p.NQ);
        __syncthreads();
        mma_4xD<D>(dK, XTs + kb * LDA, Bs2, D, ty, tx);
      }
      mma_4xD<D>(dK, YTs, Cst, LDC, ty, tx);
      mma_4xD<D>(dK, YTs + KB * LDA, Cst + KB * LDC, LDC, ty, tx);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = rc0 + ty * 4 + i;
        if (r < p.R) {
          float* dst = p.dkn + ((size_t)img * p.Rpad + r) * D;
#pragma unroll
          for (int jj = 0; jj < D / 64; ++jj) {
            const int col = jj * 64 + tx * 4;
            atomicAdd(dst + col + 0, dK[i][jj].x); atomicAdd(dst + col + 1, dK[i][jj].y);
            atomicAdd(dst + col + 2, dK[i][jj].z); atomicAdd(dst + col + 3, dK[i][jj].w);
          }
        }
      }
      __syncthreads();
      if (rn && threadIdx.x < RC) {
        const int r = rc0 + threadIdx.x;
        if (r < p.R) atomicAdd(p.drnorm + (size_t)img * p.Rpad + r, colacc[threadIdx.x]);
        colacc[threadIdx.x] = 0.f;
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row < p.NQ) {
      float* dst = p.dqn + (size_t)row * D;
#pragma unroll
      for (int jj = 0; jj < D / 64; ++jj) {
        const int col = jj * 64 + tx * 4;
        atomicAdd(dst + col + 0, dQ[i][jj].x); atomicAdd(dst + col + 1, dQ[i][jj].y);
        atomicAdd(dst + col + 2, dQ[i][jj].z); atomicAdd(dst + col + 3, dQ[i][jj].w);
      }
    }
  }
}

template <int D>
static int launch_fwd(const WrParams& p, cudaStream_t st) {
  const size_t smem = (size_t)(2 * KB * LDA + RC * LDA + KB * D) * sizeof(float);
  XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(wr_fwd_f32_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((p.NQ + BM - 1) / BM, p.Bi);
  wr_fwd_f32_kernel<D><<<grid, 256, smem, st>>>(p);
  return cuda_fail(cudaGetLastError(), "wr_fwd_f32_kernel launch");
}

template <int D>
static int launch_bwd(const WrParams& p, cudaStream_t st) {
  const size_t smem = (size_t)(2 * KB * LDA + RC * LDA + 2 * BM * LDA + KB * D + BM * (D + 4) + RC) * sizeof(float);
  XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(wr_bwd_f32_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = (p.NQ + BM - 1) / BM;
  // enough image groups to fill 148 SMs a few times over; dQ stays in registers inside a group
  int groups = (148 * 4 + tiles - 1) / tiles;
  if (groups > p.Bi) groups = p.Bi;
  if (groups < 1) groups = 1;
  dim3 grid(tiles, groups);
  wr_bwd_f32_kernel<D><<<grid, 256, smem, st>>>(p);
  return cuda_fail(cudaGetLastError(), "wr_bwd_f32_kernel launch");
}

int wordregion_f32_forward(const WrParams& p, int D, cudaStream_t st) {
  switch (D) {
    case 64: return launch_fwd<64>(p, st);
    case 128: return launch_fwd<128>(p, st);
    case 256: return launch_fwd<256>(p, st);
  }
  set_error("word-region D=%d unsupported (64, 128, 256)", D);
  return XMC_ERR_UNSUPPORTED;
}

int wordregion_f32_backward(const WrParams& p, int D, cudaStream_t st) {
  switch (D) {
    case 64: return launch_bwd<64>(p, st);
    case 128: return launch_bwd<128>(p, st);
    case 256: return launch_bwd<256>(p, st);
  }
  set_error("word-region D=%d unsupported (64, 128, 256)", D);
  return XMC_ERR_UNSUPPORTED;
}

}  // namespace xmc
