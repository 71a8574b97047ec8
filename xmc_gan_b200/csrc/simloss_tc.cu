// simloss_tc.cu — sentence–image / image–image InfoNCE on the 5th-gen tensor cores, for LARGE rectangular
// problems (global negatives: 256 local rows x 2048 gathered columns, D = 256 / 512).
//
// Replaces the same reference lines as simloss.cu (xmc_gan/train_gan.py:85-139); simloss.cu's CUDA-core kernels are
// built for the latency of a 256 x 256 problem and take 100-200 us at 256 x 2048 (they stream one side per resident
// row group and compute the score matrix twice).  Here the three products are tcgen05 GEMM tiles:
//
//   forward   S = (A B^T) * inv|a_i| * inv|b_j|             tile 128 x 128, K = D           -> scores, inverse norms
//   backward  dS(i,j) in closed form from S and the statistics (never stored), per 128 x 128 block:
//             G_a[i,:] += sum_j dS(i,j) bhat_j             A = dS  K-major,  B = bhat MN-major
//             G_b[j,:] += sum_i dS(i,j) ahat_i             A = dS  MN-major (the SAME smem bytes), B = ahat MN-major
//             then one pass of normalise-backward over the rows of G_a, G_b.
//
// Precision.  The reference computes in fp32 (torch.mm, :90) and north_star asks for rel 1e-4 with fp32 inputs, which a
// single bf16 (or tf32) product cannot give.  Every fp32 operand is therefore split on the fly into two bf16 numbers,
// x = hi + lo (|x - hi - lo| <= 2^-17 |x|), and each product runs as three MMAs, hi*hi + hi*lo + lo*hi, accumulated
// in fp32 in TMEM: relative error ~1e-5, an order inside the tolerance.  bf16 inputs are exact as `hi` in the forward
// (norms are applied to the fp32 accumulator afterwards), so that GEMM is one MMA per step.
//
// Operands are staged by the CTA's threads (global -> registers -> split -> 128B-swizzled smem tiles): the values
// need arithmetic on the way, so TMA cannot carry them.  The problems are tiny (0.3-0.5 GFLOP): one k-block in
// flight, MMAs issued by one thread and waited for by all — latency, not throughput, is what matters here.
#include <type_traits>

#include "simloss.cuh"
#include "tc_common.cuh"

namespace xmc {

using namespace tc;

namespace {

constexpr int kT = 128;                 // tile rows / columns (UMMA M and N)
constexpr int kBlk = kT * 128;          // bytes of one [128 rows x 64 bf16] swizzled block
constexpr int kTcThreads = 256;

// 16-byte chunk c (8 bf16) of row r in a [128 x 64] bf16 block, 128B swizzle: chunk position XOR (row mod 8)
__device__ __forceinline__ void st_chunk(uint8_t* blk, int r, int c, uint4 v) {
  *reinterpret_cast<uint4*>(blk + r * 128 + ((c ^ (r & 7)) << 4)) = v;
}

// x[0..8) -> bf16 hi (round to nearest) and bf16 lo = rn(x - hi)
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat162 hv = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
    const float2 hf = __bfloat1622float2(hv);
    h[e] = *reinterpret_cast<const uint32_t*>(&hv);
    l[e] = pack_bf16(x[2 * e] - hf.x, x[2 * e + 1] - hf.y);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void load8(const float* p, float (&x)[8]) {
  const float4 u = __ldg(reinterpret_cast<const float4*>(p)), v = __ldg(reinterpret_cast<const float4*>(p) + 1);
  x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w; x[4] = v.x; x[5] = v.y; x[6] = v.z; x[7] = v.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&x)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) { x[2 * e] = __uint_as_float(w[e] << 16); x[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u); }
}

struct TcShared {
  uint64_t bar;            // MMA completion
  uint32_t tmem_slot;
  int abort_flag;
};

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

// ---------------------------------------------------------------------------------------------------------
// forward: scores tile [128 x 128] = raw dot products (split bf16) scaled by the two inverse norms
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kTcThreads) sim_tc_fwd_kernel(SimParams p) {
  constexpr bool kSplit = std::is_same<T, float>::value;       // bf16 inputs are exact as `hi`
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* Ah = smem, *Al = smem + kBlk, *Bh = smem + 2 * kBlk, *Bl = smem + 3 * kBlk;
  float* ss = reinterpret_cast<float*>(smem + 4 * kBlk);       // [2 operands][2 halves][128] partial sums of squares
  float* inv_s = ss + 512;                                      // [2][128] inverse norms of the tile's rows / columns
  TcShared* sh = reinterpret_cast<TcShared*>(inv_s + 256);
  const WaitCtx wc{&sh->abort_flag, nullptr};

  const int tid = threadIdx.x, warp = warp_index(), lane = tid & 31;
  const int row = tid & 127, half = tid >> 7;
  const int i0 = blockIdx.y * kT, j0 = blockIdx.x * kT;
  const T* A = static_cast<const T*>(p.a);
  const T* B = static_cast<const T*>(p.b);

  if (tid == 0) { sh->abort_flag = 0; mbar_init(&sh->bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&sh->tmem_slot, kT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_slot;

  constexpr uint32_t idesc = idesc_bf16(kT, kT, false, false);
  float ssa = 0.f, ssb = 0.f;
  uint32_t phase = 0;
  for (int kb = 0; kb < p.D / 64; ++kb) {
    // stage the k-block: thread = (row, half of the 64 k), raw values -> hi (/ lo)
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const T* X = side ? B : A;
      const int g = (side ? j0 : i0) + row, n = side ? p.Bk : p.Bq;
      uint8_t* Xh = side ? Bh : Ah;
      uint8_t* Xl = side ? Bl : Al;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float x[8];
        if (g < n) load8(X + (size_t)g * p.D + kb * 64 + half * 32 + c * 8, x);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = fmaf(x[e], x[e], acc);
        uint4 hi, lo;
        split8(x, hi, lo);
        st_chunk(Xh, row, half * 4 + c, hi);
        if (kSplit) st_chunk(Xl, row, half * 4 + c, lo);
      }
      if (side) ssb += acc; else ssa += acc;
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const Desc ah = make_desc(smem_u32(Ah) + k * 32, 16, 1024), bh = make_desc(smem_u32(Bh) + k * 32, 16, 1024);
          mma_ss(tmem, ah, bh, idesc, kb > 0 || k > 0);
          if (kSplit) {
            const Desc al = make_desc(smem_u32(Al) + k * 32, 16, 1024), bl = make_desc(smem_u32(Bl) + k * 32, 16, 1024);
            mma_ss(tmem, ah, bl, idesc, true);
            mma_ss(tmem, al, bh, idesc, true);
          }
        }
        mma_commit(&sh->bar);
      }
      __syncwarp();
    }
    mbar_wait(&sh->bar, phase, wc, 1);        // the MMAs have read the tiles (last block: the accumulator is complete)
    phase ^= 1;
  }
  tc_fence_after();
  ss[0 * 256 + half * 128 + row] = ssa;
  ss[1 * 256 + half * 128 + row] = ssb;
  __syncthreads();
  if (tid < 256) {
    const int side = tid >> 7, r = tid & 127;
    const float n2 = ss[side * 256 + r] + ss[side * 256 + 128 + r];
    const float inv = 1.f / fmaxf(sqrtf(n2), kEps);
    inv_s[side * 128 + r] = inv;
    const int g = (side ? j0 : i0) + r;
    if (side == 0 && blockIdx.x == 0 && g < p.Bq && p.inv_a) p.inv_a[g] = inv;
    if (side == 1 && blockIdx.y == 0 && g < p.Bk && p.inv_b) p.inv_b[g] = inv;
  }
  __syncthreads();
  // epilogue: warp w reads TMEM lanes 32 (w & 3) .. +31 (= tile rows), columns 64 (w >> 2) .. +63
  {
    const int r = (warp & 3) * 32 + lane, i = i0 + r;
    const float ia = inv_s[r];
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c0 = (warp >> 2) * 64 + cc * 32;
      uint32_t v[32];
      tmem_ld32(lane_base + c0, v);
      tmem_wait_ld();
      if (i < p.Bq) {
        float* dst = p.scores + (size_t)i * p.Bk + j0 + c0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int j = j0 + c0 + 4 * q;
          float4 o;
          o.x = __uint_as_float(v[4 * q + 0]) * ia * inv_s[128 + c0 + 4 * q + 0];
          o.y = __uint_as_float(v[4 * q + 1]) * ia * inv_s[128 + c0 + 4 * q + 1];
          o.z = __uint_as_float(v[4 * q + 2]) * ia * inv_s[128 + c0 + 4 * q + 2];
          o.w = __uint_as_float(v[4 * q + 3]) * ia * inv_s[128 + c0 + 4 * q + 3];
          if (j + 3 < p.Bk) *reinterpret_cast<float4*>(dst + 4 * q) = o;          // Bk % 4 == 0 (eligibility): all or nothing
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kT);
}

// ---------------------------------------------------------------------------------------------------------
// backward: one CTA per (128 x 128 block of dS, 128 features); both products from one staged dS tile
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kTcThreads) sim_tc_bwd_kernel(SimParams p, float* __restrict__ Ga, float* __restrict__ Gb) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* Sh = smem;                       // dS hi: 2 blocks [128 i x 64 j]
  uint8_t* Sl = smem + 2 * kBlk;            // dS lo
  uint8_t* Xh = smem + 4 * kBlk;            // other operand (bhat or ahat rows) hi: 2 blocks [128 rows x 64 d]
  uint8_t* Xl = smem + 6 * kBlk;
  float* cst = reinterpret_cast<float*>(smem + 8 * kBlk);   // [3][128] column statistics of the block: lse, label sum, 1/n
  TcShared* sh = reinterpret_cast<TcShared*>(cst + 384);
  const WaitCtx wc{&sh->abort_flag, nullptr};

  const int tid = threadIdx.x, warp = warp_index(), lane = tid & 31;
  const int row = tid & 127, half = tid >> 7;
  const int i0 = blockIdx.y * kT, j0 = blockIdx.x * kT, d0 = blockIdx.z * kT;
  const T* A = static_cast<const T*>(p.a);
  const T* B = static_cast<const T*>(p.b);

  if (tid == 0) { sh->abort_flag = 0; mbar_init(&sh->bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&sh->tmem_slot, kT);
  if (tid < 128) {
    const int j = j0 + tid;
    const bool ok = j < p.Bk;
    cst[tid] = ok ? p.col_stats[j] : 0.f;
    cst[128 + tid] = ok ? p.col_stats[p.Bk + j] : 0.f;
    cst[256 + tid] = ok ? p.inv_cols_total / (p.col_div ? p.col_div[j] : p.num_pos) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_slot;

  // ---- dS block: thread = (row i, 64 columns), closed form of train_gan.py:103-113's autograd ----
  {
    const int i = i0 + row;
    const bool iok = i < p.Bq;
    const float go = __ldg(p.grad_out) * p.scale;
    const float r_lse = iok ? p.row_stats[i] : 0.f, r_sl = iok ? p.row_stats[p.Bq + i] : 0.f;
    const float r_inv = iok ? p.inv_rows_total / (p.row_div ? p.row_div[i] : p.num_pos) : 0.f;
#pragma unroll 2
    for (int c = 0; c < 8; ++c) {
      const int jl = half * 64 + c * 8, j = j0 + jl;
      float x[8];
      if (iok && j + 7 < p.Bk) {
        float s[8];
        load8(p.scores + (size_t)i * p.Bk + j, s);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          x[e] = go * dscore(p.scale * s[e], label_at(p.labels, p.Bk, i, j + e, p.diag), r_lse, r_sl, r_inv,
                             cst[jl + e], cst[128 + jl + e], cst[256 + jl + e]);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const bool ok = iok && (j + e) < p.Bk;
          x[e] = ok ? go * dscore(p.scale * __ldg(p.scores + (size_t)i * p.Bk + j + e), label_at(p.labels, p.Bk, i, j + e, p.diag),
                                  r_lse, r_sl, r_inv, cst[jl + e], cst[128 + jl + e], cst[256 + jl + e]) : 0.f;
        }
      }
      uint4 hi, lo;
      split8(x, hi, lo);
      st_chunk(Sh + half * kBlk, row, c, hi);
      st_chunk(Sl + half * kBlk, row, c, lo);
    }
  }

  uint32_t phase = 0;
  // which: 0 -> G_a (rows i, other operand bhat rows j), 1 -> G_b (rows j, other operand ahat rows i)
  for (int which = 0; which < 2; ++which) {
    float* G = which ? Gb : Ga;
    if (!G) continue;                                   // this gradient is not needed (uniform over the grid)
    const T* X = which ? A : B;
    const float* inv_x = which ? p.inv_a : p.inv_b;
    const int x0 = which ? i0 : j0, nx = which ? p.Bq : p.Bk;
    // stage xhat rows [128 x 128 features]: thread = (row, 64-feature block), unit rows = raw * 1/|x|
    {
      const int g = x0 + row;
      const float inv = g < nx ? inv_x[g] : 0.f;
#pragma unroll 2
      for (int c = 0; c < 8; ++c) {
        float x[8];
        if (g < nx) load8(X + (size_t)g * p.D + d0 + half * 64 + c * 8, x);
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = g < nx ? x[e] * inv : 0.f;
        uint4 hi, lo;
        split8(x, hi, lo);
        st_chunk(Xh + half * kBlk, row, c, hi);
        st_chunk(Xl + half * kBlk, row, c, lo);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        // contraction over the 128 rows of the xhat tile, 16 per MMA.  xhat is the B operand, MN-major: a k-step is
        // 16 rows = 2048 bytes, the second 64-feature block is kBlk further (LBO).  dS is the A operand:
        //   which = 0: K-major  (M = i rows, K = j: 4 k-steps per 64-column block, 32 bytes apart)
        //   which = 1: MN-major (M = j, 64 per block LBO apart; K = i rows, 16 rows = 2048 bytes per k-step)
        const uint32_t idesc = idesc_bf16(kT, kT, which == 1, true);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t xoff = k * 2048;
          const Desc xh = make_desc(smem_u32(Xh) + xoff, kBlk, 1024), xl = make_desc(smem_u32(Xl) + xoff, kBlk, 1024);
          Desc sh_, sl_;
          if (which == 0) {
            const uint32_t soff = (k >> 2) * kBlk + (k & 3) * 32;
            sh_ = make_desc(smem_u32(Sh) + soff, 16, 1024); sl_ = make_desc(smem_u32(Sl) + soff, 16, 1024);
          } else {
            sh_ = make_desc(smem_u32(Sh) + xoff, kBlk, 1024); sl_ = make_desc(smem_u32(Sl) + xoff, kBlk, 1024);
          }
          mma_ss(tmem, sh_, xh, idesc, k > 0);
          mma_ss(tmem, sh_, xl, idesc, true);
          mma_ss(tmem, sl_, xh, idesc, true);
        }
        mma_commit(&sh->bar);
      }
      __syncwarp();
    }
    mbar_wait(&sh->bar, phase, wc, 2);
    phase ^= 1;
    tc_fence_after();
    // partial gradient tile [128 rows x 128 features] -> fp32 adds into G (rows i for G_a, rows j for G_b)
    {
      const int r = (warp & 3) * 32 + lane;
      const int g = (which ? j0 : i0) + r, n = which ? p.Bk : p.Bq;
      const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = (warp >> 2) * 64 + cc * 32;
        uint32_t v[32];
        tmem_ld32(lane_base + c0, v);
        tmem_wait_ld();
        if (g < n) {                                     // 16-byte vector reductions: a quarter of the instructions
          float* dst = G + (size_t)g * p.D + d0 + c0;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "f"(__uint_as_float(v[4 * q])),
                         "f"(__uint_as_float(v[4 * q + 1])), "f"(__uint_as_float(v[4 * q + 2])), "f"(__uint_as_float(v[4 * q + 3]))
                         : "memory");
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                    // accumulator and xhat tiles are free for the other product
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, kT);
}

// normalise-backward over the rows of G_a / G_b: dx = (g - xhat (xhat . g)) / max(|x|, eps); one warp per row
template <typename T>
__global__ void __launch_bounds__(256) sim_tc_normbwd_kernel(SimParams p, const float* __restrict__ Ga, const float* __restrict__ Gb,
                                                             int rows_a) {
  const int lane = threadIdx.x & 31;
  int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const bool a_side = r < rows_a;
  if (!a_side) r -= rows_a;
  const float* G = a_side ? Ga : Gb;
  const int n = a_side ? p.Bq : p.Bk;
  if (!G || r >= n) return;
  const T* X = static_cast<const T*>(a_side ? p.a : p.b) + (size_t)r * p.D;
  T* DX = static_cast<T*>(a_side ? p.da : p.db) + (size_t)r * p.D;
  const float ix = (a_side ? p.inv_a : p.inv_b)[r];
  const float* g = G + (size_t)r * p.D;
  float proj = 0.f;
  for (int d = lane * 4; d < p.D; d += 128) {
    const float4 gv = *reinterpret_cast<const float4*>(g + d), xv = ld4(X + d);
    proj += dot4(gv, xv) * ix;
  }
  proj = warp_sum(proj);
  if (ix >= 1.f / kEps) proj = 0.f;            // |x| clamped to eps: x / eps is linear in x
  for (int d = lane * 4; d < p.D; d += 128) {
    const float4 gv = *reinterpret_cast<const float4*>(g + d), xv = ld4(X + d);
    st4(DX + d, make_float4((gv.x - xv.x * ix * proj) * ix, (gv.y - xv.y * ix * proj) * ix,
                            (gv.z - xv.z * ix * proj) * ix, (gv.w - xv.w * ix * proj) * ix));
  }
}

constexpr int kFwdSmem = 4 * kBlk + (512 + 256) * 4 + 64 + 1024;
constexpr int kBwdSmem = 8 * kBlk + 384 * 4 + 64 + 1024;

}  // namespace

bool sim_tc_eligible(int Bq, int Bk, int D) {
  // measured (profiles/r02_sim_tc.jsonl): at 256 x 512 the one-kernel CUDA-core form still wins (27 vs 46 us backward),
  // from 256 x 1024 on the tensor-core form does (48 vs 88 us at D = 512; 50 vs 98-148 us at 256 x 2048)
  return (long long)Bq * Bk >= 256LL * 1024 && D % 128 == 0 && D <= 768 && Bk % 8 == 0;
}

size_t sim_tc_workspace_bytes(int Bq, int Bk, int D) {
  return sim_tc_eligible(Bq, Bk, D) ? sizeof(float) * (size_t)(Bq + Bk) * D : 0;
}

int sim_tc_forward(const SimParams& p, int dtype, cudaStream_t st) {
  dim3 grid((p.Bk + kT - 1) / kT, (p.Bq + kT - 1) / kT);
  if (dtype == XMC_F32) {
    XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(sim_tc_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
    sim_tc_fwd_kernel<float><<<grid, kTcThreads, kFwdSmem, st>>>(p);
  } else {
    XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(sim_tc_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
    sim_tc_fwd_kernel<__nv_bfloat16><<<grid, kTcThreads, kFwdSmem, st>>>(p);
  }
  return cuda_fail(cudaGetLastError(), "sim_tc_fwd_kernel launch");
}

int sim_tc_backward(const SimParams& p, int dtype, void* ws, size_t ws_bytes, cudaStream_t st) {
  const size_t na = (size_t)p.Bq * p.D, nb = (size_t)p.Bk * p.D;
  XMC_REQUIRE(ws && ws_bytes >= sizeof(float) * (na + nb), XMC_ERR_WORKSPACE,
              "similarity-loss backward needs %zu workspace bytes (xmc_simloss_workspace_bytes), got %zu",
              sizeof(float) * (na + nb), ws_bytes);
  float* Ga = p.da ? static_cast<float*>(ws) : nullptr;
  float* Gb = p.db ? static_cast<float*>(ws) + na : nullptr;
  if (!Ga && !Gb) return XMC_OK;
  if (Ga) XMC_RETURN_IF_CUDA(cudaMemsetAsync(Ga, 0, na * sizeof(float), st));
  if (Gb) XMC_RETURN_IF_CUDA(cudaMemsetAsync(Gb, 0, nb * sizeof(float), st));
  dim3 grid((p.Bk + kT - 1) / kT, (p.Bq + kT - 1) / kT, p.D / kT);
  const int rows_a = Ga ? p.Bq : 0, rows = rows_a + (Gb ? p.Bk : 0);
  if (dtype == XMC_F32) {
    XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(sim_tc_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    sim_tc_bwd_kernel<float><<<grid, kTcThreads, kBwdSmem, st>>>(p, Ga, Gb);
    sim_tc_normbwd_kernel<float><<<(rows + 7) / 8, 256, 0, st>>>(p, Ga, Gb, rows_a);
  } else {
    XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(sim_tc_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
    sim_tc_bwd_kernel<__nv_bfloat16><<<grid, kTcThreads, kBwdSmem, st>>>(p, Ga, Gb);
    sim_tc_normbwd_kernel<__nv_bfloat16><<<(rows + 7) / 8, 256, 0, st>>>(p, Ga, Gb, rows_a);
  }
  return cuda_fail(cudaGetLastError(), "sim_tc_bwd_kernel launch");
}

}  // namespace xmc
