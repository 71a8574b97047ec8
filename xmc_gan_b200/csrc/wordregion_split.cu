// wordregion_split.cu — the fp32-TOLERANCE word-region path on the 5th-gen tensor cores (XMC_PATH_FP32_TCGEN05).
//
// BASELINE config 2 names two modes, fp32 (rel 1e-4) and bf16-in/fp32-accumulate (rel 2e-2).  wordregion_tc.cu is the
// second; the first used to run on CUDA cores (wordregion_f32.cu, 35-58 ms per COCO-256 step).  A single bf16 or tf32
// product cannot meet 1e-4, but every fp32 operand can be carried as TWO bf16 numbers, x = hi + lo
// (|x - hi - lo| <= 2^-17 |x|), and every product as three MMAs, hi*hi + hi*lo + lo*hi, accumulated in fp32 in TMEM.
// Emulated on the CPU against the fp64 oracle this keeps loss and gradients of the word loss within 4e-6
// (plain bf16 operands: 2e-3); measured on the GPU: tests/test_wordregion_split_gpu.py.
//
// Same mathematics and the same per-(image, word) state as the other two paths (spec: oracle/word_region.py; the
// loss is the `word_loss` the reference names at xmc_gan/train_gan.py:220-222, 267-269 and never implements):
//   forward   S = Q Khat^T -> P = exp(rho1 (S - 1)), P' = P ||v_r|| -> C += P' Khat -> lsum, cnorm, rel, and C itself
//             (kept for the backward as a hi/lo bf16 pair: as many bytes as fp32)
//   backward  S, W = C Khat^T -> X, Y (closed form) -> dQ += X Khat,  dK^T = C^T Y + Q^T X  -> fp32 reductions
// Operands are split ONCE per call into hi / lo bf16 planes in the caller's workspace (split_planes_kernel); the CTAs
// fetch tiles of those planes by TMA (3-D plane maps, 128B swizzle: no arithmetic on the way) and ONE thread issues the
// tcgen05 MMAs.  The schedule is mostly synchronous — stage, multiply, wait — : this is the precise mode, 12x faster than
// the CUDA-core kernels it replaces and simple enough to audit; the throughput path is wordregion_tc.cu.  What is
// overlapped: in the backward every staged operand half serves a dK^T product of chunk c AND a score product of chunk
// c + 1 (four stagings per chunk instead of seven) and the next region chunk lands under them.
// Shared memory bounds the shape of the backward: Q and C tiles (128 x 256, hi + lo = 128 KB each) cannot both stay
// resident, so they pass through ONE 64 KB buffer in feature halves (S and W accumulate over the halves; each half is
// one M-tile of dK^T); a second buffer — the double buffering that would hide the stagings, a third of the backward's
// cycles in profiles/r02_split_kernels.json — does not fit next to the region chunk and the X / Y tiles (192 KB).
#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"
#include "wordregion.h"

namespace xmc {

using namespace tc;

namespace {

constexpr int TMs = 128;               // word rows per tile (UMMA M)
constexpr int CHs = 64;                // regions per chunk (UMMA N of the score products)
constexpr int kABlk = TMs * 128;       // [128 rows x 64 bf16] swizzled block: 16 KB
constexpr int kKBlk = CHs * 128;       // [64 rows x 64 bf16] swizzled block: 8 KB
constexpr int kThreads = 256;
constexpr float kLog2eS = 1.4426950408889634f;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// The operand planes are read by TMA: 3-D maps (feature, row, outer), boxes of 64 features x {128 word rows | 64 region rows},
// 128B swizzle (the layout the UMMA descriptors expect); rows past the end of a plane's middle dimension arrive as zeros.
struct Maps { CUtensorMap qh, ql, ch, cl, kh, kl; };

// nblk boxes of [rows x 64 features] starting at feature block blk0, rows from row0 of slice `outer`, into consecutive
// swizzled blocks blk_bytes apart; completion (bytes) on `bar`
__device__ __forceinline__ void tma_tile(uint8_t* dst, int blk_bytes, const CUtensorMap* m, int blk0, int nblk, int row0, int outer,
                                         uint64_t* bar) {
  for (int b = 0; b < nblk; ++b) tma_load_3d(dst + b * blk_bytes, m, (blk0 + b) * 64, row0, outer, bar);
}

__device__ __forceinline__ void bulk_load_1d_s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void st_chunk16(uint8_t* blk, int r, int c, uint4 v) {
  *reinterpret_cast<uint4*>(blk + r * 128 + ((c ^ (r & 7)) << 4)) = v;
}

// x[0..8) -> bf16 hi and bf16 lo = rn(x - hi), each as one 16-byte chunk
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

// fp32 operand -> hi / lo bf16 planes (8 elements per thread)
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, size_t n8, __nv_bfloat16* __restrict__ hi,
                                                            __nv_bfloat16* __restrict__ lo) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (size_t)gridDim.x * 256) {
    const float4 u = __ldg(reinterpret_cast<const float4*>(x) + 2 * i), v = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
    const float f[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
    uint4 h, l;
    split8(f, h, l);
    reinterpret_cast<uint4*>(hi)[i] = h;
    reinterpret_cast<uint4*>(lo)[i] = l;
  }
}

struct SplitParams {
  const __nv_bfloat16* qh; const __nv_bfloat16* ql;     // [NQ, D] planes
  const __nv_bfloat16* kh; const __nv_bfloat16* kl;     // [Bi, Rpad, D] planes
  __nv_bfloat16* ch; __nv_bfloat16* cl;                 // [Bi, NQ, D] planes of the context sums (forward writes, backward reads)
  const float* rnorm;
  int NQ, Bi, R, Rpad;
  const int* nq_dev;
  float rho1;
  float* lsum; float* cnorm; float* rel;
  const float* grel;
  float* dqn; float* dkn; float* drnorm;
  int* err;
};

// Phase timing of CTA (0, 0), hooks build only (profiles/exp_split.py): cycles accumulated per phase into the
// workspace header, [16 B error word | 8 B x 14 phases].
#ifdef XMC_TEST_HOOKS
#define XMC_PHASE(k)                                                                   \
  do {                                                                                 \
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) {                      \
      const long long t_ = clock64();                                                  \
      reinterpret_cast<long long*>(p.err + 4)[k] += t_ - t_phase;                      \
      t_phase = t_;                                                                    \
    }                                                                                  \
  } while (0)
#define XMC_PHASE_INIT() long long t_phase = clock64()
#else
#define XMC_PHASE(k) do {} while (0)
#define XMC_PHASE_INIT() do {} while (0)
#endif

struct Ctl {
  uint64_t bar;            // MMA completion
  uint64_t ld;             // TMA completion (word / context tiles)
  uint64_t ldk;            // TMA completion (region chunks: in flight under other work)
  uint64_t ld2;            // forward: the second region buffer
  uint32_t tmem_slot;
  int abort_flag;
};

__device__ __forceinline__ uint8_t* align1k(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

// three MMAs of one k-step: hi*hi (+)= , hi*lo +=, lo*hi +=          (SS form)
__device__ __forceinline__ void mma3_ss(uint32_t d, Desc ah, Desc al, Desc bh, Desc bl, uint32_t idesc, bool acc) {
  mma_ss(d, ah, bh, idesc, acc);
  mma_ss(d, ah, bl, idesc, true);
  mma_ss(d, al, bh, idesc, true);
}

// =====================================================================================================
// forward
// =====================================================================================================
template <int D>
struct FwdS {
  // the word tile's lo plane (its hi plane lives in tensor memory) and TWO region buffers (hi + lo planes each): the chunk
  // after the one being multiplied lands under its MMAs and softmax
  static constexpr int kKBuf = 2 * (D / 64) * kKBlk;           // 64 KB: hi blocks, then lo blocks
  static constexpr int kOffQl = 0, kOffK = (D / 64) * kABlk, kOffRn = kOffK + 2 * kKBuf, kOffPart = kOffRn + 2 * CHs * 4,
                       kOffStg = kOffPart + 3 * 2 * TMs * 4, kOffCtl = kOffStg + (kThreads / 32) * 32 * kStgPitch * 4,
                       kBytes = kOffCtl + 64 + 1024;
  static constexpr int kColC = 0, kColS = D, kColQ = D + CHs;  // C 256 | S 64 | Q_hi 128 (bf16 pairs)
  static_assert(kColQ + D / 2 <= 512, "TMEM budget");
  static_assert(kBytes <= 232448, "shared memory budget");
};

template <int D>
__global__ void __launch_bounds__(kThreads, 1) wr_fwd_split_kernel(const __grid_constant__ Maps maps, SplitParams p) {
  using L = FwdS<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1k(smem_raw);
  uint8_t* Ql = smem + L::kOffQl;
  uint8_t* Kb = smem + L::kOffK;                                       // [2 buffers][hi blocks | lo blocks]
  float* rn_s = reinterpret_cast<float*>(smem + L::kOffRn);            // [2 buffers][64]
  float* part = reinterpret_cast<float*>(smem + L::kOffPart);          // [3: l, a, |C|^2][2 halves][128]
  uint32_t* stg_all = reinterpret_cast<uint32_t*>(smem + L::kOffStg);  // per warp: [32 rows x 16 words] on their way out
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + L::kOffCtl);
  const WaitCtx wc{&ctl->abort_flag, p.err};

  const int tid = threadIdx.x, warp = warp_index(), lane = tid & 31;
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;             // thread = TMEM lane (word row) x column half
  const int NQv = p.nq_dev ? min(p.NQ, __ldg(p.nq_dev)) : p.NQ;
  const int m0 = blockIdx.x * TMs;
  if (m0 >= NQv) return;
  const int grow = m0 + row;
  const int nch = (p.Rpad + CHs - 1) / CHs;
  const bool has_rn = p.rnorm != nullptr;
  const float c1 = p.rho1 * kLog2eS;

  // ctl->ld: the word tile's lo plane; ctl->ldk / ctl->bar2: the two region buffers
  uint64_t* kfull[2] = {&ctl->ldk, &ctl->ld2};
  if (tid == 0) {
    ctl->abort_flag = 0;
    mbar_init(&ctl->bar, 1); mbar_init(&ctl->ld, 1); mbar_init(&ctl->ldk, 1); mbar_init(&ctl->ld2, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 2 * CHs; i += kThreads) rn_s[i] = has_rn ? 0.f : 1.f;      // slots past a chunk's rows stay finite
  if (warp == 0) tmem_alloc(&ctl->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl->tmem_slot;
  const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t phase = 0, kphase[2] = {0, 0};

  // region chunk (img_, c_) -> buffer b_, asynchronously: both planes by TMA (rows past Rpad arrive as zeros), its norms by a bulk copy
  auto load_regions = [&](int img_, int c_, int b_) {
    if (tid == 0) {
      const int n_ = min(CHs, p.Rpad - c_ * CHs);
      uint8_t* kh = Kb + b_ * L::kKBuf;
      mbar_expect_tx(kfull[b_], L::kKBuf + (has_rn ? n_ * 4 : 0));
      tma_tile(kh, kKBlk, &maps.kh, 0, D / 64, c_ * CHs, img_, kfull[b_]);
      tma_tile(kh + (D / 64) * kKBlk, kKBlk, &maps.kl, 0, D / 64, c_ * CHs, img_, kfull[b_]);
      if (has_rn) bulk_load_1d_s(rn_s + b_ * CHs, p.rnorm + (size_t)img_ * p.Rpad + c_ * CHs, n_ * 4, kfull[b_]);
    }
  };

  // the word tile: lo plane -> shared memory (TMA), hi plane -> tensor memory (this thread's half of its row: 32-bit column c
  // holds elements 2c, 2c + 1), so that two of the three MMAs of every score k-step read only the region tile from shared memory
  if (tid == 0) {
    mbar_expect_tx(&ctl->ld, (D / 64) * kABlk);
    tma_tile(Ql, kABlk, &maps.ql, 0, D / 64, m0, 0, &ctl->ld);
  }
  load_regions(blockIdx.y, 0, 0);
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.qh + (size_t)grow * D + half * (D / 2));
#pragma unroll
    for (int b = 0; b < D / 64; ++b) {                                  // 32 bf16 = 16 columns per store
      uint32_t v[16];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint4 w = (grow < p.NQ) ? __ldg(src + b * 4 + u) : make_uint4(0, 0, 0, 0);
        v[4 * u + 0] = w.x; v[4 * u + 1] = w.y; v[4 * u + 2] = w.z; v[4 * u + 3] = w.w;
      }
      tmem_st16(lane_base + L::kColQ + half * (D / 4) + b * 16, v);
    }
    tmem_wait_st();
  }
  mbar_wait(&ctl->ld, 0, wc, 30);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  int x = 0;                                                    // running chunk index: buffer x & 1
  for (int img = blockIdx.y; img < p.Bi; img += gridDim.y) {
    float l = 0.f, a = 0.f;
    for (int c = 0; c < nch; ++c, ++x) {
      const int n = min(CHs, p.Rpad - c * CHs);
      const int b = x & 1;
      uint8_t* Kh = Kb + b * L::kKBuf;
      uint8_t* Kl = Kh + (D / 64) * kKBlk;
      // the other buffer is free (the MMAs of the previous chunk were waited for): fetch the next chunk into it
      if (c + 1 < nch) load_regions(img, c + 1, b ^ 1);
      else if (img + (int)gridDim.y < p.Bi) load_regions(img + gridDim.y, 0, b ^ 1);
      mbar_wait(kfull[b], kphase[b], wc, 33); kphase[b] ^= 1;
      if (warp == 0) {
        if (elect_one()) {                                      // S = Q Khat^T: hi*hi, hi*lo (A from tensor memory), lo*hi (A from shared memory)
          tc_fence_after();
          constexpr uint32_t idesc = idesc_bf16(TMs, CHs, false, false);
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t ao = (k >> 2) * kABlk + (k & 3) * 32, bo = (k >> 2) * kKBlk + (k & 3) * 32;
            const Desc bh = make_desc(smem_u32(Kh) + bo, 16, 1024), bl = make_desc(smem_u32(Kl) + bo, 16, 1024);
            mma_ts(tmem + L::kColS, tmem + L::kColQ + k * 8, bh, idesc, k > 0);
            mma_ts(tmem + L::kColS, tmem + L::kColQ + k * 8, bl, idesc, true);
            mma_ss(tmem + L::kColS, make_desc(smem_u32(Ql) + ao, 16, 1024), bh, idesc, true);
          }
          mma_commit(&ctl->bar);
        }
        __syncwarp();
      }
      mbar_wait(&ctl->bar, phase, wc, 31); phase ^= 1;
      tc_fence_after();
      {   // softmax numerators of this thread's 32 columns; P' = hi + lo written back over S (A operand of the next product)
        uint32_t sv[32];
        tmem_ld32(lane_base + L::kColS + half * 32, sv);
        tmem_wait_ld();
        uint32_t ph[16], pl[16];
        const int r0 = c * CHs + half * 32;
        const float* rn = rn_s + b * CHs + half * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float pw[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float s = __uint_as_float(sv[j + e]);
            const float pv = (r0 + j + e) < p.R ? ex2f(fmaf(s, c1, -c1)) : 0.f;
            l += pv;
            pw[e] = pv * rn[j + e];
            a = fmaf(pw[e], s, a);
          }
          const __nv_bfloat162 hv = __floats2bfloat162_rn(pw[0], pw[1]);
          const float2 hf = __bfloat1622float2(hv);
          ph[j >> 1] = *reinterpret_cast<const uint32_t*>(&hv);
          pl[j >> 1] = pack_bf16(pw[0] - hf.x, pw[1] - hf.y);
        }
        tmem_st16(lane_base + L::kColS + half * 32, ph);             // hi: 16 packed columns
        tmem_st16(lane_base + L::kColS + half * 32 + 16, pl);        // lo: the next 16
        tmem_wait_st();
      }
      tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        if (elect_one()) {                                      // C += P' Khat (A from tensor memory, B = the same smem bytes, MN-major)
          tc_fence_after();
          constexpr uint32_t idesc2 = idesc_bf16(TMs, D, false, true);
#pragma unroll
          for (int ks = 0; ks < CHs / 16; ++ks) {
            if (ks * 16 < n) {
              const uint32_t ah = tmem + L::kColS + (ks >> 1) * 32 + (ks & 1) * 8, al = ah + 16;
              const Desc bh = make_desc(smem_u32(Kh) + ks * 2048, kKBlk, 1024), bl = make_desc(smem_u32(Kl) + ks * 2048, kKBlk, 1024);
              mma_ts(tmem + L::kColC, ah, bh, idesc2, c > 0 || ks > 0);
              mma_ts(tmem + L::kColC, ah, bl, idesc2, true);
              mma_ts(tmem + L::kColC, al, bh, idesc2, true);
            }
          }
          mma_commit(&ctl->bar);
        }
        __syncwarp();
      }
      mbar_wait(&ctl->bar, phase, wc, 32); phase ^= 1;           // this region buffer and S are free again
      tc_fence_after();
    }
    // ---- per image: context sums -> hi/lo planes, |C|^2; statistics of the row ----
    {
      float c2 = 0.f;
      // a thread owns one ROW of C: its 16-byte pieces go through the warp's staging block and leave as 64-byte row segments
      // (stored straight from registers every warp instruction touched 32 rows, half a sector each)
      const int wrow0 = m0 + (warp & 3) * 32;                   // first row of this warp's 32
      const size_t ow = ((size_t)img * p.NQ + wrow0) * D + half * (D / 2);
      uint32_t* stg = stg_all + warp * 32 * kStgPitch;
#pragma unroll 1
      for (int b = 0; b < D / 64; ++b) {                        // this thread's D/2 columns, 32 at a time
        uint32_t cv[32];
        tmem_ld32(lane_base + L::kColC + half * (D / 2) + b * 32, cv);
        tmem_wait_ld();
        uint32_t hw[16], lw[16];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { x[e] = __uint_as_float(cv[8 * u + e]); c2 = fmaf(x[e], x[e], c2); }
          uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;          // rows of a live tile past the count: zeros (the backward's TMA reads them)
          if (grow < NQv) split8(x, hi, lo);
          hw[4 * u] = hi.x; hw[4 * u + 1] = hi.y; hw[4 * u + 2] = hi.z; hw[4 * u + 3] = hi.w;
          lw[4 * u] = lo.x; lw[4 * u + 1] = lo.y; lw[4 * u + 2] = lo.z; lw[4 * u + 3] = lo.w;
        }
        if (p.ch) {                                            // uniform: the contexts are saved for a backward
          warp_rows_out(stg, lane, hw, [&](int r, int wo, uint4 val) {
            if (wrow0 + r < p.NQ) *reinterpret_cast<uint4*>(p.ch + ow + (size_t)r * D + b * 32 + 2 * wo) = val;
          });
          warp_rows_out(stg, lane, lw, [&](int r, int wo, uint4 val) {
            if (wrow0 + r < p.NQ) *reinterpret_cast<uint4*>(p.cl + ow + (size_t)r * D + b * 32 + 2 * wo) = val;
          });
        }
      }
      part[(0 * 2 + half) * TMs + row] = l;
      part[(1 * 2 + half) * TMs + row] = a;
      part[(2 * 2 + half) * TMs + row] = c2;
      tc_fence_before();
      __syncthreads();                                          // C may be overwritten by the next image; partials visible
      tc_fence_after();
      if (half == 0 && grow < NQv) {
        const float lt = part[row] + part[TMs + row], at = part[2 * TMs + row] + part[3 * TMs + row];
        const float ct = part[4 * TMs + row] + part[5 * TMs + row];
        const float cn = sqrtf(ct) / lt;
        const size_t o2 = (size_t)img * p.NQ + grow;
        p.lsum[o2] = lt;
        p.cnorm[o2] = cn;
        p.rel[o2] = (at / lt) / fmaxf(cn, kEps);
      }
      __syncthreads();                                          // partials are rewritten by the next image
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// =====================================================================================================
// backward
// =====================================================================================================
template <int D>
struct BwdS {
  static_assert(D == 256, "feature halves of 128 columns: each is one M = 128 tile of dK^T");
  static constexpr int kHalf = D / 2;                    // features per half
  static constexpr int kHB = kHalf / 64;                 // 64-feature blocks per half
  static constexpr int kOffAh = 0, kOffAl = kHB * kABlk, kOffKh = 2 * kHB * kABlk, kOffKl = kOffKh + (D / 64) * kKBlk,
                       kOffXh = kOffKl + (D / 64) * kKBlk, kOffXl = kOffXh + kABlk, kOffYh = kOffXl + kABlk, kOffYl = kOffYh + kABlk,
                       kOffRn = kOffYl + kABlk, kOffCol = kOffRn + CHs * 4, kOffCtl = kOffCol + CHs * 4, kBytes = kOffCtl + 64 + 1024;
  // dK^T has one 64-column tile per feature half: the drain of one half runs under the MMAs of the other
  static constexpr int kColDQ = 0, kColS = D, kColW = D + CHs, kColDK = D + 2 * CHs;
  static_assert(kColDK + 2 * CHs <= 512, "TMEM budget");
  static_assert(kBytes <= 232448, "shared memory budget");
};

template <int D>
__global__ void __launch_bounds__(kThreads, 1) wr_bwd_split_kernel(const __grid_constant__ Maps maps, SplitParams p) {
  using L = BwdS<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1k(smem_raw);
  uint8_t* Ah = smem + L::kOffAh; uint8_t* Al = smem + L::kOffAl;      // one feature half of Q or of C, hi / lo
  uint8_t* Kh = smem + L::kOffKh; uint8_t* Kl = smem + L::kOffKl;
  uint8_t* Xh = smem + L::kOffXh; uint8_t* Xl = smem + L::kOffXl;
  uint8_t* Yh = smem + L::kOffYh; uint8_t* Yl = smem + L::kOffYl;
  float* rn_s = reinterpret_cast<float*>(smem + L::kOffRn);
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + L::kOffCtl);
  const WaitCtx wc{&ctl->abort_flag, p.err};

  const int tid = threadIdx.x, warp = warp_index(), lane = tid & 31;
  const int row = (warp & 3) * 32 + lane, half = warp >> 2;
  const int NQv = p.nq_dev ? min(p.NQ, __ldg(p.nq_dev)) : p.NQ;
  const int m0 = blockIdx.x * TMs;
  if (m0 >= NQv) return;
  const int grow = m0 + row;
  const int nch = (p.Rpad + CHs - 1) / CHs;
  const bool has_rn = p.rnorm != nullptr;
  const float c1 = p.rho1 * kLog2eS;

  if (tid == 0) { ctl->abort_flag = 0; mbar_init(&ctl->bar, 1); mbar_init(&ctl->ld, 1); mbar_init(&ctl->ldk, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&ctl->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl->tmem_slot;
  const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t phase = 0, ld_phase = 0, ldk_phase = 0;
  bool dq_started = false;
  XMC_PHASE_INIT();

  // one feature half of a [128 x D] operand (hi / lo planes) -> the A buffer.  Every MMA that read the buffer has been
  // waited for by all threads (issue()), so one thread may refill it.
  auto stage_begin = [&](const CUtensorMap* hi, const CUtensorMap* lo, int row0, int outer, int h) {
    if (tid == 0) {
      mbar_expect_tx(&ctl->ld, 2 * L::kHB * kABlk);
      tma_tile(Ah, kABlk, hi, h * L::kHB, L::kHB, row0, outer, &ctl->ld);
      tma_tile(Al, kABlk, lo, h * L::kHB, L::kHB, row0, outer, &ctl->ld);
    }
  };
  auto stage_wait = [&]() { mbar_wait(&ctl->ld, ld_phase, wc, 40); ld_phase ^= 1; };
  auto stage_half = [&](const CUtensorMap* hi, const CUtensorMap* lo, int row0, int outer, int h) {
    stage_begin(hi, lo, row0, outer, h);
    stage_wait();
  };
  auto issue_start = [&](auto&& body) {                 // one elected thread issues ...
    if (warp == 0) {
      if (elect_one()) { tc_fence_after(); body(); mma_commit(&ctl->bar); }
      __syncwarp();
    }
  };
  auto issue_wait = [&]() {                             // ... everybody waits for completion
    mbar_wait(&ctl->bar, phase, wc, 41); phase ^= 1;
    tc_fence_after();
  };
  auto issue = [&](auto&& body) { issue_start(body); issue_wait(); };
  // [128 x 64] += A_half[128 x kHalf] . Khat_half^T : score product over one feature half
  auto scores_half = [&](uint32_t d_col, int h, bool first) {
    constexpr uint32_t idesc = idesc_bf16(TMs, CHs, false, false);
#pragma unroll
    for (int k = 0; k < L::kHalf / 16; ++k) {
      const uint32_t ao = (k >> 2) * kABlk + (k & 3) * 32, bo = (h * L::kHB + (k >> 2)) * kKBlk + (k & 3) * 32;
      mma3_ss(tmem + d_col, make_desc(smem_u32(Ah) + ao, 16, 1024), make_desc(smem_u32(Al) + ao, 16, 1024),
              make_desc(smem_u32(Kh) + bo, 16, 1024), make_desc(smem_u32(Kl) + bo, 16, 1024), idesc, !(first && k == 0));
    }
  };
  // dK^T[half] [kHalf x 64] (+)= A_half^T . Z   (A and Z as MN-major operands; contraction over the 128 word rows)
  auto dk_half = [&](int h, uint8_t* Zh, uint8_t* Zl, bool first) {
    constexpr uint32_t idesc = idesc_bf16(L::kHalf, CHs, true, true);
#pragma unroll
    for (int kt = 0; kt < TMs / 16; ++kt) {
      const uint32_t o = kt * 2048;
      mma3_ss(tmem + L::kColDK + h * CHs, make_desc(smem_u32(Ah) + o, kABlk, 1024), make_desc(smem_u32(Al) + o, kABlk, 1024),
              make_desc(smem_u32(Zh) + o, kABlk, 1024), make_desc(smem_u32(Zl) + o, kABlk, 1024), idesc, !(first && kt == 0));
    }
  };

  // Region chunk (c_, img_) -> the region buffers, asynchronously (TMA); its norms -> rn_s.  Called once the previous
  // chunk's dQ product has executed (the last reader of the buffers); wait_regions() before the first use.
  auto load_regions = [&](int img_, int c_) {
    const int n_ = min(CHs, p.Rpad - c_ * CHs);
    if (tid == 0) {
      mbar_expect_tx(&ctl->ldk, 2 * (D / 64) * kKBlk);
      tma_tile(Kh, kKBlk, &maps.kh, 0, D / 64, c_ * CHs, img_, &ctl->ldk);
      tma_tile(Kl, kKBlk, &maps.kl, 0, D / 64, c_ * CHs, img_, &ctl->ldk);
    }
    if (tid < CHs) rn_s[tid] = (has_rn && tid < n_) ? __ldg(p.rnorm + (size_t)img_ * p.Rpad + c_ * CHs + tid) : (has_rn ? 0.f : 1.f);
  };
  auto wait_regions = [&]() { mbar_wait(&ctl->ldk, ldk_phase, wc, 42); ldk_phase ^= 1; };

  // Order of one chunk c (S, W of the chunk already in tensor memory, the A buffer holding the second half of C):
  //   elementwise X, Y | dQ(c) + dK^T[1] = C_1^T Y | prefetch regions(c+1)
  //   A <- Q_1:  dK^T[1] += Q_1^T X,  S(c+1)  = Q_1 K_1^T      | drain dK^T[1]
  //   A <- Q_0:  dK^T[0]  = Q_0^T X,  S(c+1) += Q_0 K_0^T
  //   A <- C_0:  dK^T[0] += C_0^T Y,  W(c+1)  = C_0 K_0^T      | drain dK^T[0]
  //   A <- C_1:                       W(c+1) += C_1 K_1^T
  // so every staged half serves a dK^T product of chunk c AND a score product of chunk c+1: four stagings per chunk
  // instead of seven, and the region chunk arrives under them.  The first chunk of an image computes its scores alone.
  bool regions_pending = false;
  for (int img = blockIdx.y; img < p.Bi; img += gridDim.y) {
    float inv_l = 1.f, gam = 0.f, ngrl = 0.f;
    if (grow < NQv) {
      const size_t o = (size_t)img * p.NQ + grow;
      const float inv_cn = 1.f / fmaxf(p.cnorm[o], kEps);
      inv_l = 1.f / p.lsum[o];
      gam = p.grel[o] * inv_cn;
      ngrl = -gam * p.rel[o] * inv_cn * inv_l;          // the saved context is the unscaled sum C = l c
    }
    // ---- scores of the image's first chunk ----
    if (!regions_pending) load_regions(img, 0);
    wait_regions();
    regions_pending = false;
    XMC_PHASE(0);
    for (int h = 0; h < 2; ++h) {
      stage_half(&maps.qh, &maps.ql, m0, 0, h);
      XMC_PHASE(1);
      issue([&] { scores_half(L::kColS, h, h == 0); });
      XMC_PHASE(2);
    }
    for (int h = 0; h < 2; ++h) {
      stage_half(&maps.ch, &maps.cl, m0, img, h);
      XMC_PHASE(3);
      issue([&] { scores_half(L::kColW, h, h == 0); });
      XMC_PHASE(4);
    }
    for (int c = 0; c < nch; ++c) {
      const int n = min(CHs, p.Rpad - c * CHs);
      const bool more = c + 1 < nch;                    // another chunk of this image follows: its scores ride along
      __syncthreads();                                  // rn_s of this chunk visible
      // ---- X = dS + gamma alpha', Y = -gamma rel alpha' / (l |c|): this thread's 32 columns, split into hi / lo tiles ----
      {
        uint32_t sv[32], wv[32];
        tmem_ld32(lane_base + L::kColS + half * 32, sv);
        tmem_ld32(lane_base + L::kColW + half * 32, wv);
        tmem_wait_ld();
        const int r0 = c * CHs + half * 32;
        float z[32];                                                                        // alpha * d alpha' -> d |v_r|
        // packed fp32x2 arithmetic on column pairs (the same operations in the same order as the scalar form, half the issue slots)
        const float2 cc = make_float2(c1, c1), ncc = make_float2(-c1, -c1), il2 = make_float2(inv_l, inv_l);
        const float2 gam2 = make_float2(gam, gam), ng2 = make_float2(ngrl, ngrl), rho2v = make_float2(p.rho1, p.rho1);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float xv[8], yv[8];
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const int j = u * 8 + e;
            const float2 s2 = make_float2(__uint_as_float(sv[j]), __uint_as_float(sv[j + 1]));
            const float2 w2 = make_float2(__uint_as_float(wv[j]), __uint_as_float(wv[j + 1]));
            const float2 arg = __ffma2_rn(s2, cc, ncc);
            float2 al = __fmul2_rn(make_float2(ex2f(arg.x), ex2f(arg.y)), il2);               // alpha
            al.x = (r0 + j) < p.R ? al.x : 0.f;
            al.y = (r0 + j + 1) < p.R ? al.y : 0.f;
            const float2 rn2 = *reinterpret_cast<const float2*>(rn_s + half * 32 + j);
            const float2 alp = __fmul2_rn(al, rn2);                                          // alpha' = alpha |v_r|
            const float2 dap = __ffma2_rn(ng2, w2, __fmul2_rn(gam2, s2));                    // d loss / d alpha'
            const float2 x2 = __fmul2_rn(alp, __ffma2_rn(rho2v, dap, gam2));
            const float2 y2 = __fmul2_rn(ng2, alp);
            const float2 zz = __fmul2_rn(al, dap);
            xv[e] = x2.x; xv[e + 1] = x2.y; yv[e] = y2.x; yv[e + 1] = y2.y; z[j] = zz.x; z[j + 1] = zz.y;
          }
          uint4 hi, lo;
          split8(xv, hi, lo);
          st_chunk16(Xh, row, half * 4 + u, hi); st_chunk16(Xl, row, half * 4 + u, lo);
          split8(yv, hi, lo);
          st_chunk16(Yh, row, half * 4 + u, hi); st_chunk16(Yl, row, half * 4 + u, lo);
        }
        if (has_rn) {                 // column sums over this warp's 32 word rows (31 shuffles), one add per column and warp
          const float colsum = warp_transpose_sum32(z, lane);
          if (r0 + lane < p.R) atomicAdd(p.drnorm + (size_t)img * p.Rpad + r0 + lane, colsum);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      XMC_PHASE(5);
      // ---- dQ += X Khat_chunk;  dK^T[half 1] = C_1^T Y (the A buffer still holds the second half of C) ----
      issue([&] {
        constexpr uint32_t idesc_dq = idesc_bf16(TMs, D, false, true);
#pragma unroll
        for (int ks = 0; ks < CHs / 16; ++ks) {
          if (ks * 16 < n) {
            const Desc xh = make_desc(smem_u32(Xh) + ks * 32, 16, 1024), xl = make_desc(smem_u32(Xl) + ks * 32, 16, 1024);
            const Desc bh = make_desc(smem_u32(Kh) + ks * 2048, kKBlk, 1024), bl = make_desc(smem_u32(Kl) + ks * 2048, kKBlk, 1024);
            mma3_ss(tmem + L::kColDQ, xh, xl, bh, bl, idesc_dq, dq_started || ks > 0);
          }
        }
        dk_half(1, Yh, Yl, true);
      });
      dq_started = true;
      XMC_PHASE(6);
      // the region buffers are free: the next chunk (of this image, or the first of the CTA's next image) lands under the rest
      {
        const int nimg = more ? img : img + (int)gridDim.y, nc = more ? c + 1 : 0;
        if (nimg < p.Bi) { load_regions(nimg, nc); regions_pending = true; }
      }
      // ---- A <- Q_1: dK^T[1] += Q_1^T X; S(c+1) = Q_1 K_1^T ----
      stage_half(&maps.qh, &maps.ql, m0, 0, 1);
      if (more) { wait_regions(); regions_pending = false; }    // both loads were in flight together
      XMC_PHASE(7);
      issue([&] { dk_half(1, Xh, Xl, false); if (more) scores_half(L::kColS, 1, true); });
      XMC_PHASE(8);
      auto drain = [&](int h) {   // thread = feature (TMEM lane) x 32 regions: a warp adds 32 consecutive features of one region
                                  // row (128 bytes).  (Staging the tile transposed in the dead region buffers and one bulk
                                  // reduce-add per region row was tried: 4.03 ms instead of 3.74 ms for the kernel.)
        uint32_t dv[32];
        tmem_ld32(lane_base + L::kColDK + h * CHs + half * 32, dv);
        tmem_wait_ld();
        float* dst = p.dkn + ((size_t)img * p.Rpad + c * CHs + half * 32) * D + h * L::kHalf + row;
        const int nr = p.R - (c * CHs + half * 32);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nr) atomicAdd(dst + (size_t)j * D, __uint_as_float(dv[j]));
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
      };
      // ---- A <- Q_0: dK^T[0] = Q_0^T X; S(c+1) += Q_0 K_0^T; the finished dK^T[1] drains under these MMAs ----
      stage_half(&maps.qh, &maps.ql, m0, 0, 0);
      XMC_PHASE(7);
      issue_start([&] { dk_half(0, Xh, Xl, true); if (more) scores_half(L::kColS, 0, false); });
      drain(1);
      XMC_PHASE(9);
      issue_wait();
      XMC_PHASE(8);
      // ---- A <- C_0: dK^T[0] += C_0^T Y; W(c+1) = C_0 K_0^T ----
      stage_half(&maps.ch, &maps.cl, m0, img, 0);
      XMC_PHASE(7);
      issue([&] { dk_half(0, Yh, Yl, false); if (more) scores_half(L::kColW, 0, true); });
      XMC_PHASE(8);
      // ---- A <- C_1: W(c+1) += C_1 K_1^T (and the buffer is where the next chunk's dK^T[1] expects the second half of C);
      //      the finished dK^T[0] drains under these MMAs ----
      if (more) {
        stage_half(&maps.ch, &maps.cl, m0, img, 1);
        XMC_PHASE(3);
        issue_start([&] { scores_half(L::kColW, 1, false); });
        drain(0);
        XMC_PHASE(9);
        issue_wait();
        XMC_PHASE(4);
      } else {
        drain(0);
        XMC_PHASE(9);
      }
    }
  }
  // ---- dQ of this word tile (summed over the CTA's images) -> fp32 adds.  A thread owns one ROW of the tile: its pieces go
  //      through a per-warp staging block (the X tiles are dead by now) and leave as 64-byte row segments ----
  if (dq_started) {
    uint32_t* stg = reinterpret_cast<uint32_t*>(Xh) + warp * 32 * kStgPitch;
    const int wrow0 = m0 + (warp & 3) * 32;
    float* dstw = p.dqn + (size_t)wrow0 * D + half * (D / 2);
#pragma unroll 1
    for (int b = 0; b < D / 64; ++b) {
      uint32_t dv[32];
      tmem_ld32(lane_base + L::kColDQ + half * (D / 2) + b * 32, dv);
      tmem_wait_ld();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
        warp_rows_out(stg, lane, dv + 16 * hf, [&](int r, int wo, uint4 val) {
          if (wrow0 + r < NQv)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dstw + (size_t)r * D + b * 32 + 16 * hf + wo),
                         "f"(__uint_as_float(val.x)), "f"(__uint_as_float(val.y)), "f"(__uint_as_float(val.z)), "f"(__uint_as_float(val.w))
                         : "memory");
        });
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// workspace: [256 B header: error word, phase counters of the hooks build | qh | ql | kh | kl]   (bf16 planes, 256-byte aligned)
struct Planes { __nv_bfloat16 *qh, *ql, *kh, *kl; };
size_t plane_bytes(size_t n) { return (n * 2 + 255) & ~size_t(255); }
Planes carve(void* ws, int NQ, int Bi, int Rpad, int D) {
  uint8_t* b = static_cast<uint8_t*>(ws) + 256;
  const size_t nq = plane_bytes((size_t)NQ * D), nk = plane_bytes((size_t)Bi * Rpad * D);
  Planes pl;
  pl.qh = reinterpret_cast<__nv_bfloat16*>(b); pl.ql = reinterpret_cast<__nv_bfloat16*>(b + nq);
  pl.kh = reinterpret_cast<__nv_bfloat16*>(b + 2 * nq); pl.kl = reinterpret_cast<__nv_bfloat16*>(b + 2 * nq + nk);
  return pl;
}

int split_operands(const WrParams& w, int D, const Planes& pl, cudaStream_t st) {
  const size_t nq8 = (size_t)w.NQ * D / 8, nk8 = (size_t)w.Bi * w.Rpad * D / 8;
  split_planes_kernel<<<(unsigned)std::min<size_t>((nq8 + 255) / 256, 148 * 8), 256, 0, st>>>(static_cast<const float*>(w.qn), nq8, pl.qh, pl.ql);
  split_planes_kernel<<<(unsigned)std::min<size_t>((nk8 + 255) / 256, 148 * 8), 256, 0, st>>>(static_cast<const float*>(w.kn), nk8, pl.kh, pl.kl);
  return cuda_fail(cudaGetLastError(), "split_planes_kernel launch");
}

typedef CUresult (*PFN_encodeTiledS)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// bf16 plane [outer][rows][D] as a 3-D tensor (feature, row, outer); box = 64 features x box_rows x 1, 128B swizzle
int make_plane_map(CUtensorMap* m, const void* base, int outer, int rows, int D, int box_rows) {
  static PFN_encodeTiledS enc = nullptr;
  if (!enc) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<PFN_encodeTiledS>(ptr);
  }
  XMC_REQUIRE(enc != nullptr, XMC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)rows, (cuuint64_t)outer};
  cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)rows * D * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XMC_REQUIRE(r == CUDA_SUCCESS, XMC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return XMC_OK;
}

// A CTA keeps one word tile and walks every `groups`-th image.  The number of live tiles is only known on the device
// (compacted rows), so the grid cannot be sized to whole waves; four images per CTA keeps the tail below a few percent
// (~1 500 CTAs at COCO-256) while a tile's operands (128 KB) are staged once per four images (320 KB of region chunks each).
int grid_groups(int Bi) { return std::max(1, (Bi + 3) / 4); }

SplitParams fill(const WrParams& w, const Planes& pl, void* ws, int D) {
  SplitParams p{};
  p.qh = pl.qh; p.ql = pl.ql; p.kh = pl.kh; p.kl = pl.kl;
  p.ch = static_cast<__nv_bfloat16*>(w.chat);                        // two planes: hi [Bi, NQ, D], then lo
  p.cl = w.chat ? p.ch + (size_t)w.Bi * w.NQ * D : nullptr;
  p.rnorm = w.rnorm; p.NQ = w.NQ; p.Bi = w.Bi; p.R = w.R; p.Rpad = w.Rpad; p.nq_dev = w.nq_dev; p.rho1 = w.rho1;
  p.lsum = w.lsum; p.cnorm = w.cnorm; p.rel = w.rel; p.grel = w.grel;
  p.dqn = w.dqn; p.dkn = w.dkn; p.drnorm = w.drnorm;
  p.err = static_cast<int*>(ws);
  return p;
}

int make_maps(Maps* m, const SplitParams& p, const WrParams& w, int D) {
  if (int rc = make_plane_map(&m->qh, p.qh, 1, w.NQ, D, TMs)) return rc;
  if (int rc = make_plane_map(&m->ql, p.ql, 1, w.NQ, D, TMs)) return rc;
  if (int rc = make_plane_map(&m->kh, p.kh, w.Bi, w.Rpad, D, CHs)) return rc;
  if (int rc = make_plane_map(&m->kl, p.kl, w.Bi, w.Rpad, D, CHs)) return rc;
  // without saved contexts (forward only, no gradient wanted) the two maps are never used: point them at the words
  if (int rc = make_plane_map(&m->ch, p.ch ? p.ch : p.qh, p.ch ? w.Bi : 1, w.NQ, D, TMs)) return rc;
  if (int rc = make_plane_map(&m->cl, p.cl ? p.cl : p.ql, p.cl ? w.Bi : 1, w.NQ, D, TMs)) return rc;
  return XMC_OK;
}

template <int D>
int launch_fwd(const WrParams& w, void* ws, cudaStream_t st) {
  const Planes pl = carve(ws, w.NQ, w.Bi, w.Rpad, D);
  if (int rc = split_operands(w, D, pl, st)) return rc;
  SplitParams p = fill(w, pl, ws, D);
  Maps maps;
  if (int rc = make_maps(&maps, p, w, D)) return rc;
  const int tiles = (w.NQ + TMs - 1) / TMs;
  dim3 grid(tiles, grid_groups(w.Bi));
  XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(wr_fwd_split_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdS<D>::kBytes));
  wr_fwd_split_kernel<D><<<grid, kThreads, FwdS<D>::kBytes, st>>>(maps, p);
  return cuda_fail(cudaGetLastError(), "wr_fwd_split_kernel launch");
}

template <int D>
int launch_bwd(const WrParams& w, void* ws, cudaStream_t st) {
  const Planes pl = carve(ws, w.NQ, w.Bi, w.Rpad, D);
  if (int rc = split_operands(w, D, pl, st)) return rc;
  SplitParams p = fill(w, pl, ws, D);
  Maps maps;
  if (int rc = make_maps(&maps, p, w, D)) return rc;
  const int tiles = (w.NQ + TMs - 1) / TMs;
  dim3 grid(tiles, grid_groups(w.Bi));
  XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(wr_bwd_split_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdS<D>::kBytes));
  wr_bwd_split_kernel<D><<<grid, kThreads, BwdS<D>::kBytes, st>>>(maps, p);
  return cuda_fail(cudaGetLastError(), "wr_bwd_split_kernel launch");
}

}  // namespace

size_t wordregion_split_workspace_bytes(int NQ, int Bi, int, int Rpad, int D) {
  return 256 + 2 * plane_bytes((size_t)NQ * D) + 2 * plane_bytes((size_t)Bi * Rpad * D);
}

int wordregion_split_forward(const WrParams& w, int D, void* ws, size_t ws_bytes, cudaStream_t st) {
  XMC_REQUIRE(ws && ws_bytes >= wordregion_split_workspace_bytes(w.NQ, w.Bi, w.R, w.Rpad, D), XMC_ERR_WORKSPACE,
              "workspace too small (%zu bytes)", ws_bytes);
  if (D == 256) return launch_fwd<256>(w, ws, st);
  set_error("split-bf16 word-region path supports D = 256 (got %d)", D);
  return XMC_ERR_UNSUPPORTED;
}

int wordregion_split_backward(const WrParams& w, int D, void* ws, size_t ws_bytes, cudaStream_t st) {
  XMC_REQUIRE(ws && ws_bytes >= wordregion_split_workspace_bytes(w.NQ, w.Bi, w.R, w.Rpad, D), XMC_ERR_WORKSPACE,
              "workspace too small (%zu bytes)", ws_bytes);
  XMC_REQUIRE(w.chat != nullptr, XMC_ERR_INVALID_ARG, "the split-bf16 backward needs the contexts saved by the forward (chat)");
  if (D == 256) return launch_bwd<256>(w, ws, st);
  set_error("split-bf16 word-region path supports D = 256 (got %d)", D);
  return XMC_ERR_UNSUPPORTED;
}

}  // namespace xmc
