// wordregion_tc.cu — word–region attention statistics and their fused backward on the 5th-gen tensor
// cores (sm_100a): bf16 operands, fp32 accumulation in TMEM, region tiles streamed by TMA.
//
// Forward.  Work item = (128 consecutive rows of the compacted unit-word matrix [NQ, D]) x (one image).
// One persistent CTA per SM walks its share of the items (SegIter below); it keeps the 128 word rows
// resident in TMEM as the A operand (bf16, D/2 columns) while it stays on a word tile; the regions of
// an image arrive in chunks of 64 rows ([64 x D] bf16, 128B-swizzled TMA boxes) through a 4-stage
// mbarrier ring.
//
//   GEMM1 (TS)  S[128 x 64]  = Q(tmem) . Khat_chunk^T        B = K-major  smem descriptor
//   softmax warps: P' = exp2(c1 (S-1)) * ||v_r||  -> bf16, written over S in TMEM;  l += P, a += P' S
//   GEMM2 (TS)  C[128 x D] += P'(tmem) . Khat_chunk           B = MN-major smem descriptor (same bytes)
//   epilogue warpgroup: C -> bf16 -> swizzled smem boxes -> TMA store (saved for the backward);
//                       ||C||, lsum, cnorm, rel  (per image, per word row)
//
// The key and the value of the attention are the SAME smem tile (raw values are folded into P' as a
// per-column scale), cosines are bounded so the softmax uses the constant shift rho1 (no running
// max, no rescale of C), and the [Bi,Bc,T,R] score tensor lives only in TMEM.
//
// Warp roles (512 threads): warp 0 = TMA producer and warp 1 = MMA issuer (one elected thread each runs
// the whole role), warp 2 = TMEM allocator, warps 4-11 = two softmax warpgroups (thread = TMEM lane =
// word row, 32 chunk columns each), warps 12-15 = epilogue warpgroup.
// TMEM map (512 columns): C [0,D) | Q [D, D+D/2) | S0 | S1 (64 columns each, P' aliases S).
// The backward kernel is described in front of its code.
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"
#include "wordregion.h"

namespace xmc {

using namespace tc;

constexpr int TM = 128;            // word rows per tile (UMMA M)
constexpr int CH = 64;             // region rows per chunk
constexpr float kLog2eTc = 1.4426950408889634f;

// Persistent schedule shared by both kernels.  Work items are (word tile, image) pairs; the grid is one
// CTA per SM and every CTA gets the same number of items (+-1).  A CTA walks SEGMENTS = runs of images
// of ONE word tile (the stationary operand and its accumulator change only between segments); the
// images of a segment are img0 + j*stride.
// The images of a tile are dealt into s = ceil(grid / tiles) interleaved ROWS (row c = images c, c+s,
// c+2s, ...; s = 1 when a CTA owns at least a whole tile), the item list is tile-major, row-major, and
// CTA k owns the contiguous span [k*per, (k+1)*per) of it, which it walks starting at the first row
// start inside the span (wrapping to the span's head at the end).  Every CTA therefore sweeps the
// images in the same direction at nearly the same pace: at any moment the grid works on a window of a
// few images whose region tiles and gradient rows stay L2-resident (a contiguous image split spreads
// the grid over all images: +35% DRAM traffic, measured), and the load is still balanced to one item.
// The number of word rows is read from device memory (compacted captions: the host does not know it),
// so the schedule is computed in the kernel.
struct Seg { int tile, img0, stride, nimg; };
struct SegIter {
  int Bi, s, q, r;                 // rows per tile; Bi = q*s + r: rows c < r hold q+1 images, the others q
  int P, Pend, P0, Pwrap, phase;
  __device__ __forceinline__ void decode(int pos, int& tile, int& c, int& col, int& len) const {
    tile = pos / Bi;
    const int pp = pos - tile * Bi;
    const int big = r * (q + 1);
    if (pp < big) { c = pp / (q + 1); col = pp - c * (q + 1); len = q + 1; }
    else { const int p2 = pp - big; c = r + p2 / q; col = p2 - (c - r) * q; len = q; }
  }
  __device__ __forceinline__ bool next(Seg& sg) {
    if (P >= Pend) {
      if (phase != 0) return false;
      phase = 1; P = P0; Pend = Pwrap;
      if (P >= Pend) return false;
    }
    int tile, c, col, len;
    decode(P, tile, c, col, len);
    const int cnt = min(len - col, Pend - P);
    sg.tile = tile; sg.img0 = col * s + c; sg.stride = s; sg.nimg = cnt;
    P += cnt;
    return true;
  }
};
__device__ __forceinline__ SegIter seg_iter(int NQ, int Bi) {
  const int G = (int)gridDim.x, k = (int)blockIdx.x;
  const int tiles = (NQ + TM - 1) / TM;
  const int W = tiles * Bi;
  const int per = (W + G - 1) / G;
  SegIter it;
  it.Bi = Bi;
  it.s = tiles < G ? min(Bi, (G + tiles - 1) / max(tiles, 1)) : 1;
  it.q = Bi / it.s; it.r = Bi - it.q * it.s;
  it.P0 = min(W, k * per);
  const int P1 = min(W, it.P0 + per);
  int start = it.P0;
  if (it.P0 < P1) {
    int tile, c, col, len;
    it.decode(it.P0, tile, c, col, len);
    if (col != 0 && it.P0 + (len - col) < P1) start = it.P0 + (len - col);    // first row start inside the span
  }
  it.P = start; it.Pend = P1; it.Pwrap = start; it.phase = 0;
  return it;
}
__device__ __forceinline__ int device_rows(const int* nq_dev, int NQ) { return nq_dev ? min(NQ, __ldg(nq_dev)) : NQ; }

constexpr int kFwdThreads = 512;     // 4 role warps + 2 softmax warpgroups + 1 epilogue warpgroup

template <int D>
struct FwdCfg {
  static constexpr int kStageBytes = CH * D * 2;
  static constexpr int kBlockBytes = CH * 128;                    // one [64 rows x 64 bf16] swizzled box
  static constexpr int kStages = (D == 256) ? 4 : 6;
  static constexpr int kOutBytes = TM * 128;                      // context staging: [128 rows x 64 bf16] swizzled box
  static constexpr int kOutBoxes = D / 64;                        // one box per 64 features: a whole image's contexts
  static constexpr int kOffOut = kStages * kStageBytes;           // staging boxes (1024-aligned)
  static constexpr int kOffRn = kOffOut + kOutBoxes * kOutBytes;  // [kStages][64] fp32 region norms of the chunk
  static constexpr int kOffEx = kOffRn + kStages * CH * 4;        // [2 parity][2 wg][2][128] fp32 partial l, a
  static constexpr int kOffBar = kOffEx + 2 * 2 * 2 * TM * 4;
  static constexpr int kSmemBytes = kOffBar + 256 + 1024 /*align*/;
  static constexpr int kColC = 0, kColQ = D, kColS0 = D + D / 2, kColS1 = kColS0 + CH;
  static_assert(kColS1 + CH <= 512, "TMEM budget");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

struct TcFwdParams {
  const __nv_bfloat16* qn;     // [NQ, D]
  const float* rnorm;          // [Bi, Rpad] or null
  int NQ, Bi, R, Rpad;
  float rho1;
  float* lsum; float* cnorm; float* rel;
  int save_ctx;                // context sums C = l * c_t go out through tm_ctx (bf16 [Bi, NQ, D]) for the backward
  const int* nq_dev;           // device count of valid word rows (<= NQ, the row stride of every buffer) or null
  int poison;                  // tests (debug flag 16): fill TMEM with NaNs first, so that a read of a never-written column shows
  int* err;
  float* dbg;                  // optional: S chunk 0 and C of the first tile/image (tests)
  long long* trace;            // perf experiments only (flag 4): clock64 timeline of CTA (0,0), [4 roles][64][4]
};

#define XMC_TRACE(role, g, k)                                                             \
  do {                                                                                   \
    if (p.trace && blockIdx.x == 0 && (g) < 64)                      \
      p.trace[((role) * 64 + (g)) * 4 + (k)] = clock64();                               \
  } while (0)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ long long global_timer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// columns [c0, c0 + n) of this warp's lane quadrant <- 0: score columns beyond a narrow chunk's N are never
// written by the MMAs but are read (and multiplied by a zero weight), so they must hold finite bits
__device__ __forceinline__ void zero_tmem(int quadrant, int c0, int n) {
  uint32_t v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = 0u;
  for (int c = 0; c < n; c += 16) tmem_st16((static_cast<uint32_t>(quadrant * 32) << 16) + c0 + c, v);
  tmem_wait_st();
  tc_fence_before();
}
// tests only: every TMEM column of this warp's lane quadrant <- quiet NaNs
__device__ __forceinline__ void poison_tmem(int quadrant) {
  uint32_t v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = 0x7fc00000u;
  for (int c = 0; c < 512; c += 16) tmem_st16((static_cast<uint32_t>(quadrant * 32) << 16) + c, v);
  tmem_wait_st();
  tc_fence_before();
}

template <int D>
__global__ void __launch_bounds__(kFwdThreads, 1)
wr_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_ctx, TcFwdParams p) {
  using Cfg = FwdCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* kv = smem;                                               // kStages x kStageBytes
  uint8_t* out_s = smem + Cfg::kOffOut;                             // 2 x kOutBytes
  float* rn_s = reinterpret_cast<float*>(smem + Cfg::kOffRn);
  float* ex_s = reinterpret_cast<float*>(smem + Cfg::kOffEx);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* kv_full = bars;                       // [kStages]
  uint64_t* kv_empty = bars + Cfg::kStages;       // [kStages]
  uint64_t* s_full = kv_empty + Cfg::kStages;     // [2]
  uint64_t* p_full = s_full + 2;                  // [2]
  uint64_t* la_full = p_full + 2;                 // [2]  softmax partials of an image are in ex_s
  uint64_t* la_empty = la_full + 2;               // [2]
  uint64_t* c_full = la_empty + 2;
  uint64_t* c_empty = c_full + 1;
  uint64_t* q_ready = c_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);
  int* abort_flag = reinterpret_cast<int*>(tmem_slot + 1);
  const WaitCtx wc{abort_flag, p.err};

  const int warp = warp_index(), lane = threadIdx.x & 31;
  const int NQ = device_rows(p.nq_dev, p.NQ);
  const int nch = (p.Rpad + CH - 1) / CH;
  const bool has_rn = p.rnorm != nullptr;

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(kv_full + s, 1); mbar_init(kv_empty + s, 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(s_full + s, 1); mbar_init(p_full + s, 256);
      mbar_init(la_full + s, 256); mbar_init(la_empty + s, 128);
    }
    mbar_init(c_full, 1); mbar_init(c_empty, 128); mbar_init(q_ready, 256);
    fence_barrier_init();
  }
  // region-norm slots past a chunk's rows are never loaded: make them finite (they multiply weights that are 0)
  for (int i = threadIdx.x; i < Cfg::kStages * CH; i += kFwdThreads) rn_s[i] = 0.f;
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_ctx); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // The allocation is the whole tensor memory (512 columns, one CTA per SM), so its base is lane 0 /
  // column 0; using the literal keeps every TMEM address an immediate for the MMA issue path.
  constexpr uint32_t tmem = 0;
  if (threadIdx.x == 0 && *tmem_slot != 0) { *abort_flag = 1; if (p.err) atomicExch(p.err, 999); }
  if (warp >= 12) {
    if (p.poison) poison_tmem(warp & 3);
    zero_tmem(warp & 3, Cfg::kColS0, 2 * CH);
  }
  __syncthreads();

  {
    if (warp == 0) {
      // ===== TMA producer: region chunks of every (segment, image), one continuous ring =====
      if (elect_one()) {
        int st = 0, ph = 1;                          // ph: parity to wait on kv_empty (first pass free)
        SegIter it = seg_iter(NQ, p.Bi);
        Seg sg;
        while (it.next(sg)) {
          for (int ii = 0; ii < sg.nimg; ++ii) {
            const int img = sg.img0 + ii * sg.stride;
            for (int c = 0; c < nch; ++c) {
              const int n = min(CH, p.Rpad - c * CH);
              mbar_wait(kv_empty + st, ph, wc, 1);
              mbar_expect_tx(kv_full + st, Cfg::kStageBytes + (has_rn ? n * 4 : 0));
              uint8_t* dst = kv + st * Cfg::kStageBytes;
#pragma unroll
              for (int kb = 0; kb < D / 64; ++kb)
                tma_load_3d(dst + kb * Cfg::kBlockBytes, &tm_k, kb * 64, c * CH, img, kv_full + st);
              if (has_rn) bulk_load_1d(rn_s + st * CH, p.rnorm + (size_t)img * p.Rpad + c * CH, n * 4, kv_full + st);
              if (++st == Cfg::kStages) { st = 0; ph ^= 1; }
            }
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ===== MMA issuer: one elected thread runs the whole role =====
      if (elect_one()) {
        const uint32_t kv_addr = smem_u32(kv);
        constexpr uint32_t idesc2 = idesc_bf16(TM, D, false, true);
        // descriptor of stage 0; stages / k-steps are reached by adding (byte offset >> 4) to the low word
        const Desc kdesc0 = make_desc(kv_addr, 16, 1024);                    // K-major   (GEMM1 B)
        const Desc vdesc0 = make_desc(kv_addr, Cfg::kBlockBytes, 1024);      // MN-major  (GEMM2 B)
        auto issue_g1 = [&](int st, int n, int sb) {
          const uint32_t idesc1 = idesc_bf16(TM, n, false, false);
          const Desc base = kdesc0 + ((uint32_t)(st * Cfg::kStageBytes) >> 4);
          const uint32_t d_tmem = tmem + (sb ? Cfg::kColS1 : Cfg::kColS0);
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            mma_ts(d_tmem, tmem + Cfg::kColQ + k * 8, base + (((k >> 2) * Cfg::kBlockBytes + (k & 3) * 32) >> 4), idesc1, k > 0);
        };
        // Issue order on the (in-order) tensor pipe, per segment: G1(0) G1(1) | G2(0) G1(2) | G2(1) G1(3) | ...
        // One exposed synchronisation point per chunk: P(g) written (and, at the first chunk of an image,
        // the previous image's C drained); region chunk g+2 has normally landed long before.
        // Running counters (never reset): x = chunks issued so far (S/P buffer = x & 1, their barrier
        // parity = (x >> 1) & 1), ic = images finished (parity of c_full / c_empty), stage ring (st, st2, ph2).
        int x = 0, ic = 0, seg = 0;
        int st = 0;                                  // stage of chunk g
        int st2 = 0, ph2 = 0;                        // stage / kv_full parity of the chunk G1 is issued for next
        auto advance2 = [&]() { if (++st2 == Cfg::kStages) { st2 = 0; ph2 ^= 1; } };
        SegIter it = seg_iter(NQ, p.Bi);
        Seg sg;
        while (it.next(sg)) {
          const int G = sg.nimg * nch;
          mbar_wait(q_ready, seg & 1, wc, 2);
          int c2 = 0;                                // chunk-in-image of the chunk G1 is issued for next
          mbar_wait(kv_full + st2, ph2, wc, 3);
          tc_fence_after();
          issue_g1(st2, min(CH, p.Rpad), x & 1);
          mma_commit(s_full + (x & 1));
          advance2(); if (++c2 == nch) c2 = 0;
          if (G > 1) {
            mbar_wait(kv_full + st2, ph2, wc, 4);
            tc_fence_after();
            issue_g1(st2, min(CH, p.Rpad - c2 * CH), (x + 1) & 1);
            mma_commit(s_full + ((x + 1) & 1));
            advance2(); if (++c2 == nch) c2 = 0;
          }
          int c = 0;
          for (int g = 0; g < G; ++g, ++x) {
            const int sb = x & 1;
            mbar_wait(p_full + sb, (x >> 1) & 1, wc, 5);
            if (c == 0 && ic > 0) mbar_wait(c_empty, (ic - 1) & 1, wc, 6);
            tc_fence_after();
            XMC_TRACE(0, x, 0);
            const int n = min(CH, p.Rpad - c * CH);
            const Desc vbase = vdesc0 + ((uint32_t)(st * Cfg::kStageBytes) >> 4);
            const uint32_t a_tmem = tmem + (sb ? Cfg::kColS1 : Cfg::kColS0);
            // P' of region cols [0,32) sits at S cols [0,16), of [32,64) at S cols [32,48)
#pragma unroll
            for (int ks = 0; ks < CH / 16; ++ks)
              if (ks * 16 < n)
                mma_ts(tmem + Cfg::kColC, a_tmem + (ks >> 1) * 32 + (ks & 1) * 8, vbase + ((ks * 2048) >> 4), idesc2, (c > 0) || (ks > 0));
            mma_commit(kv_empty + st);
            if (c == nch - 1) mma_commit(c_full);
            XMC_TRACE(0, x, 1);
            if (g + 2 < G) {
              mbar_wait(kv_full + st2, ph2, wc, 4);               // its latency hides behind G2(g) on the pipe
              tc_fence_after();
              issue_g1(st2, min(CH, p.Rpad - c2 * CH), sb);       // overwrites P(g) after G2(g): same pipe, in order
              mma_commit(s_full + sb);
              advance2(); if (++c2 == nch) c2 = 0;
            }
            XMC_TRACE(0, x, 2);
            if (++st == Cfg::kStages) st = 0;
            if (++c == nch) { c = 0; ++ic; }
          }
          ++seg;
        }
      }
      __syncwarp();
    } else if (warp >= 4 && warp < 12) {
      // ===== softmax warpgroups: thread = TMEM lane = word row; WG h owns chunk cols [32h, 32h+32)
      const int h = (warp - 4) >> 2;
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const uint32_t lane_base = tmem + (static_cast<uint32_t>(q * 32) << 16);
      const float c1 = p.rho1 * kLog2eTc;
      const int col0 = h * 32;
      int st = 0, x = 0, ic = 0;                     // running: rn_s stage, chunk count, image count
      SegIter it = seg_iter(NQ, p.Bi);
      Seg sg;
      while (it.next(sg)) {
        const int grow = sg.tile * TM + row;
        // Q row -> TMEM (A operand: 32-bit column c holds bf16 elements 2c, 2c+1); each WG writes half of D.
        // Every G1 of the previous segment has completed (this thread consumed its S), so Q may be replaced.
        {
          const uint4* src = reinterpret_cast<const uint4*>(p.qn + (size_t)grow * D);
#pragma unroll
          for (int b = 0; b < D / 64; ++b) {                 // 32 bf16 = 16 columns per store
            const int blk = h * (D / 64) + b;
            uint32_t v[16];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint4 w = (grow < NQ) ? __ldg(src + blk * 4 + u) : make_uint4(0, 0, 0, 0);
              v[4 * u + 0] = w.x; v[4 * u + 1] = w.y; v[4 * u + 2] = w.z; v[4 * u + 3] = w.w;
            }
            tmem_st16(lane_base + Cfg::kColQ + blk * 16, v);
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(q_ready);
        }
        for (int ii = 0; ii < sg.nimg; ++ii, ++ic) {
          float l = 0.f, a = 0.f;
          for (int c = 0; c < nch; ++c, ++x) {
            const int n = min(CH, p.Rpad - c * CH);
            const uint32_t s_col = (x & 1) ? Cfg::kColS1 : Cfg::kColS0;
            mbar_wait(s_full + (x & 1), (x >> 1) & 1, wc, 7);
            tc_fence_after();
            if (threadIdx.x == 128) XMC_TRACE(1, x, 0);
            if (col0 < n) {
              uint32_t sv[32];
              tmem_ld32(lane_base + s_col + col0, sv);
              tmem_wait_ld();
              if (threadIdx.x == 128) XMC_TRACE(1, x, 1);
              if (p.dbg && blockIdx.x == 0 && x == 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) p.dbg[row * CH + col0 + j] = __uint_as_float(sv[j]);
              }
              uint32_t pk[16];
              const int r0 = c * CH + col0;
              const float* wsm = rn_s + st * CH + col0;
              // phase 1: all exponentials back to back (packed fp32x2 FMAs feed the MUFU)
              float pv[32];
              {
                const float2 cc = make_float2(c1, c1), ncc = make_float2(-c1, -c1);
#pragma unroll
                for (int j2 = 0; j2 < 16; ++j2) {
                  const float2 arg = __ffma2_rn(make_float2(__uint_as_float(sv[2 * j2]), __uint_as_float(sv[2 * j2 + 1])), cc, ncc);
                  pv[2 * j2] = ex2_approx(arg.x);
                  pv[2 * j2 + 1] = ex2_approx(arg.y);
                }
              }
              if (r0 + 32 > p.R) {
                // ragged tail of the image: padded columns weigh 0.  The scores of columns the MMA never wrote are
                // finite (the S buffers are zeroed at kernel start and only ever hold cosines after that)
#pragma unroll
                for (int j = 0; j < 32; ++j) pv[j] = (r0 + j) < p.R ? pv[j] : 0.f;
              }
              if (threadIdx.x == 128) XMC_TRACE(3, x, 0);
              // phase 2: l += p, p' = p * ||v_r||, a += p' s, pack (packed fp32x2, two chains each)
              float2 l2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
              float2 a2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                float4 mr = make_float4(1.f, 1.f, 1.f, 1.f);
                if (has_rn) mr = *reinterpret_cast<const float4*>(wsm + j4 * 4);   // rows past Rpad of the chunk: pv = 0
                const float2 mr2[2] = {make_float2(mr.x, mr.y), make_float2(mr.z, mr.w)};
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int j = j4 * 4 + 2 * u;
                  const float2 p2 = make_float2(pv[j], pv[j + 1]);
                  const float2 s2 = make_float2(__uint_as_float(sv[j]), __uint_as_float(sv[j + 1]));
                  l2[u] = __fadd2_rn(l2[u], p2);
                  const float2 pw = has_rn ? __fmul2_rn(p2, mr2[u]) : p2;
                  a2[u] = __ffma2_rn(pw, s2, a2[u]);
                  pk[j4 * 2 + u] = pack_bf16(pw.x, pw.y);
                }
              }
              l += (l2[0].x + l2[0].y) + (l2[1].x + l2[1].y);
              a += (a2[0].x + a2[0].y) + (a2[1].x + a2[1].y);
              if (threadIdx.x == 128) XMC_TRACE(3, x, 1);
              tmem_st16(lane_base + s_col + col0, pk);        // WG0 -> S cols [0,16), WG1 -> [32,48)
              tmem_wait_st();
            }
            if (threadIdx.x == 128) XMC_TRACE(1, x, 2);
            tc_fence_before();
            mbar_arrive(p_full + (x & 1));
            if (threadIdx.x == 128) XMC_TRACE(1, x, 3);
            if (++st == Cfg::kStages) st = 0;
          }
          // hand this warpgroup's partial l, a of the image to the epilogue warpgroup
          mbar_wait(la_empty + (ic & 1), ((ic >> 1) & 1) ^ 1, wc, 9);
          float* ex = ex_s + (ic & 1) * (2 * 2 * TM);
          ex[(h * 2 + 0) * TM + row] = l;
          ex[(h * 2 + 1) * TM + row] = a;
          mbar_arrive(la_full + (ic & 1));                    // release semantics: the stores above are visible
        }
      }
    } else if (warp >= 12) {
      // ===== epilogue warpgroup: thread = TMEM lane = word row.  One pass over C per image:
      //       ||C|| -> statistics;  bf16 C -> swizzled smem boxes -> TMA store (saved for the backward)
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const uint32_t lane_base = tmem + (static_cast<uint32_t>(q * 32) << 16);
      const bool issuer = (threadIdx.x == 384);
      int ic = 0;
      long long clk0 = 0, ns0 = 0;                   // trace only: SM cycles vs wall time of this CTA's main loop
      if (p.trace && issuer && blockIdx.x == 0) { clk0 = clock64(); ns0 = global_timer_ns(); }
      SegIter it = seg_iter(NQ, p.Bi);
      Seg sg;
      while (it.next(sg)) {
        const int m0 = sg.tile * TM;
        const int grow = m0 + row;
        for (int ii = 0; ii < sg.nimg; ++ii, ++ic) {
          const int img = sg.img0 + ii * sg.stride;
          mbar_wait(la_full + (ic & 1), (ic >> 1) & 1, wc, 10);
          const float* ex = ex_s + (ic & 1) * (2 * 2 * TM);
          const float l = ex[0 * TM + row] + ex[2 * TM + row];
          const float a = ex[1 * TM + row] + ex[3 * TM + row];
          mbar_arrive(la_empty + (ic & 1));
          const float inv_l = 1.f / l;
          if (p.save_ctx) {
            if (issuer) bulk_wait_read<0>();                  // the previous image's stores have read the boxes
            named_bar_sync(2, 128);
          }
          mbar_wait(c_full, ic & 1, wc, 8);
          tc_fence_after();
          if (issuer) XMC_TRACE(2, ic, 0);
          uint32_t cva[32], cvb[32];
          // Critical section (C blocks the next image's GEMM2): 32 columns -> bf16 -> four 16-byte units of
          // the row's 128-byte box line, nothing else.  The saved tile is the UNSCALED sum C = l * c; the
          // backward folds 1/(l ||c||) into its coefficients.  ||C|| is taken afterwards from the bf16 copy.
          auto consume = [&](const uint32_t (&cv)[32], int blk32) {
            if (p.dbg && blockIdx.x == 0 && ic == 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) p.dbg[TM * CH + row * D + blk32 * 32 + j] = __uint_as_float(cv[j]);
            }
            uint8_t* line = out_s + (blk32 >> 1) * Cfg::kOutBytes + row * 128;
#pragma unroll
            for (int u = 0; u < 4; ++u) {                     // 8 bf16 = one 16-byte unit
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                w[e] = pack_bf16(__uint_as_float(cv[8 * u + 2 * e]), __uint_as_float(cv[8 * u + 2 * e + 1]));
              *reinterpret_cast<uint4*>(line + ((((blk32 & 1) * 4 + u) ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          };
          tmem_ld32(lane_base + Cfg::kColC, cva);
#pragma unroll 1
          for (int b2 = 0; b2 < D / 64; ++b2) {               // two 32-column loads in flight / being consumed
            tmem_wait_ld();
            tmem_ld32(lane_base + Cfg::kColC + b2 * 64 + 32, cvb);
            consume(cva, 2 * b2);
            tmem_wait_ld();
            if (b2 + 1 < D / 64) {
              tmem_ld32(lane_base + Cfg::kColC + b2 * 64 + 64, cva);
            } else {                                          // C is in registers: the next image may overwrite it
              tc_fence_before();
              mbar_arrive(c_empty);
              if (issuer) XMC_TRACE(2, ic, 1);
            }
            consume(cvb, 2 * b2 + 1);
          }
          // off the critical path: ||C||^2 of this thread's row from its own bf16 copy (packed fp32x2)
          float2 c2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll 2
          for (int b = 0; b < Cfg::kOutBoxes; ++b) {
            const uint8_t* line = out_s + b * Cfg::kOutBytes + row * 128;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint4 w = *reinterpret_cast<const uint4*>(line + ((u ^ (row & 7)) << 4));
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 v = make_float2(__uint_as_float(ww[e] << 16), __uint_as_float(ww[e] & 0xffff0000u));
                c2[e & 1] = __ffma2_rn(v, v, c2[e & 1]);
              }
            }
          }
          if (p.save_ctx) {
            fence_proxy_async_smem();
            named_bar_sync(2, 128);
            if (issuer) {
#pragma unroll
              for (int b = 0; b < Cfg::kOutBoxes; ++b) tma_store_3d(&tm_ctx, out_s + b * Cfg::kOutBytes, b * 64, m0, img);
              bulk_commit();
            }
          }
          if (issuer) XMC_TRACE(2, ic, 3);
          if (grow < NQ) {
            const float cn = sqrtf((c2[0].x + c2[0].y) + (c2[1].x + c2[1].y)) * inv_l;
            const size_t o2 = (size_t)img * p.NQ + grow;
            p.lsum[o2] = l;
            p.cnorm[o2] = cn;
            p.rel[o2] = (a * inv_l) / fmaxf(cn, kEps);
          }
        }
      }
      if (issuer) bulk_wait<0>();                             // all context stores performed before exit
      if (p.trace && issuer && blockIdx.x == 0) {
        long long* t = p.trace + (2 * 64 + 63) * 4;
        t[0] = clock64() - clk0; t[1] = global_timer_ns() - ns0; t[2] = ic;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(*tmem_slot, 512);
}

// Debug flags (TMEM dump, NaN poisoning, clock64 traces, A/B switches).  The product library has no way to set
// them: the setter exists only in the -DXMC_TEST_HOOKS build that tests and profiles/exp_*.py load.
#ifdef XMC_TEST_HOOKS
static int g_debug_dump = 0;
#else
static constexpr int g_debug_dump = 0;
#endif

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// kn [Bi, Rpad, D] bf16 as a 3-D tensor (d, r, image); box = 64 d x 64 rows x 1 image, 128B swizzle.
static int make_region_map(CUtensorMap* m, const void* kn, int Bi, int Rpad, int D) {
  PFN_encodeTiled enc = get_encode();
  XMC_REQUIRE(enc != nullptr, XMC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)Rpad, (cuuint64_t)Bi};
  cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)Rpad * D * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)CH, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(kn), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XMC_REQUIRE(r == CUDA_SUCCESS, XMC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return XMC_OK;
}

// 2-D [rows, D] / 3-D [outer, rows, D] bf16 tensor maps, box = 64 d x box_rows (x 1), 128B swizzle
static int make_rows_map(CUtensorMap* m, const void* base, int outer, int rows, int D, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  XMC_REQUIRE(enc != nullptr, XMC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)rows, (cuuint64_t)outer};
  cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)rows * D * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, outer > 0 ? 3 : 2, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XMC_REQUIRE(r == CUDA_SUCCESS, XMC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return XMC_OK;
}

// 2-D [rows, D] fp32 tensor map, box = 32 floats (128 B) x box_rows, 128B swizzle (reduce-add target)
static int make_f32_rows_map(CUtensorMap* m, const void* base, int rows, int D, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  XMC_REQUIRE(enc != nullptr, XMC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)D * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XMC_REQUIRE(r == CUDA_SUCCESS, XMC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return XMC_OK;
}

static int num_sms() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

template <int D>
static int launch_fwd_tc(const WrParams& w, void* ws, size_t ws_bytes, cudaStream_t st) {
  using Cfg = FwdCfg<D>;
  XMC_REQUIRE(ws && ws_bytes >= 64, XMC_ERR_WORKSPACE, "workspace too small (%zu bytes)", ws_bytes);
  CUtensorMap tm, tctx;
  if (int rc = make_region_map(&tm, w.kn, w.Bi, w.Rpad, D)) return rc;
  // without a context buffer the map is never used; point it at the words so that it encodes
  if (int rc = make_rows_map(&tctx, w.chat ? w.chat : w.qn, w.chat ? w.Bi : 1, w.NQ, D, TM)) return rc;
  TcFwdParams p{};
  p.qn = static_cast<const __nv_bfloat16*>(w.qn);
  p.rnorm = w.rnorm; p.NQ = w.NQ; p.Bi = w.Bi; p.R = w.R; p.Rpad = w.Rpad; p.rho1 = w.rho1;
  p.lsum = w.lsum; p.cnorm = w.cnorm; p.rel = w.rel;
  p.save_ctx = w.chat != nullptr;
  p.poison = (g_debug_dump & 16) != 0;
  p.err = static_cast<int*>(ws);
  p.dbg = ((g_debug_dump & 1) && ws_bytes >= 64 + sizeof(float) * (size_t)(TM * CH + TM * D))
              ? reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 64) : nullptr;
  p.trace = ((g_debug_dump & 4) && ws_bytes >= 64 + 4 * 64 * 4 * 8) ? reinterpret_cast<long long*>(static_cast<uint8_t*>(ws) + 64) : nullptr;
  p.nq_dev = w.nq_dev;
  const int grid = num_sms();                                             // persistent: the schedule adapts in the kernel
  XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(wr_fwd_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  wr_fwd_tc_kernel<D><<<grid, kFwdThreads, Cfg::kSmemBytes, st>>>(tm, tctx, p);
  return cuda_fail(cudaGetLastError(), "wr_fwd_tc_kernel launch");
}

size_t wordregion_tc_workspace_bytes(int, int, int, int, int D) {
  return 64 + sizeof(float) * (size_t)(TM * CH + TM * D);   // error word + optional debug dump
}

int wordregion_tc_forward(const WrParams& p, int D, void* ws, size_t ws_bytes, cudaStream_t st) {
  switch (D) {
    case 64: return launch_fwd_tc<64>(p, ws, ws_bytes, st);
    case 128: return launch_fwd_tc<128>(p, ws, ws_bytes, st);
    case 256: return launch_fwd_tc<256>(p, ws, ws_bytes, st);
  }
  set_error("word-region D=%d unsupported (64, 128, 256)", D);
  return XMC_ERR_UNSUPPORTED;
}

// =====================================================================================================
// Backward.  Tile = (128 word rows) x (one image); the CTA keeps its word tile Q in smem and the
// gradient dQ [128 x D] in TMEM across ALL the images it visits; per image the saved context sums
// Chat (C = l c_t, bf16; 1/(l ||c||) is folded into the elementwise coefficients) arrive by TMA, the
// regions in 64-row chunks through a 2-stage ring.  Per chunk g:
//   S = Q Khat^T, W = Chat Khat^T                     (SS MMAs, K-major operands)          -> TMEM
//   X = dS + gamma*alpha', Y = -gamma*rel*alpha'      (2 warpgroups, bf16, 128B-swizzled smem tiles)
//   dQ   += X Khat_chunk                              (A = X K-major, B = Khat MN-major)   -> TMEM, persistent
//   dK^T  = Chat^T Y + Q^T X   [D x 64]               (A = Chat / Q as MN-major, B = Y / X MN-major) -> TMEM
//   dK^T -> registers -> red.global.add.f32 into dkn (a warp adds 128 contiguous bytes per instruction);
//           column sums of alpha*d alpha' -> drnorm
// The same smem bytes serve as K-major and as MN-major operands; nothing of size Bi x Bc x T x R
// ever reaches global memory.
//
// Software pipeline (tensor pipe order):  S,W(g+1) | dQ(g) dK^T(g) | S,W(g+2) | dQ(g+1) dK^T(g+1) ...
//   * S,W(g+1) is issued as soon as the elementwise warps hold S,W(g) in registers (sw_consumed), so it
//     runs under their arithmetic;
//   * the elementwise warps do the arithmetic of chunk g+1 while dQ(g), dK^T(g) execute, then drain dK^T(g);
//   * at an image boundary W(g+1) needs the next image's Chat, which can only be fetched once the last
//     dK^T of the current image has consumed the old one: there the order is S(g+1) | dQ dK^T(g) | W(g+1).
//   * dK^T(g) is drained by a dedicated warpgroup: the adds into dkn run at the L2's atomic rate
//     (~3000 clk per chunk with every SM active, measured) and would otherwise stall the arithmetic.
// Warps: 0 TMA, 1 MMA, 2 TMEM alloc, 4-7 elementwise WG0 (chunk cols 0-31), 8-11 elementwise WG1
// (cols 32-63), 12-15 drain WG (dK^T rows d<128, then d>=128).  Registers are re-partitioned with
// setmaxnreg (the sum must stay within the CTA's launch allocation, 512 x 128): 168 per elementwise
// thread, 96 per drain thread, 80 per role thread.
// =====================================================================================================
constexpr int kBwdThreads = 512;
constexpr int kBwdRegsRole = 80, kBwdRegsEw = 168, kBwdRegsDrain = 96;
static_assert(128 * kBwdRegsRole + 256 * kBwdRegsEw + 128 * kBwdRegsDrain <= 512 * 128, "register pool of the CTA");

template <int D>
struct BwdCfg {
  static constexpr int kQBytes = TM * D * 2;       // Q / Chat tile: D/64 blocks of [128 rows x 128 B]
  static constexpr int kQBlock = TM * 128;
  static constexpr int kKvStage = CH * D * 2;      // region chunk: D/64 blocks of [64 rows x 128 B]
  static constexpr int kKvBlock = CH * 128;
  static constexpr int kXBytes = TM * 128;         // X / Y tile: [128 t x 64 r] bf16
  static constexpr int kOffQ = 0, kOffC = kQBytes, kOffKv = 2 * kQBytes, kOffX = kOffKv + 2 * kKvStage,
                       kOffY = kOffX + kXBytes, kOffBar = kOffY + kXBytes, kOffRn = kOffBar + 256;
  static constexpr int kSmemBytes = kOffRn + 2 * CH * 4 + 1024;
  static constexpr int kTilesD = D / 128;          // M-tiles of dK^T
  static constexpr int kColDQ = 0, kColS = D, kColW = D + CH, kColDK = D + 2 * CH;
  static_assert(kColDK + kTilesD * CH <= 512, "TMEM budget");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

struct TcBwdParams {
  const float* rnorm;
  int NQ, Bi, R, Rpad;
  float rho1;
  const float* lsum; const float* cnorm; const float* rel; const float* grel;
  float* dqn; float* dkn; float* drnorm;
  const int* nq_dev;           // device count of valid word rows (<= NQ) or null
  int* err;
  int dbg_flags;               // xmc_internal_set_debug_dump bits (tests / perf experiments only): 2 = skip the dK reduce,
                               // 4 = clock64 trace, 8 = per-CTA timing log, 16 = fill TMEM with NaNs first (tests),
                               // 64 = hand the context tile over whole, 128 = S waits for the whole region stage
  long long* trace;            // perf experiments only: flag 4 = clock64 timeline of CTA 0, [4 roles][64 chunks][4];
                               // flag 8 = per-CTA {start ns, end ns, segments, images} after it
};

template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

template <int D>
__global__ void __launch_bounds__(kBwdThreads, 1)
wr_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_c,
                 const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_dq, TcBwdParams p) {
  using Cfg = BwdCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem + Cfg::kOffQ;
  uint8_t* Cs = smem + Cfg::kOffC;
  uint8_t* kv = smem + Cfg::kOffKv;
  uint8_t* Xs = smem + Cfg::kOffX;
  uint8_t* Ys = smem + Cfg::kOffY;
  float* rn_s = reinterpret_cast<float*>(smem + Cfg::kOffRn);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 3;    // [2]
  uint64_t* kv_empty = bars + 5;   // [2]
  uint64_t* sw_full = bars + 7;
  uint64_t* sw_consumed = bars + 8;
  uint64_t* xy_full = bars + 9;
  uint64_t* dk_full = bars + 10;
  uint64_t* dk_empty = bars + 11;
  uint64_t* dq_full = bars + 12;   // per segment: every MMA of the segment has executed
  uint64_t* dq_empty = bars + 13;  // per segment: dQ has been read out of TMEM
  // The context tile is handed over in kCH feature halves (one per M-tile of dK^T): at the end of an image the
  // half that dK^T's first M-tile has finished with is refilled while the second M-tile still runs, and the
  // next image's W starts on the first half while the second is in flight.
  constexpr int kCH = Cfg::kTilesD;
  uint64_t* ch_full = bars + 14;   // [kCH]
  uint64_t* ch_empty = bars + 16;  // [kCH]
  // A region stage lands in two feature halves as well (kv_full: features [0, D/2) and the region norms,
  // kv_full_b: features [D/2, D)): S starts on the first half while the second is in flight
  uint64_t* kv_full_b = bars + 18; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  int* abort_flag = reinterpret_cast<int*>(tmem_slot + 1);
  const WaitCtx wc{abort_flag, p.err};

  const int warp = warp_index(), lane = threadIdx.x & 31;
  const int NQ = device_rows(p.nq_dev, p.NQ);
  const int nch = (p.Rpad + CH - 1) / CH;
  const bool has_rn = p.rnorm != nullptr;

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(ch_full + s, 1); mbar_init(ch_empty + s, 1); mbar_init(kv_full_b + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(kv_full + s, 1); mbar_init(kv_empty + s, 1); }
    mbar_init(sw_full, 1); mbar_init(sw_consumed, 256); mbar_init(xy_full, 256);
    mbar_init(dk_full, 1); mbar_init(dk_empty, 128); mbar_init(dq_full, 1); mbar_init(dq_empty, 256);
    fence_barrier_init();
  }
  // region-norm slots past a chunk's rows are never loaded: make them finite (they multiply weights that are 0)
  for (int i = threadIdx.x; i < 2 * CH; i += kBwdThreads) rn_s[i] = 0.f;
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_c); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_dq); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // The allocation is the whole tensor memory (512 columns, one CTA per SM), so its base is lane 0 /
  // column 0; using the literal keeps every TMEM address an immediate for the MMA issue path.
  constexpr uint32_t tmem = 0;
  if (threadIdx.x == 0 && *tmem_slot != 0) { *abort_flag = 1; if (p.err) atomicExch(p.err, 999); }
  if (warp >= 12) {
    if (p.dbg_flags & 16) poison_tmem(warp & 3);
    zero_tmem(warp & 3, Cfg::kColS, 2 * CH);
  }
  __syncthreads();

  // Running counters used by every role (never reset across segments): x = chunk index (region stage
  // x & 1 with parity (x >> 1) & 1; parity x & 1 of the per-chunk barriers), ic = image index (parity
  // of ch_full / ch_empty), seg = segment index (parity of q_full / dq_full / dq_empty).
  //
  // setmaxnreg sits at the top of each role's branch so that it dominates the role's code (ptxas
  // allocates registers per region only then)
  if (warp < 4) {
    setmaxnreg_dec<kBwdRegsRole>();
    if (warp == 0) {
      // ===== TMA producer =====
      if (elect_one()) {
        int x = 0, ic = 0, seg = 0;
        SegIter it = seg_iter(NQ, p.Bi);
        Seg sg;
        while (it.next(sg)) {
          const int m0 = sg.tile * TM;
          if (seg > 0) mbar_wait(dq_full, (seg - 1) & 1, wc, 10);   // no MMA reads the old word tile any more
          mbar_expect_tx(q_full, Cfg::kQBytes);
#pragma unroll
          for (int kb = 0; kb < D / 64; ++kb) tma_load_2d(Qs + kb * Cfg::kQBlock, &tm_q, kb * 64, m0, q_full);
          for (int ii = 0; ii < sg.nimg; ++ii, ++ic) {
            const int img = sg.img0 + ii * sg.stride;
            for (int c = 0; c < nch; ++c, ++x) {
              const int st = x & 1;
              const int n = min(CH, p.Rpad - c * CH);
              mbar_wait(kv_empty + st, ((x >> 1) & 1) ^ 1, wc, 11);
              mbar_expect_tx(kv_full + st, Cfg::kKvStage / 2 + (has_rn ? n * 4 : 0));
              mbar_expect_tx(kv_full_b + st, Cfg::kKvStage / 2);
#pragma unroll
              for (int kb = 0; kb < D / 64; ++kb)
                tma_load_3d(kv + st * Cfg::kKvStage + kb * Cfg::kKvBlock, &tm_k, kb * 64, c * CH, img,
                            (kb < D / 128 ? kv_full : kv_full_b) + st);
              if (has_rn) bulk_load_1d(rn_s + st * CH, p.rnorm + (size_t)img * p.Rpad + c * CH, n * 4, kv_full + st);
              if (c == 0) {
                if (p.dbg_flags & 64)                      // A/B timing only: hand the tile over whole
                  for (int hh = 0; hh < kCH; ++hh) mbar_wait(ch_empty + hh, (ic & 1) ^ 1, wc, 12);
#pragma unroll
                for (int hh = 0; hh < kCH; ++hh) {
                  mbar_wait(ch_empty + hh, (ic & 1) ^ 1, wc, 12);
                  mbar_expect_tx(ch_full + hh, Cfg::kQBytes / kCH);
#pragma unroll
                  for (int kb = hh * (D / 64 / kCH); kb < (hh + 1) * (D / 64 / kCH); ++kb)
                    tma_load_3d(Cs + kb * Cfg::kQBlock, &tm_c, kb * 64, m0, img, ch_full + hh);
                }
              }
              // the next image's contexts come from HBM and are needed the moment this image ends (the single
              // buffer cannot be refilled earlier): pull them into L2 shortly before
              if (c == max(0, nch - 3) && ii + 1 < sg.nimg) {
#pragma unroll
                for (int kb = 0; kb < D / 64; ++kb) tma_prefetch_l2_3d(&tm_c, kb * 64, m0, img + sg.stride);
              }
            }
          }
          ++seg;
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ===== MMA issuer: one elected thread runs the whole role =====
      if (elect_one()) {
        constexpr uint32_t idesc_dq = idesc_bf16(TM, D, false, true);
        // base descriptors; k-steps / blocks / stages are reached by adding (byte offset >> 4)
        const Desc q_k = make_desc(smem_u32(Qs), 16, 1024), q_mn = make_desc(smem_u32(Qs), Cfg::kQBlock, 1024);
        const Desc c_k = make_desc(smem_u32(Cs), 16, 1024), c_mn = make_desc(smem_u32(Cs), Cfg::kQBlock, 1024);
        const Desc kv_k = make_desc(smem_u32(kv), 16, 1024), kv_mn = make_desc(smem_u32(kv), Cfg::kKvBlock, 1024);
        const Desc x_k = make_desc(smem_u32(Xs), 16, 1024), x_mn = make_desc(smem_u32(Xs), Cfg::kXBytes, 1024);
        const Desc y_mn = make_desc(smem_u32(Ys), Cfg::kXBytes, 1024);
        auto issue_scores_part = [&](int st, int n, Desc a_k, uint32_t d_col, int part, int parts) {
          const uint32_t idesc = idesc_bf16(TM, n, false, false);   // [128 x n] = A[128 x D] . Khat_chunk^T, k-steps of one part
          const Desc b_k = kv_k + ((uint32_t)(st * Cfg::kKvStage) >> 4);
          const int k0 = part * (D / 16 / parts);
#pragma unroll
          for (int kk = 0; kk < D / 16 / parts; ++kk) {
            const int k = k0 + kk;
            mma_ss(tmem + d_col, a_k + (((k >> 2) * Cfg::kQBlock + (k & 3) * 32) >> 4),
                   b_k + (((k >> 2) * Cfg::kKvBlock + (k & 3) * 32) >> 4), idesc, k > 0);
          }
        };
        auto issue_scores = [&](int st, int n, Desc a_k, uint32_t d_col) { issue_scores_part(st, n, a_k, d_col, 0, 1); };
        // S of a chunk whose region stage may still be landing: each feature half as it arrives
        auto issue_s_arriving = [&](int st, int n, int parity) {
          const bool whole = (p.dbg_flags & 128) != 0;     // A/B timing only: wait for the whole stage first
          if (whole) mbar_wait(kv_full_b + st, parity, wc, 18);
          mbar_wait(kv_full + st, parity, wc, 18);
          tc_fence_after();
          issue_scores_part(st, n, q_k, Cfg::kColS, 0, 2);
          mbar_wait(kv_full_b + st, parity, wc, 18);
          tc_fence_after();
          issue_scores_part(st, n, q_k, Cfg::kColS, 1, 2);
        };
        // W of the FIRST chunk of an image: each feature half of the context tile as it lands
        auto issue_w_arriving = [&](int st, int n, int parity) {
          if (p.dbg_flags & 64)                            // A/B timing only
            for (int hh = 0; hh < kCH; ++hh) mbar_wait(ch_full + hh, parity, wc, 15);
#pragma unroll
          for (int hh = 0; hh < kCH; ++hh) {
            mbar_wait(ch_full + hh, parity, wc, 15);
            tc_fence_after();
            issue_scores_part(st, n, c_k, Cfg::kColW, hh, kCH);
          }
        };
        int x = 0, ic = 0, seg = 0;
        SegIter it = seg_iter(NQ, p.Bi);
        Seg sg;
        while (it.next(sg)) {
          const int G = sg.nimg * nch;
          mbar_wait(q_full, seg & 1, wc, 13);
          issue_s_arriving(x & 1, min(CH, p.Rpad), (x >> 1) & 1);
          issue_w_arriving(x & 1, min(CH, p.Rpad), ic & 1);
          mma_commit(sw_full);
          int c = 0;
          for (int g = 0; g < G; ++g, ++x) {
            const int st = x & 1, n = min(CH, p.Rpad - c * CH);
            const bool last_chunk = (c == nch - 1);
            const bool has_next = g + 1 < G;
            const int n1 = last_chunk ? min(CH, p.Rpad) : min(CH, p.Rpad - (c + 1) * CH);
            if (has_next) {
              // S,W(g) are in the elementwise warps' registers: the next scores may overwrite them now
              mbar_wait(sw_consumed, x & 1, wc, 16);
              XMC_TRACE(0, x, 0);
              issue_s_arriving(st ^ 1, n1, ((x + 1) >> 1) & 1);
              if (!last_chunk) {
                issue_scores(st ^ 1, n1, c_k, Cfg::kColW);
                mma_commit(sw_full);
              }
            }
            mbar_wait(xy_full, x & 1, wc, 17);
            if (x > 0) mbar_wait(dk_empty, (x - 1) & 1, wc, 19);
            if (g == 0 && seg > 0) mbar_wait(dq_empty, (seg - 1) & 1, wc, 9);   // the previous tile's dQ has left TMEM
            tc_fence_after();
            XMC_TRACE(0, x, 1);
            const uint32_t idesc_dk = idesc_bf16(TM, n, true, true);
            auto issue_dq = [&]() {                       // dQ += X Khat_chunk
              const Desc b_mn = kv_mn + ((uint32_t)(st * Cfg::kKvStage) >> 4);
#pragma unroll
              for (int ks = 0; ks < CH / 16; ++ks)
                if (ks * 16 < n) mma_ss(tmem + Cfg::kColDQ, x_k + ((ks * 32) >> 4), b_mn + ((ks * 2048) >> 4), idesc_dq, (g > 0) || (ks > 0));
              mma_commit(kv_empty + st);
            };
            auto issue_dk_c = [&]() {                     // dK^T = Chat^T Y   [128 d x n] per M-tile, contraction over the 128 word rows
#pragma unroll
              for (int h = 0; h < Cfg::kTilesD; ++h) {
#pragma unroll
                for (int kt = 0; kt < TM / 16; ++kt)
                  mma_ss(tmem + Cfg::kColDK + h * CH, c_mn + ((2 * h * Cfg::kQBlock + kt * 2048) >> 4), y_mn + ((kt * 2048) >> 4), idesc_dk, kt > 0);
                if (last_chunk) mma_commit(ch_empty + h);   // this feature half of the context tile is free
              }
            };
            if (last_chunk) {
              // end of the image: release the context buffer first, the refill overlaps dQ and the Q part of dK^T
              issue_dk_c();
              issue_dq();
            } else {
              // inside the image: release the region stage first, the next chunk's load overlaps dK^T
              issue_dq();
              issue_dk_c();
            }
#pragma unroll
            for (int h = 0; h < Cfg::kTilesD; ++h) {      // dK^T += Q^T X
#pragma unroll
              for (int kt = 0; kt < TM / 16; ++kt)
                mma_ss(tmem + Cfg::kColDK + h * CH, q_mn + ((2 * h * Cfg::kQBlock + kt * 2048) >> 4), x_mn + ((kt * 2048) >> 4), idesc_dk, true);
            }
            mma_commit(dk_full);          // X / Y are dead once this fires
            XMC_TRACE(0, x, 2);
            if (has_next && last_chunk) {
              issue_w_arriving(st ^ 1, n1, (ic + 1) & 1);
              mma_commit(sw_full);
            }
            XMC_TRACE(0, x, 3);
            if (++c == nch) { c = 0; ++ic; }
          }
          mma_commit(dq_full);
          ++seg;
        }
      }
      __syncwarp();
    }
  } else if (warp < 12) {
    setmaxnreg_inc<kBwdRegsEw>();
    {
      // ===== elementwise warpgroups: thread = TMEM lane = word row =====
      const int h = (warp - 4) >> 2;              // 0: chunk cols 0-31, 1: cols 32-63
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const uint32_t lane_base = tmem + (static_cast<uint32_t>(q * 32) << 16);
      const float c1 = p.rho1 * kLog2eTc;
      const int col0 = h * 32;
      const bool tracer = (threadIdx.x == 128);

      int xm = 0;                                 // running index of the chunk whose arithmetic is next
      int x = 0, seg = 0;
      SegIter it = seg_iter(NQ, p.Bi);
      Seg sg;
      long long* cta_log = (p.dbg_flags & 8) && tracer ? p.trace + 4 * 64 * 4 + 4 * (long long)blockIdx.x : nullptr;
      if (cta_log) { cta_log[0] = (long long)global_ns(); cta_log[2] = 0; cta_log[3] = 0; }
      while (it.next(sg)) {
        const int grow = sg.tile * TM + row;
        const int G = sg.nimg * nch;
        if (cta_log) { cta_log[2] += 1; cta_log[3] += sg.nimg; }
        // ---- arithmetic of one chunk: S,W -> X,Y (packed bf16, registers) and the drnorm column sum ----
        int cm = 0, iim = 0;                      // chunk-in-image / image-in-segment of the arithmetic chunk
        float inv_l = 1.f, gam = 0.f, ngrl = 0.f;
        uint32_t xp[16], yp[16];
        float z[32];                              // alpha * d alpha' of the chunk: column-summed into drnorm after X, Y left
        // per-(image, word row) statistics of the NEXT image are fetched one image ahead: at the first chunk of
        // an image the elementwise warps are on the critical path (the pipeline drains at image boundaries)
        float pf_l = 1.f, pf_cn = 1.f, pf_g = 0.f, pf_rel = 0.f;
        auto fetch_stats = [&](int ii_) {
          if (grow < NQ && ii_ < sg.nimg) {
            const size_t o = (size_t)(sg.img0 + ii_ * sg.stride) * p.NQ + grow;
            pf_l = __ldg(p.lsum + o); pf_cn = __ldg(p.cnorm + o); pf_g = __ldg(p.grel + o); pf_rel = __ldg(p.rel + o);
          }
        };
        fetch_stats(0);
        auto arithmetic = [&]() {
          if (cm == 0) {
            inv_l = 1.f; gam = 0.f; ngrl = 0.f;
            if (grow < NQ) {
              const float inv_cn = 1.f / fmaxf(pf_cn, kEps);
              inv_l = 1.f / pf_l;
              gam = pf_g * inv_cn;
              // the saved context is the unscaled sum C = l c (not the unit vector): W = C K^T and the
              // Chat^T Y term both carry 1 / (l ||c||), folded into the coefficient that multiplies them
              ngrl = -gam * pf_rel * inv_cn * inv_l;
            }
            fetch_stats(iim + 1);
          }
          const int n = min(CH, p.Rpad - cm * CH);
          const int r0 = cm * CH + col0;
          mbar_wait(sw_full, xm & 1, wc, 20);
          tc_fence_after();
          if (tracer) XMC_TRACE(1, xm, 0);
          if (col0 < n) {                               // warp-uniform: this warpgroup has columns in the chunk
            const float* wsm = rn_s + (xm & 1) * CH + col0;
            // S and W of the warpgroup's 32 columns -> registers, then they are released at once (the next scores
            // may overwrite them while the arithmetic below runs); the arithmetic goes in two halves of 16 columns
            uint32_t svv[2][16], wvv[2][16];
            tmem_ld16(lane_base + Cfg::kColS + col0, svv[0]);
            tmem_ld16(lane_base + Cfg::kColW + col0, wvv[0]);
            tmem_ld16(lane_base + Cfg::kColS + col0 + 16, svv[1]);
            tmem_ld16(lane_base + Cfg::kColW + col0 + 16, wvv[1]);
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(sw_consumed);
            if (tracer) XMC_TRACE(1, xm, 1);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const uint32_t (&sv)[16] = svv[hf];
              const uint32_t (&wv)[16] = wvv[hf];
              auto elementwise = [&](auto full_tag) {
                constexpr bool kFull = decltype(full_tag)::value;   // every column is a real region: no predicates
                const float2 cc = make_float2(c1, c1), ncc = make_float2(-c1, -c1);
                const float2 il2 = make_float2(inv_l, inv_l), gam2 = make_float2(gam, gam);
                const float2 ng2 = make_float2(ngrl, ngrl), rho2v = make_float2(p.rho1, p.rho1);
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                  float4 mr = make_float4(1.f, 1.f, 1.f, 1.f);
                  if (has_rn) mr = *reinterpret_cast<const float4*>(wsm + hf * 16 + j4 * 4);
                  const float2 mr2[2] = {make_float2(mr.x, mr.y), make_float2(mr.z, mr.w)};
#pragma unroll
                  for (int u = 0; u < 2; ++u) {               // packed fp32x2 arithmetic on column pairs
                    const int j = j4 * 4 + 2 * u;
                    const float2 s2 = make_float2(__uint_as_float(sv[j]), __uint_as_float(sv[j + 1]));
                    const float2 w2 = make_float2(__uint_as_float(wv[j]), __uint_as_float(wv[j + 1]));
                    const float2 arg = __ffma2_rn(s2, cc, ncc);
                    float2 al = __fmul2_rn(make_float2(ex2_approx(arg.x), ex2_approx(arg.y)), il2);   // alpha
                    // padded region rows: alpha = 0 zeroes X, Y, z.  s, w of columns the MMAs never wrote are
                    // finite: the score columns are zeroed at kernel start and only ever hold cosines after that
                    if (!kFull) {
                      al.x = (r0 + hf * 16 + j) < p.R ? al.x : 0.f;
                      al.y = (r0 + hf * 16 + j + 1) < p.R ? al.y : 0.f;
                    }
                    const float2 alp = __fmul2_rn(al, mr2[u]);                             // alpha' = alpha * ||v_r||
                    const float2 dap = __ffma2_rn(ng2, w2, __fmul2_rn(gam2, s2));          // d loss / d alpha'
                    const float2 xv = __fmul2_rn(alp, __ffma2_rn(rho2v, dap, gam2));       // dS + gamma*alpha'
                    const float2 yv = __fmul2_rn(ng2, alp);
                    const float2 zz = __fmul2_rn(al, dap);
                    z[hf * 16 + j] = zz.x; z[hf * 16 + j + 1] = zz.y;
                    xp[hf * 8 + j4 * 2 + u] = pack_bf16(xv.x, xv.y);
                    yp[hf * 8 + j4 * 2 + u] = pack_bf16(yv.x, yv.y);
                  }
                }
              };
              // With raw values the padded region rows carry norm 0 (and never-loaded norm slots / never-written
              // score columns are zeroed at kernel start): alpha' = alpha * 0 and s = w = 0 already make X, Y, z
              // exact zeros there, so the predicate-free path serves every chunk.
              if (has_rn || r0 + 32 <= p.R) elementwise(std::true_type{}); else elementwise(std::false_type{});
            }
          } else {
            tc_fence_before();
            mbar_arrive(sw_consumed);
          }
          if (tracer) XMC_TRACE(1, xm, 2);
          ++xm;
          if (++cm == nch) { cm = 0; ++iim; }
        };

        // One call site for the arithmetic (the code of this role is large: every copy costs instruction cache):
        // iteration g = -1 only does the arithmetic of chunk 0.
        int c = 0, ii = 0;
        for (int g = -1; g < G; ++g) {
          if (g >= 0) {
            const int img = sg.img0 + ii * sg.stride;
            const int n = min(CH, p.Rpad - c * CH);
            const bool active = col0 < n;
            const int r0 = c * CH + col0;
            // ---- X,Y(g) -> smem.  Safe: dk_full of the previous chunk was waited for below, so the MMAs that read X,Y are done.
            if (active) {
              // row `row` of the [128 x 64] bf16 tiles, 16-byte chunks XOR-swizzled by (row & 7)
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int phys = ((col0 >> 3) + u) ^ (row & 7);
                *reinterpret_cast<uint4*>(Xs + row * 128 + phys * 16) = make_uint4(xp[4 * u], xp[4 * u + 1], xp[4 * u + 2], xp[4 * u + 3]);
                *reinterpret_cast<uint4*>(Ys + row * 128 + phys * 16) = make_uint4(yp[4 * u], yp[4 * u + 1], yp[4 * u + 2], yp[4 * u + 3]);
              }
            }
            fence_proxy_async_smem();
            mbar_arrive(xy_full);
            if (tracer) XMC_TRACE(2, x, 0);
            // off the critical path (X, Y are already with the tensor pipe): column sums of z -> drnorm
            if (active && has_rn) {
              const float colsum = warp_transpose_sum32(z, lane);      // column (r0 + lane) over this warp's 32 rows
              if (r0 + lane < p.R) atomicAdd(p.drnorm + (size_t)img * p.Rpad + r0 + lane, colsum);
            }
          }
          // ---- arithmetic of chunk g+1 while dQ(g), dK^T(g) run on the tensor pipe ----
          if (g + 1 < G) arithmetic();
          if (g >= 0) {
            if (tracer) XMC_TRACE(2, x, 1);
            // ---- X,Y(g) stay live until dQ(g), dK^T(g) have executed ----
            mbar_wait(dk_full, x & 1, wc, 21);
            if (tracer) XMC_TRACE(2, x, 2);
            if (++c == nch) { c = 0; ++ii; }
            ++x;
          }
        }
        // ---- dQ of this segment's word tile (summed over its images): TMEM -> swizzled fp32 boxes in the
        //      (now dead) X|Y bytes -> TMA reduce-add into dqn; warpgroup h owns columns [128h, 128h+128) ----
        mbar_wait(dq_full, seg & 1, wc, 22);
        tc_fence_after();
        {
          uint8_t* box = Xs + h * Cfg::kXBytes;         // [128 rows x 32 fp32] = 16 KB, 128B-swizzled rows
          const bool issuer = (q == 0 && lane == 0);
#pragma unroll 1
          for (int blk = 0; blk < D / 64; ++blk) {
            uint32_t dv[32];
            tmem_ld32(lane_base + Cfg::kColDQ + h * (D / 2) + blk * 32, dv);   // warp-collective: never predicate
            tmem_wait_ld();
            if (blk == D / 64 - 1) {                    // dQ is in registers: the next segment may overwrite it
              tc_fence_before();
              mbar_arrive(dq_empty);
            }
            if (blk > 0) {                              // the previous reduce has read the box
              if (issuer) bulk_wait_read<0>();
              named_bar_sync(2 + h, 128);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
              *reinterpret_cast<uint4*>(box + row * 128 + ((u ^ (row & 7)) << 4)) = make_uint4(dv[4 * u], dv[4 * u + 1], dv[4 * u + 2], dv[4 * u + 3]);
            fence_proxy_async_smem();
            named_bar_sync(2 + h, 128);
            if (issuer) {
              tma_reduce_add_2d(&tm_dq, box, h * (D / 2) + blk * 32, sg.tile * TM);   // rows past the buffer are clipped
              bulk_commit();
            }
          }
          if (issuer) bulk_wait_read<0>();              // X|Y are rewritten by the next segment
          named_bar_sync(2 + h, 128);
        }
        ++seg;
      }
      if (q == 0 && lane == 0) bulk_wait<0>();          // all dQ reductions performed before exit
      if (cta_log) cta_log[1] = (long long)global_ns();
    }
  } else {
    setmaxnreg_dec<kBwdRegsDrain>();
    {
      // ===== drain warpgroup: dK^T(g) TMEM -> registers -> red.global.add.f32 into dkn.  Thread = TMEM
      //       lane = feature d of an M-tile; a warp adds 32 consecutive d of one region row (128 bytes). =====
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const uint32_t lane_base = tmem + (static_cast<uint32_t>(q * 32) << 16);
      const bool tracer = (threadIdx.x == 384);
      int x = 0;
      SegIter it = seg_iter(NQ, p.Bi);
      Seg sg;
      while (it.next(sg)) {
        for (int ii = 0; ii < sg.nimg; ++ii) {
          for (int c = 0; c < nch; ++c, ++x) {
            const int n = min(CH, p.R - c * CH);            // real regions of the chunk: padding rows get no add
            mbar_wait(dk_full, x & 1, wc, 23);
            tc_fence_after();
            if (tracer) XMC_TRACE(3, x, 0);
#pragma unroll
            for (int h = 0; h < Cfg::kTilesD; ++h) {
              float* dst0 = p.dkn + ((size_t)(sg.img0 + ii * sg.stride) * p.Rpad + c * CH) * D + h * 128 + row;
              uint32_t dva[32], dvb[32];
              tmem_ld32(lane_base + Cfg::kColDK + h * CH, dva);            // rows of the chunk past n hold stale data:
              if (n > 32) tmem_ld32(lane_base + Cfg::kColDK + h * CH + 32, dvb);   // never added below
              tmem_wait_ld();
              if (h == Cfg::kTilesD - 1) {                                 // dK^T(g) is in registers
                tc_fence_before();
                mbar_arrive(dk_empty);
                if (tracer) XMC_TRACE(3, x, 1);
              }
              if (!(p.dbg_flags & 2)) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < n) red_add_f32(dst0 + (size_t)j * D, __uint_as_float(dva[j]));
                if (n > 32) {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (32 + j < n) red_add_f32(dst0 + (size_t)(32 + j) * D, __uint_as_float(dvb[j]));
                }
              }
            }
            if (tracer) XMC_TRACE(3, x, 2);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(*tmem_slot, 512);
}

template <int D>
static int launch_bwd_tc(const WrParams& w, void* ws, size_t ws_bytes, cudaStream_t st) {
  using Cfg = BwdCfg<D>;
  XMC_REQUIRE(ws && ws_bytes >= 64, XMC_ERR_WORKSPACE, "workspace too small (%zu bytes)", ws_bytes);
  XMC_REQUIRE(w.chat != nullptr, XMC_ERR_INVALID_ARG, "the tcgen05 backward needs the contexts saved by the forward (chat)");
  CUtensorMap tq, tcm, tk, tdq;
  if (int rc = make_f32_rows_map(&tdq, w.dqn, w.NQ, D, TM)) return rc;
  if (int rc = make_rows_map(&tq, w.qn, 0, w.NQ, D, TM)) return rc;
  if (int rc = make_rows_map(&tcm, w.chat, w.Bi, w.NQ, D, TM)) return rc;
  if (int rc = make_region_map(&tk, w.kn, w.Bi, w.Rpad, D)) return rc;
  TcBwdParams p{};
  p.rnorm = w.rnorm; p.NQ = w.NQ; p.Bi = w.Bi; p.R = w.R; p.Rpad = w.Rpad; p.rho1 = w.rho1;
  p.lsum = w.lsum; p.cnorm = w.cnorm; p.rel = w.rel; p.grel = w.grel;
  p.dqn = w.dqn; p.dkn = w.dkn; p.drnorm = w.drnorm;
  p.err = static_cast<int*>(ws);
  p.dbg_flags = g_debug_dump;
  p.trace = ((g_debug_dump & 12) && ws_bytes >= 64 + (4 * 64 * 4 + 4 * 160) * 8) ? reinterpret_cast<long long*>(static_cast<uint8_t*>(ws) + 64) : nullptr;
  if (!p.trace) p.dbg_flags &= ~12;
  p.nq_dev = w.nq_dev;
  const int grid = num_sms();                                             // persistent: the schedule adapts in the kernel
  XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(wr_bwd_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  wr_bwd_tc_kernel<D><<<grid, kBwdThreads, Cfg::kSmemBytes, st>>>(tq, tcm, tk, tdq, p);
  return cuda_fail(cudaGetLastError(), "wr_bwd_tc_kernel launch");
}

int wordregion_tc_backward(const WrParams& p, int D, void* ws, size_t ws_bytes, cudaStream_t st) {
  switch (D) {
    case 128: return launch_bwd_tc<128>(p, ws, ws_bytes, st);
    case 256: return launch_bwd_tc<256>(p, ws, ws_bytes, st);
  }
  set_error("tcgen05 word-region backward supports D = 128, 256 (got %d)", D);
  return XMC_ERR_UNSUPPORTED;
}

}  // namespace xmc

#ifdef XMC_TEST_HOOKS
// Not part of the public ABI (not in include/xmc_loss.h, not in libxmcloss.so): test hook for the debug flags.
extern "C" void xmc_internal_set_debug_dump(int on) { xmc::g_debug_dump = on; }
#endif
