// region_head.cu — the producer of the word loss's region operand, fused into the loss prologue.
//
// The word loss attends over region features that a 1x1 convolution ("region head") projects out of the
// discriminator's 16 x 16 stage (xmc_gan/model/df_gan.py:106-132 produces the [B, 512, 16, 16] map; the head is the
// region-side counterpart of proj_match, df_gan.py:143-145, 165-168).  Unfused, that is a convolution writing
// y[B, D, H, W], a layout kernel re-reading it and writing unit rows — the region map crosses HBM three times.  Here
//
//   forward   y_b = feat_b^T W^T + bias   (M = pixels, N = D = 256, K = Cin)  ->  epilogue: ||y_r||, unit rows
//             kn[B, Rpad, D] bf16 + rnorm[B, Rpad] fp32, exactly what wr_fwd_tc_kernel / wr_bwd_tc_kernel consume;
//             y itself never reaches HBM.
//   backward  dy[B, R, D] (in the map's dtype, from xmc_normalize_rows_backward on the word-region kernels' dkn / drnorm):
//             dfeat_b = W^T dy_b^T        (M = Cin, N = pixels, K = D)        written in feat's own [B, Cin, H, W] layout
//             dW      = sum_b dy_b^T feat_b^T (M = D, N = Cin, K = B * R)     split over CTAs, fp32 red into dW
//             dbias   = sum_{b,r} dy      a column-sum kernel beside it
//
// All three products have ONE operand form on the tensor cores: C[m, n] = sum_k A[k, m] * B[n, k] with A stored
// [K rows][M contiguous] (MN-major smem tile) and B stored [N rows][K contiguous] (K-major tile) — feat_b is
// [Cin][pixels], W is [D][Cin], dy_b is [pixels][D]: every operand is consumed where it lies, no transposes.
// Staging (fast path: both operands of one dtype, rows 16-byte aligned): one thread issues TMA boxes ([k rows x 128 bytes],
// 3-D maps built per call) straight into the swizzled tiles of a 4-stage ring, 192 KB per SM in flight.  bf16 operands
// run as kind::f16 MMAs (64 k per stage); fp32 operands stay fp32 in shared memory and run as kind::tf32 MMAs (32 k per
// stage, 10-bit mantissa: inside the bf16 tolerance; the MN-major fp32 tile needs the 128B_BASE32B / 128B_ATOM_32B
// swizzle), so an fp32 discriminator needs neither a cast pass nor a conversion in registers.  (Two earlier versions
// staged through the LSU — conversion in registers: 100 us per launch, bound by one load round trip per group of chunks;
// 16-byte cp.async: 40-60 us, bound by the ~48 KB of requests an SM keeps outstanding on that path.)  Generic path
// (mixed dtypes, unaligned rows such as a 17 x 17 map): global -> registers -> bf16 -> swizzled smem by eight staging
// warps.  fp32 accumulation in TMEM, two 256-column accumulators: the epilogue of one tile (eight warps: lane quadrant x
// column half) runs under the MMAs of the next and stores whole 64-byte row segments through a per-warp staging block.
//
// Roofline: HBM by bytes (forward at B = 256, 16 x 16, Cin = 512, D = 256: 134 MB (fp32 map) or 67 MB (bf16) in + 33.6 MB
// out against 17.2 GFLOP); measured bound: shared-memory bandwidth — with K = Cin = 512 every staged byte is written once
// (TMA) and read once (MMA operand), 192 B/clk against the SM's 128 B/clk while the MMAs run.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace xmc {

using namespace tc;

namespace {

constexpr int kM = 128, kN = 256, kStages = 4;
constexpr int kABytes = 128 * 128;                 // A tile: 128 m x (64 bf16 | 32 fp32) k = 16 KB
constexpr int kBBytes = kN * 128;                  // B tile: 256 n rows x 128 bytes of k = 32 KB
constexpr int kStageBytes = kABytes + kBBytes;     // 48 KB
constexpr int kEpiWarps = 8, kMmaWarp = 8, kStageWarp0 = 9, kStageWarps = 8;
constexpr int kStageThreads = kStageWarps * 32;    // 256
constexpr int kThreads = (kStageWarp0 + kStageWarps) * 32;   // 544
constexpr int kAccCols = 256;

enum Mode { kFwd = 0, kDFeat = 1, kDW = 2 };

struct HeadParams {
  const void* a; long long a_batch; int lda; int a_bf16;   // A_b[k][m], m contiguous
  const void* b; long long b_batch; int ldb; int b_bf16;   // B_b[n][k], k contiguous
  int M, N, K;                 // valid extents of one batch item's product
  int m_tiles, n_tiles, batches, splits;
  // forward
  const float* bias; __nv_bfloat16* kn; float* rnorm; int R, Rpad;
  // dfeat
  void* dfeat; int out_bf16;   // [B, M = Cin, N = R]
  // dW
  float* dw; float* dbias;     // [M = D, N = Cin] fp32, [D]
  int* err;
};

struct HeadShared {
  alignas(16) uint32_t stg[kEpiWarps][32 * kStgPitch];    // per epilogue warp: a [32 rows x 16 words] block on its way out
  float ss_part[2][2][kM];                                // forward: partial sums of squares [accumulator][column half][row]
  uint64_t full[kStages], empty[kStages], acc_full[2], acc_empty[2];
  uint32_t tmem_slot;
  int abort_flag;
  float bias[kN];
};

// one work item = one accumulator tile: (m0, n0) and the batch items [b0, b1) it sums over
struct Item { int m0, n0, b0, b1, out_b; };

template <int MODE>
__device__ __forceinline__ Item decode_item(const HeadParams& p, int item) {
  Item it;
  if (MODE == kDW) {
    const int s = item % p.splits, t = item / p.splits;
    const int per = (p.batches + p.splits - 1) / p.splits;
    it.m0 = (t / p.n_tiles) * kM; it.n0 = (t % p.n_tiles) * kN;
    it.b0 = s * per; it.b1 = min(p.batches, it.b0 + per); it.out_b = 0;
  } else {
    const int nt = item % p.n_tiles, t = item / p.n_tiles;
    it.m0 = (t % p.m_tiles) * kM; it.n0 = nt * kN;
    it.b0 = t / p.m_tiles; it.b1 = it.b0 + 1; it.out_b = it.b0;
  }
  return it;
}

// Four 8-element chunks of one operand, as they come out of global memory: fp32 -> two uint4 per chunk, bf16 -> one.
struct Raw { uint4 v[8]; };

// 8 elements src[0..8) of a row, `valid` (0..8) of them inside the matrix, rest zero
__device__ __forceinline__ void load_chunk(const void* base, long long off, int valid, bool bf16, bool vec_ok, uint4& r0, uint4& r1) {
  if (bf16) {
    const __nv_bfloat16* s = static_cast<const __nv_bfloat16*>(base) + off;
    if (valid >= 8 && vec_ok) { r0 = __ldg(reinterpret_cast<const uint4*>(s)); return; }
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (e < valid) w[e >> 1] |= static_cast<uint32_t>(__bfloat16_as_ushort(s[e])) << ((e & 1) * 16);
    r0 = make_uint4(w[0], w[1], w[2], w[3]);
  } else {
    const float* s = static_cast<const float*>(base) + off;
    if (valid >= 8 && vec_ok) {
      r0 = __ldg(reinterpret_cast<const uint4*>(s)); r1 = __ldg(reinterpret_cast<const uint4*>(s) + 1);
      return;
    }
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = e < valid ? __ldg(s + e) : 0.f;
    r0 = make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3]));
    r1 = make_uint4(__float_as_uint(x[4]), __float_as_uint(x[5]), __float_as_uint(x[6]), __float_as_uint(x[7]));
  }
}
__device__ __forceinline__ uint4 to_bf16x8(bool bf16, const uint4& r0, const uint4& r1) {
  if (bf16) return r0;
  return make_uint4(pack_bf16(__uint_as_float(r0.x), __uint_as_float(r0.y)), pack_bf16(__uint_as_float(r0.z), __uint_as_float(r0.w)),
                    pack_bf16(__uint_as_float(r1.x), __uint_as_float(r1.y)), pack_bf16(__uint_as_float(r1.z), __uint_as_float(r1.w)));
}
__device__ __forceinline__ void st_chunk(uint8_t* blk, int r, int c, uint4 v) {
  *reinterpret_cast<uint4*>(blk + r * 128 + ((c ^ (r & 7)) << 4)) = v;
}

// FAST: cp.async staging (operands of one dtype, aligned); TF32: fp32 operands as kind::tf32 (FAST only)
// CL = 2 (FAST only; instantiated in the hooks build, see launch_head): CTA pairs (thread-block clusters of two, cta_group::2) on two items that share their B operand
// (the weights in the forward, the image's dy tile in dfeat).  One tcgen05.mma of the pair's leader multiplies a 256-row
// tile over both SMs: each CTA stages its own 128 A rows and HALF of the B rows, so a k-block brings 32 KB into an SM
// instead of 48 KB.  The leader's full barrier collects the bytes of both CTAs, its commits release the stage / publish the
// accumulator in both, and both epilogues arrive on the leader's accumulator-empty barrier.  (An earlier pair variant
// multicast the B stage instead — half the L2 reads, the same bytes into each SM; neither variant is faster than single
// CTAs: see launch_head.)
template <int MODE, bool FAST, bool TF32, int CL>
__global__ void __launch_bounds__(kThreads, 1) region_head_kernel(const __grid_constant__ CUtensorMap tm_a,
                                                                  const __grid_constant__ CUtensorMap tm_b, const HeadParams p) {
  constexpr int KB = TF32 ? 32 : 64;                           // k per stage: 128 bytes of K per B row either way
  constexpr int ES = TF32 ? 4 : 2;                             // element size in shared memory
  constexpr int kABlk = KB * 128;                              // one MN block of the A tile: [KB k rows x 128 bytes of m]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  HeadShared* sh = reinterpret_cast<HeadShared*>(smem + kStages * kStageBytes);
  const WaitCtx wc{&sh->abort_flag, p.err};

  const int tid = threadIdx.x, warp = warp_index(), lane = tid & 31;
  const int n_items = (MODE == kDW ? p.m_tiles * p.n_tiles * p.splits : p.batches * p.m_tiles * p.n_tiles);
  // item walk: cluster c of the grid takes item pairs c, c + #clusters, ...; its CTA of rank r the r-th item of the pair
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int item0 = ((int)blockIdx.x / CL) * CL + crank, item_step = (int)gridDim.x;
  const int kbpb = (p.K + KB - 1) / KB;                       // k-blocks per batch item

  if (tid == 0) {
    sh->abort_flag = 0;
    for (int s = 0; s < kStages; ++s) { mbar_init(&sh->full[s], FAST ? 1 : kStageThreads); mbar_init(&sh->empty[s], 1); }
    if (FAST) { tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sh->acc_full[i], 1); mbar_init(&sh->acc_empty[i], CL * kEpiWarps * 32); }
    fence_barrier_init();
  }
  if (MODE == kFwd)
    for (int i = tid; i < kN; i += kThreads) sh->bias[i] = (p.bias && i < p.N) ? p.bias[i] : 0.f;
  if (CL > 1) {
    __syncthreads();
    cluster_sync_all();                                       // both CTAs' barriers exist before the pair-wide allocation / any signal
    if (warp == kMmaWarp) tmem_alloc_pair(&sh->tmem_slot, 2 * kAccCols);
  } else if (warp == kMmaWarp) {
    tmem_alloc(&sh->tmem_slot, 2 * kAccCols);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = sh->tmem_slot;

  if (warp >= kStageWarp0) {
    // =============================== staging warps: global -> bf16 -> swizzled smem ===============================
    const int t = tid - kStageWarp0 * 32;                      // 0..255
    if (FAST) {
      // ---- TMA producer: one thread; boxes of [KB k rows x 128 bytes] land in the swizzled layouts the MMAs read ----
      if (warp == kStageWarp0 && elect_one()) {
        constexpr int kBoxes = kABytes / kABlk;                  // A: 128 m in boxes of 128 bytes of m (2 bf16 / 4 fp32)
        uint32_t stage = 0, phase = 0;
        for (int item = item0; item < n_items; item += item_step) {
          const Item it = decode_item<MODE>(p, item);
          const int steps = (it.b1 - it.b0) * kbpb;
          for (int step = 0; step < steps; ++step) {
            const int b = it.b0 + step / kbpb, k0 = (step % kbpb) * KB;
            uint8_t* sa = smem + stage * kStageBytes;
            mbar_wait(&sh->empty[stage], phase ^ 1, wc, 11);    // the MMAs that read this stage are done
            if (CL > 1) {
              // pair: own A rows + own half of the B rows into OWN shared memory; all bytes complete on the leader's barrier
              if (crank == 0) mbar_expect_tx(&sh->full[stage], CL * (kABytes + kBBytes / CL));
#pragma unroll
              for (int bx = 0; bx < kBoxes; ++bx)
                tma_load_3d_pair(sa + bx * kABlk, &tm_a, it.m0 + bx * (128 / ES), k0, p.a_batch ? b : 0, &sh->full[stage]);
              tma_load_3d_pair(sa + kABytes, &tm_b, k0, it.n0 + crank * (kN / CL), p.b_batch ? b : 0, &sh->full[stage]);
            } else {
              mbar_expect_tx(&sh->full[stage], kStageBytes);     // out-of-range rows / columns arrive as zeros and count
#pragma unroll
              for (int bx = 0; bx < kBoxes; ++bx)
                tma_load_3d(sa + bx * kABlk, &tm_a, it.m0 + bx * (128 / ES), k0, p.a_batch ? b : 0, &sh->full[stage]);
              tma_load_3d(sa + kABytes, &tm_b, k0, it.n0, p.b_batch ? b : 0, &sh->full[stage]);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
      __syncwarp();
    } else {
      const bool a_bf16 = p.a_bf16 != 0, b_bf16 = p.b_bf16 != 0;
      const bool a_vec = (p.lda % (a_bf16 ? 8 : 4) == 0) && (p.a_batch % (a_bf16 ? 8 : 4) == 0);
      const bool b_vec = (p.ldb % (b_bf16 ? 8 : 4) == 0) && (p.b_batch % (b_bf16 ? 8 : 4) == 0);
      const int a_o = t & 15, a_k = t >> 4;                      // A chunk: m octet, k row (+16 per chunk)
      const int b_j = t & 7, b_n = t >> 3;                       // B chunk: k octet, n row (+32 per chunk)
      uint32_t stage = 0, phase = 0;

      // group g of a k-block: 0 = the four A chunks, 1 / 2 = B chunks 0-3 / 4-7
      auto issue = [&](const Item& it, int step, int g, Raw& raw) {
        const int b = it.b0 + step / kbpb, k0 = (step % kbpb) * KB;
        if (g == 0) {
          const int m = it.m0 + a_o * 8;
          const int valid_m = max(0, min(8, p.M - m));
  #pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int k = k0 + a_k + 16 * c;
            const int valid = k < p.K ? valid_m : 0;
            load_chunk(p.a, (long long)b * p.a_batch + (long long)k * p.lda + m, valid, a_bf16, a_vec, raw.v[2 * c], raw.v[2 * c + 1]);
          }
        } else {
          const int k = k0 + b_j * 8;
          const int valid_k = max(0, min(8, p.K - k));
  #pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int n = it.n0 + b_n + 32 * (c + 4 * (g - 1));
            const int valid = n < p.N ? valid_k : 0;
            load_chunk(p.b, (long long)b * p.b_batch + (long long)n * p.ldb + k, valid, b_bf16, b_vec, raw.v[2 * c], raw.v[2 * c + 1]);
          }
        }
      };
      auto commit = [&](int g, const Raw& raw) {
        uint8_t* sa = smem + stage * kStageBytes;
        uint8_t* sb = sa + kABytes;
        if (g == 0) {
          mbar_wait(&sh->empty[stage], phase ^ 1, wc, 11);        // the MMAs that read this stage are done
  #pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 v = to_bf16x8(a_bf16, raw.v[2 * c], raw.v[2 * c + 1]);
            st_chunk(sa + (a_o >> 3) * kABlk, a_k + 16 * c, a_o & 7, v);
          }
        } else {
  #pragma unroll
          for (int c = 0; c < 4; ++c)
            st_chunk(sb, b_n + 32 * (c + 4 * (g - 1)), b_j, to_bf16x8(b_bf16, raw.v[2 * c], raw.v[2 * c + 1]));
          if (g == 2) {
            fence_proxy_async_smem();
            mbar_arrive(&sh->full[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      };

      // rolling pipeline over all groups of all k-blocks of all items: the loads of group i+1 are issued before
      // group i is converted and stored (two groups = 256 bytes per thread in flight); two register sets, ping-pong
      struct Pos { Item it; int item, steps, step, g; };
      auto advance = [&](Pos& q) -> bool {
        if (++q.g == 3) { q.g = 0; ++q.step; }
        if (q.step == q.steps) {
          q.item += gridDim.x; q.step = 0;
          if (q.item >= n_items) return false;
          q.it = decode_item<MODE>(p, q.item); q.steps = (q.it.b1 - q.it.b0) * kbpb;
        }
        return true;
      };
      if ((int)blockIdx.x < n_items) {
        Raw r0, r1;
        Pos pos;
        pos.item = blockIdx.x; pos.it = decode_item<MODE>(p, pos.item); pos.steps = (pos.it.b1 - pos.it.b0) * kbpb; pos.step = 0; pos.g = 0;
        issue(pos.it, pos.step, pos.g, r0);
        while (true) {
          Pos np = pos;
          bool more = advance(np);
          if (more) issue(np.it, np.step, np.g, r1);
          commit(pos.g, r0);
          if (!more) break;
          pos = np;
          more = advance(np);
          if (more) issue(np.it, np.step, np.g, r0);
          commit(pos.g, r1);
          if (!more) break;
          pos = np;
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =============================== MMA issue: one elected thread ===============================
    if ((CL == 1 || crank == 0) && elect_one()) {               // a pair's MMAs are issued by its leader alone
      constexpr uint32_t idesc = TF32 ? idesc_tf32(kM * CL, kN, true, false) : idesc_bf16(kM * CL, kN, true, false);
      uint32_t stage = 0, phase = 0, n_acc = 0;
      for (int item = item0; item < n_items; item += item_step, ++n_acc) {
        const Item it = decode_item<MODE>(p, item);
        const int steps = (it.b1 - it.b0) * kbpb;
        const uint32_t buf = n_acc & 1;
        mbar_wait(&sh->acc_empty[buf], ((n_acc >> 1) & 1) ^ 1, wc, 12);      // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem + buf * kAccCols;
        for (int step = 0; step < steps; ++step) {
          mbar_wait(&sh->full[stage], phase, wc, 13);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes), sb = sa + kABytes;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {                                   // 16 k (bf16) / 8 k (tf32) per MMA
            // MN-major: a quarter of the k rows per step; tf32: 4-row swizzle groups (512 bytes), bf16: 8-row groups
            const Desc da = TF32 ? smem_desc_base32b(sa + ks * (kABlk / 4), kABlk, 512) : make_desc(sa + ks * (kABlk / 4), kABlk, 1024);
            const Desc db = make_desc(sb + ks * 32, 16, 1024);              // K-major: 32 bytes inside the row per step
            if (CL > 1) {
              if (TF32) mma_ss_tf32_pair(acc, da, db, idesc, step > 0 || ks > 0);
              else mma_ss_pair(acc, da, db, idesc, step > 0 || ks > 0);
            } else {
              if (TF32) mma_ss_tf32(acc, da, db, idesc, step > 0 || ks > 0);
              else mma_ss(acc, da, db, idesc, step > 0 || ks > 0);
            }
          }
          if (CL > 1) mma_commit_pair(&sh->empty[stage]);                    // releases the stage in both CTAs
          else mma_commit(&sh->empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (CL > 1) mma_commit_pair(&sh->acc_full[buf]);                     // both epilogues
        else mma_commit(&sh->acc_full[buf]);
      }
    }
    __syncwarp();
  } else {
    // ======= epilogue warps 0..7: TMEM lane quadrant = warp & 3 (32 tile rows), column half = warp >> 2 (128 columns) =======
    uint32_t n_acc = 0;
    const int quad = warp & 3, half = warp >> 2;
    const int rl = quad * 32 + lane;
    for (int item = item0; item < n_items; item += item_step, ++n_acc) {
      const Item it = decode_item<MODE>(p, item);
      const uint32_t buf = n_acc & 1;
      mbar_wait(&sh->acc_full[buf], (n_acc >> 1) & 1, wc, 14);
      tc_fence_after();
      const uint32_t acc = tmem + buf * kAccCols + (static_cast<uint32_t>(quad * 32) << 16);
      const int m = it.m0 + rl;
      uint32_t* stg = sh->stg[warp];
      if (MODE == kFwd) {
        // row = pixel m of image out_b: y = acc + bias, ||y||, unit row (bf16) + norm; pad rows [R, Rpad) are zeros
        float ss = 0.f;
#pragma unroll 1
        for (int c = half * 4; c < half * 4 + 4; ++c) {
          uint32_t v[32];
          tmem_ld32(acc + c * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) { const float y = __uint_as_float(v[e]) + sh->bias[c * 32 + e]; ss = fmaf(y, y, ss); }
        }
        sh->ss_part[buf][half][rl] = ss;                       // the other column half of this row lives in warp (warp ^ 4)
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        ss += sh->ss_part[buf][half ^ 1][rl];
        // a timed-out pipeline wait (abort flag) must not pass for a result: NaN norms poison the loss downstream
        const float nrm = *wc.abort_flag ? __int_as_float(0x7fc00000) : fmaxf(sqrtf(ss), kEps), inv = (m < p.R) ? 1.f / nrm : 0.f;
        const bool wr = m < p.Rpad;
        const int mw = it.m0 + quad * 32;                      // first row of this warp's block
        __nv_bfloat16* dstw = p.kn + ((size_t)it.out_b * p.Rpad + mw) * kN;
#pragma unroll 1
        for (int c = half * 4; c < half * 4 + 4; ++c) {
          uint32_t v[32];
          tmem_ld32(acc + c * 32, v);
          tmem_wait_ld();
          uint32_t w[16];
#pragma unroll
          for (int e = 0; e < 16; ++e)
            w[e] = pack_bf16((__uint_as_float(v[2 * e]) + sh->bias[c * 32 + 2 * e]) * inv,
                             (__uint_as_float(v[2 * e + 1]) + sh->bias[c * 32 + 2 * e + 1]) * inv);
          warp_rows_out(stg, lane, w, [&](int r, int wo, uint4 val) {
            if (mw + r < p.Rpad) *reinterpret_cast<uint4*>(dstw + (size_t)r * kN + c * 32 + 2 * wo) = val;
          });
        }
        if (wr && half == 0) p.rnorm[(size_t)it.out_b * p.Rpad + m] = (m < p.R) ? nrm : 0.f;
      } else if (MODE == kDFeat) {
        // row = input channel m, columns = pixels n0 + c: dfeat[b][m][n]
        const bool row_ok = m < p.M;
        const size_t row_off = ((size_t)it.out_b * p.M + m) * p.N;
        const bool vec = (p.N % 8 == 0);
#pragma unroll 1
        for (int c = half * 4; c < half * 4 + 4; ++c) {
          const int n = it.n0 + c * 32;
          if (n >= p.N) break;                                   // uniform over the warp
          uint32_t v[32];
          tmem_ld32(acc + c * 32, v);
          tmem_wait_ld();
          if (*wc.abort_flag) v[0] = 0x7fc00000u;                // timed-out wait: never a plausible gradient
          if (vec && n + 32 <= p.N) {                            // warp-uniform: whole row segments through the staging block
            const int mw = it.m0 + quad * 32;
            const size_t roww = ((size_t)it.out_b * p.M + mw) * p.N + n;
            if (p.out_bf16) {
              uint32_t w[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) w[e] = pack_bf16(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
              __nv_bfloat16* dstw = static_cast<__nv_bfloat16*>(p.dfeat) + roww;
              warp_rows_out(stg, lane, w, [&](int r, int wo, uint4 val) {
                if (mw + r < p.M) *reinterpret_cast<uint4*>(dstw + (size_t)r * p.N + 2 * wo) = val;
              });
            } else {
              float* dstw = static_cast<float*>(p.dfeat) + roww;
#pragma unroll
              for (int hf = 0; hf < 2; ++hf)                     // 32 fp32 columns = two blocks of 16 words
                warp_rows_out(stg, lane, v + 16 * hf, [&](int r, int wo, uint4 val) {
                  if (mw + r < p.M) *reinterpret_cast<uint4*>(dstw + (size_t)r * p.N + 16 * hf + wo) = val;
                });
            }
          } else if (!row_ok) {
            continue;
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (n + e < p.N) {
                if (p.out_bf16) static_cast<__nv_bfloat16*>(p.dfeat)[row_off + n + e] = __float2bfloat16_rn(__uint_as_float(v[e]));
                else static_cast<float*>(p.dfeat)[row_off + n + e] = __uint_as_float(v[e]);
              }
          }
        }
      } else {
        // row = output feature m (d), columns = input channels n0 + c: partial dW over this CTA's batch items
        const bool row_ok = m < p.M;
#pragma unroll 1
        for (int c = half * 4; c < half * 4 + 4; ++c) {
          const int n = it.n0 + c * 32;
          if (n >= p.N) break;
          uint32_t v[32];
          tmem_ld32(acc + c * 32, v);
          tmem_wait_ld();
          if (*wc.abort_flag) v[0] = 0x7fc00000u;
          float* dst = p.dw + (size_t)m * p.N + n;
          if (n + 32 <= p.N && p.N % 4 == 0) {                   // warp-uniform: 128-byte row segments per reduction group
            const int mw = it.m0 + quad * 32;
            float* dstw = p.dw + (size_t)mw * p.N + n;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
              warp_rows_out(stg, lane, v + 16 * hf, [&](int r, int wo, uint4 val) {
                if (mw + r < p.M)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dstw + (size_t)r * p.N + 16 * hf + wo),
                               "f"(__uint_as_float(val.x)), "f"(__uint_as_float(val.y)), "f"(__uint_as_float(val.z)), "f"(__uint_as_float(val.w))
                               : "memory");
              });
          } else if (!row_ok) {
            continue;
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (n + e < p.N) atomicAdd(dst + e, __uint_as_float(v[e]));
          }
        }
      }
      tc_fence_before();
      if (CL > 1 && crank != 0) mbar_arrive_remote(&sh->acc_empty[buf], 0);   // the leader's MMA thread waits for both epilogues
      else mbar_arrive(&sh->acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) {
    cluster_sync_all();                                       // no CTA leaves while its peer may still signal into it
    if (warp == kMmaWarp) tmem_dealloc_pair(tmem, 2 * kAccCols);
  } else if (warp == kMmaWarp) {
    tmem_dealloc(tmem, 2 * kAccCols);
  }
}

constexpr int kHeadSmem = kStages * kStageBytes + (int)sizeof(HeadShared) + 1024;

// column sums of dy[rows, 256] -> dbias (fp32 atomics into a zero-filled vector): thread = (8 columns, one of 8 row lanes)
template <typename T>
__global__ void __launch_bounds__(256) head_colsum_kernel(const T* __restrict__ dy, long long rows, float* __restrict__ dbias) {
  __shared__ float part[8][kN];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const long long per = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 4
  for (long long r = r0 + rl; r < r1; r += 8) {
    const float4 a = ld4_nc(dy + r * kN + cg * 8), b = ld4_nc(dy + r * kN + cg * 8 + 4);
    acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[rl][cg * 8 + e] = acc[e];
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += part[k][threadIdx.x];
  atomicAdd(dbias + threadIdx.x, s);
}

// ---- pooled embedding: x[B, C, P] -> out[B, C] = mean over the P pixels ------------------------------------------
// F.avg_pool2d(x, kernel_size = H).view(B, -1) on the discriminator's last 4 x 4 stage: the producer of sent_loss's image
// operand (df_gan.py:165-166) and of both img_loss operands (train_gan.py:271-276).  One thread per (b, c): P
// contiguous elements in, one out (optionally cast to bf16 for the bf16 similarity path); HBM-bound, 8.4 MB at
// B = 256, C = 512, P = 16.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) avgpool_rows_kernel(const TI* __restrict__ x, long long n, int P, float inv, TO* __restrict__ out) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const TI* s = x + i * P;
  float acc = 0.f;
  if (P % 4 == 0) {
    for (int q = 0; q < P; q += 4) { const float4 v = ld4_nc(s + q); acc += (v.x + v.y) + (v.z + v.w); }
  } else {
    for (int q = 0; q < P; ++q) acc += ld1(s + q);
  }
  st1(out + i, acc * inv);
}
template <typename TG, typename TO>
__global__ void __launch_bounds__(256) avgpool_rows_bwd_kernel(const TG* __restrict__ dout, long long n, int P, float inv, TO* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float g = ld1(dout + i) * inv;
  TO* d = dx + i * P;
  if (P % 4 == 0) {
    for (int q = 0; q < P; q += 4) st4(d + q, make_float4(g, g, g, g));
  } else {
    for (int q = 0; q < P; ++q) st1(d + q, g);
  }
}

// fast path: both operands of one dtype, every row start 16-byte aligned (cp.async)
bool fast_ok(const HeadParams& p) {
  if (p.a_bf16 != p.b_bf16) return false;
  const int epc = p.a_bf16 ? 8 : 4;
  return p.lda % epc == 0 && p.ldb % epc == 0 && p.a_batch % epc == 0 && p.b_batch % epc == 0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled head_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// 3-D map over an operand stored [batch][rows][inner contiguous]: box = 128 bytes of the inner axis x box_rows x 1.
// mn_major_f32: the MN-major fp32 operand of kind::tf32 needs the 32-byte-atom flavour of the 128-byte swizzle.
int head_map(CUtensorMap* m, const void* base, bool bf16, int inner, int rows, int batches, long long row_stride,
             long long batch_stride, int box_rows, bool mn_major_f32) {
  PFN_encodeTiled enc = head_encode();
  XMC_REQUIRE(enc != nullptr, XMC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int es = bf16 ? 2 : 4;
  if (batch_stride == 0) { batches = 1; batch_stride = (long long)rows * row_stride; }
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batches};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride * es, (cuuint64_t)batch_stride * es};
  cuuint32_t box[3] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major_f32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XMC_REQUIRE(r == CUDA_SUCCESS, XMC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return XMC_OK;
}

template <int MODE, bool FAST, bool TF32, int CL>
int launch_head_as(const HeadParams& p, int grid, cudaStream_t st) {
  CUtensorMap ta{}, tb{};
  if (FAST) {
    // A_b[k][m]: inner = m (extent M), rows = K;  B_b[n][k]: inner = k (extent K), rows = N (a pair loads half the rows each)
    if (int rc = head_map(&ta, p.a, !TF32, p.M, p.K, p.batches, p.lda, p.a_batch, TF32 ? 32 : 64, TF32)) return rc;
    if (int rc = head_map(&tb, p.b, !TF32, p.K, p.N, p.batches, p.ldb, p.b_batch, kN / CL, false)) return rc;
  }
  auto kern = region_head_kernel<MODE, FAST, TF32, CL>;
  XMC_RETURN_IF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kHeadSmem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = CL > 1 ? 1 : 0;
  XMC_RETURN_IF_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  return cuda_fail(cudaGetLastError(), "region_head_kernel launch");
}

template <int MODE>
int launch_head(const HeadParams& p, int n_items, cudaStream_t st) {
  int dev = 0, sms = 148;
  XMC_RETURN_IF_CUDA(cudaGetDevice(&dev));
  XMC_RETURN_IF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = n_items < sms ? n_items : sms;
  if (!fast_ok(p)) return launch_head_as<MODE, false, false, 1>(p, grid, st);
  // pairs: consecutive items must share their B operand and there must be an even number of them
  //   forward: B = the weights, shared by every item;  dfeat: B = dy_b, shared by the channel tiles of one image
#ifdef XMC_TEST_HOOKS
  // Hooks build only (libxmcloss_hooks.so, XMC_HEAD_PAIR=1): the 2-SM form.  Measured no faster than one CTA per tile (forward
  // 43.0 / 32.8 us against 43.0 / 30.7 us) — per staged byte the kernel writes shared memory once (TMA) and reads it once (MMA
  // operands), 192 B/clk against the SM's 128 B/clk while the MMAs run, and pairing halves only the B part of it.  Kept as
  // the tested 2-SM form of the kernel; the product library has the single-CTA form only.
  const bool pairs = MODE != kDW && n_items % 2 == 0 && n_items >= 2 && (MODE == kFwd || (p.n_tiles == 1 && p.m_tiles % 2 == 0)) &&
                     getenv("XMC_HEAD_PAIR") != nullptr;
  if (pairs) {
    grid &= ~1;
    return p.a_bf16 ? launch_head_as<MODE, true, false, 2>(p, grid, st) : launch_head_as<MODE, true, true, 2>(p, grid, st);
  }
#endif
  return p.a_bf16 ? launch_head_as<MODE, true, false, 1>(p, grid, st) : launch_head_as<MODE, true, true, 1>(p, grid, st);
}

int check_dt(int dt) { return dt == XMC_F32 || dt == XMC_BF16; }

}  // namespace
}  // namespace xmc

using namespace xmc;

extern "C" int xmc_region_head_forward(const void* feat, int feat_dtype, const void* weight, int weight_dtype,
                                       const float* bias, int B, int Cin, int R, int Rpad, int D,
                                       void* kn, float* rnorm, void* stream) {
  XMC_REQUIRE(feat && weight && kn && rnorm, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(check_dt(feat_dtype) && check_dt(weight_dtype), XMC_ERR_UNSUPPORTED, "dtype must be XMC_F32 or XMC_BF16");
  XMC_REQUIRE(B > 0 && Cin > 0 && R > 0 && Rpad >= R && Rpad % 16 == 0, XMC_ERR_INVALID_ARG, "bad shape B=%d Cin=%d R=%d Rpad=%d", B, Cin, R, Rpad);
  XMC_REQUIRE(D == kN, XMC_ERR_UNSUPPORTED, "region head: D=%d unsupported (256)", D);
  XMC_REQUIRE(aligned16(feat) && aligned16(weight) && aligned16(kn), XMC_ERR_ALIGNMENT, "pointers must be 16-byte aligned");
  HeadParams p{};
  p.a = feat; p.a_batch = (long long)Cin * R; p.lda = R; p.a_bf16 = feat_dtype == XMC_BF16;
  p.b = weight; p.b_batch = 0; p.ldb = Cin; p.b_bf16 = weight_dtype == XMC_BF16;
  p.M = R; p.N = D; p.K = Cin;
  p.m_tiles = (Rpad + kM - 1) / kM; p.n_tiles = 1; p.batches = B; p.splits = 1;
  p.bias = bias; p.kn = static_cast<__nv_bfloat16*>(kn); p.rnorm = rnorm; p.R = R; p.Rpad = Rpad;
  return launch_head<kFwd>(p, B * p.m_tiles, as_stream(stream));
}

extern "C" int xmc_region_head_backward_input(const void* weight, int weight_dtype, const void* dy, int dy_dtype, int B, int Cin,
                                              int R, int D, void* dfeat, int out_dtype, void* stream) {
  XMC_REQUIRE(weight && dy && dfeat, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(check_dt(weight_dtype) && check_dt(out_dtype) && check_dt(dy_dtype), XMC_ERR_UNSUPPORTED, "dtype must be XMC_F32 or XMC_BF16");
  XMC_REQUIRE(B > 0 && Cin > 0 && R > 0, XMC_ERR_INVALID_ARG, "bad shape B=%d Cin=%d R=%d", B, Cin, R);
  XMC_REQUIRE(D == kN, XMC_ERR_UNSUPPORTED, "region head: D=%d unsupported (256)", D);
  XMC_REQUIRE(aligned16(weight) && aligned16(dy) && aligned16(dfeat), XMC_ERR_ALIGNMENT, "pointers must be 16-byte aligned");
  HeadParams p{};
  p.a = weight; p.a_batch = 0; p.lda = Cin; p.a_bf16 = weight_dtype == XMC_BF16;      // A[k = d][m = channel]
  p.b = dy; p.b_batch = (long long)R * D; p.ldb = D; p.b_bf16 = dy_dtype == XMC_BF16;  // B[n = pixel][k = d]
  p.M = Cin; p.N = R; p.K = D;
  p.m_tiles = (Cin + kM - 1) / kM; p.n_tiles = (R + kN - 1) / kN; p.batches = B; p.splits = 1;
  p.dfeat = dfeat; p.out_bf16 = out_dtype == XMC_BF16;
  return launch_head<kDFeat>(p, B * p.m_tiles * p.n_tiles, as_stream(stream));
}

extern "C" int xmc_region_head_backward_weight(const void* feat, int feat_dtype, const void* dy, int dy_dtype, int B, int Cin, int R,
                                               int D, float* dweight, float* dbias, void* stream) {
  XMC_REQUIRE(feat && dy && dweight, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(check_dt(feat_dtype) && check_dt(dy_dtype), XMC_ERR_UNSUPPORTED, "dtype must be XMC_F32 or XMC_BF16");
  XMC_REQUIRE(B > 0 && Cin > 0 && R > 0, XMC_ERR_INVALID_ARG, "bad shape B=%d Cin=%d R=%d", B, Cin, R);
  XMC_REQUIRE(D == kN, XMC_ERR_UNSUPPORTED, "region head: D=%d unsupported (256)", D);
  XMC_REQUIRE(aligned16(feat) && aligned16(dy) && aligned16(dweight), XMC_ERR_ALIGNMENT, "pointers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  XMC_RETURN_IF_CUDA(cudaMemsetAsync(dweight, 0, sizeof(float) * (size_t)D * Cin, st));
  if (dbias) {
    XMC_RETURN_IF_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)D, st));
    const long long rows = (long long)B * R;
    const unsigned grid = (unsigned)(rows < 296 * 8 ? (rows + 7) / 8 : 296);
    if (dy_dtype == XMC_BF16) head_colsum_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dy), rows, dbias);
    else head_colsum_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(dy), rows, dbias);
    XMC_RETURN_IF_CUDA(cudaGetLastError());
  }
  HeadParams p{};
  p.a = dy; p.a_batch = (long long)R * D; p.lda = D; p.a_bf16 = dy_dtype == XMC_BF16;   // A[k = pixel][m = d]
  p.b = feat; p.b_batch = (long long)Cin * R; p.ldb = R; p.b_bf16 = feat_dtype == XMC_BF16;   // B[n = channel][k = pixel]
  p.M = D; p.N = Cin; p.K = R;
  p.m_tiles = D / kM; p.n_tiles = (Cin + kN - 1) / kN; p.batches = B;
  int dev = 0, sms = 148;
  XMC_RETURN_IF_CUDA(cudaGetDevice(&dev));
  XMC_RETURN_IF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int tiles = p.m_tiles * p.n_tiles;
  int splits = sms / tiles;
  if (splits < 1) splits = 1;
  if (splits > B) splits = B;
  // every split must own at least one batch item (an empty split would publish an unwritten accumulator)
  const int per = (B + splits - 1) / splits;
  splits = (B + per - 1) / per;
  p.splits = splits;
  p.dw = dweight; p.dbias = dbias;
  return launch_head<kDW>(p, tiles * splits, st);
}

/* pooled embedding and its backward (see avgpool_rows_kernel) */
extern "C" int xmc_avgpool_rows(const void* x, int in_dtype, int B, int C, int P, void* out, int out_dtype, void* stream) {
  XMC_REQUIRE(x && out, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(check_dt(in_dtype) && check_dt(out_dtype), XMC_ERR_UNSUPPORTED, "dtype must be XMC_F32 or XMC_BF16");
  XMC_REQUIRE(B > 0 && C > 0 && P > 0, XMC_ERR_INVALID_ARG, "bad shape B=%d C=%d P=%d", B, C, P);
  XMC_REQUIRE(aligned16(x), XMC_ERR_ALIGNMENT, "x must be 16-byte aligned");
  const long long n = (long long)B * C;
  const unsigned grid = (unsigned)((n + 255) / 256);
  const float inv = 1.f / (float)P;
  cudaStream_t st = as_stream(stream);
  if (in_dtype == XMC_F32) {
    if (out_dtype == XMC_F32) avgpool_rows_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(x), n, P, inv, static_cast<float*>(out));
    else avgpool_rows_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(x), n, P, inv, static_cast<__nv_bfloat16*>(out));
  } else {
    if (out_dtype == XMC_F32) avgpool_rows_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), n, P, inv, static_cast<float*>(out));
    else avgpool_rows_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), n, P, inv, static_cast<__nv_bfloat16*>(out));
  }
  return cuda_fail(cudaGetLastError(), "avgpool_rows_kernel launch");
}

extern "C" int xmc_avgpool_rows_backward(const void* dout, int g_dtype, int B, int C, int P, void* dx, int out_dtype, void* stream) {
  XMC_REQUIRE(dout && dx, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(check_dt(g_dtype) && check_dt(out_dtype), XMC_ERR_UNSUPPORTED, "dtype must be XMC_F32 or XMC_BF16");
  XMC_REQUIRE(B > 0 && C > 0 && P > 0, XMC_ERR_INVALID_ARG, "bad shape B=%d C=%d P=%d", B, C, P);
  XMC_REQUIRE(aligned16(dx), XMC_ERR_ALIGNMENT, "dx must be 16-byte aligned");
  const long long n = (long long)B * C;
  const unsigned grid = (unsigned)((n + 255) / 256);
  const float inv = 1.f / (float)P;
  cudaStream_t st = as_stream(stream);
  if (g_dtype == XMC_F32) {
    if (out_dtype == XMC_F32) avgpool_rows_bwd_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(dout), n, P, inv, static_cast<float*>(dx));
    else avgpool_rows_bwd_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(dout), n, P, inv, static_cast<__nv_bfloat16*>(dx));
  } else {
    if (out_dtype == XMC_F32) avgpool_rows_bwd_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dout), n, P, inv, static_cast<float*>(dx));
    else avgpool_rows_bwd_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dout), n, P, inv, static_cast<__nv_bfloat16*>(dx));
  }
  return cuda_fail(cudaGetLastError(), "avgpool_rows_bwd_kernel launch");
}
