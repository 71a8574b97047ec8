// simloss.cu — sentence–image / image–image InfoNCE (sent_loss, img_loss) for sm_100a.
//
// Replaces xmc_gan/train_gan.py:85-139 of the reference: cosine_scores (:85-91) and the
// bidirectional label-weighted log-softmax tail (:103-113 / :127-137).
//
// The problem is tiny (Bq,Bk = 256..2048, D = 256/512: 0.1–0.8 GFLOP, ~1–7 MB) so it is
// launch-latency / HBM bound; no tensor cores.  Forward is ONE kernel: each CTA keeps ROWS unit
// vectors of one side in registers and streams the other side once, one warp per streamed
// vector, 128-bit coalesced loads, warp-shuffle dot products and an online log-sum-exp.  "Row"
// CTAs own rows of the score matrix, "column" CTAs own columns, so both softmax directions are
// complete inside the launch with no atomics and no second pass.  Backward is ONE kernel with
// the same ownership: d scores in closed form on the fly, rank-1 updates into register
// accumulators, normalise-backward in the epilogue.
#include "common.cuh"
#include "simloss.cuh"

namespace xmc {

constexpr int kSimThreads = 256;
constexpr int kSimWarps = kSimThreads / 32;

// Load ROWS resident vectors (rows r0.. of X[N,D]) into registers as unit vectors.
template <typename T, int ROWS, int DCH>
__device__ __forceinline__ void load_resident(const T* X, int N, int D, int r0, int lane,
                                              float4 (&xr)[ROWS][DCH], float (&inv)[ROWS]) {
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    float nn = 0.f;
    const bool ok = (r0 + r) < N;
#pragma unroll
    for (int c = 0; c < DCH; ++c) {
      int col = c * 128 + lane * 4;
      xr[r][c] = (ok && col < D) ? ld4(X + (size_t)(r0 + r) * D + col) : make_float4(0, 0, 0, 0);
      nn += dot4(xr[r][c], xr[r][c]);
    }
    nn = warp_sum(nn);
    inv[r] = 1.f / fmaxf(sqrtf(nn), kEps);
#pragma unroll
    for (int c = 0; c < DCH; ++c) {
      xr[r][c].x *= inv[r]; xr[r][c].y *= inv[r]; xr[r][c].z *= inv[r]; xr[r][c].w *= inv[r];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Fused forward: normalise + cosine + scale + row/column log-sum-exp + label sums.
// ------------------------------------------------------------------------------------------
template <typename T, int ROWS, int DCH>
__global__ void __launch_bounds__(kSimThreads) sim_fwd_kernel(SimParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool row_side = (int)blockIdx.x < p.n_row_blocks;
  const int blk = row_side ? blockIdx.x : blockIdx.x - p.n_row_blocks;
  const T* X = static_cast<const T*>(row_side ? p.a : p.b);   // resident side
  const T* Y = static_cast<const T*>(row_side ? p.b : p.a);   // streamed side
  const int NX = row_side ? p.Bq : p.Bk, NY = row_side ? p.Bk : p.Bq;
  const int r0 = blk * ROWS;

  float4 xr[ROWS][DCH];
  float inv[ROWS];
  load_resident<T, ROWS, DCH>(X, NX, p.D, r0, lane, xr, inv);
  float* inv_out = row_side ? p.inv_a : p.inv_b;
  if (inv_out && warp == 0 && lane < ROWS && r0 + lane < NX) {
    float v = inv[0];
#pragma unroll
    for (int r = 1; r < ROWS; ++r) v = (lane == r) ? inv[r] : v;
    inv_out[r0 + lane] = v;
  }

  const int mine = owner_row<ROWS>(lane);        // resident vector whose dot this lane ends up with
  const int xi = r0 + mine;
  const bool xi_ok = xi < NX;
  Stat st; st.init();

  // the streamed vector of the NEXT iteration is fetched before this one is consumed (the loop is
  // otherwise a chain of exposed L2 latencies)
  auto fetch = [&](int y, float4 (&yv)[DCH]) {
#pragma unroll
    for (int c = 0; c < DCH; ++c) {
      const int col = c * 128 + lane * 4;
      yv[c] = (y < NY && col < p.D) ? ld4_nc(Y + (size_t)y * p.D + col) : make_float4(0, 0, 0, 0);
    }
  };
  float4 ynext[DCH];
  fetch(warp, ynext);
  for (int y = warp; y < NY; y += kSimWarps) {
    float4 ycur[DCH];
#pragma unroll
    for (int c = 0; c < DCH; ++c) ycur[c] = ynext[c];
    fetch(y + kSimWarps, ynext);
    float dots[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) dots[r] = 0.f;
    float nn = 0.f;
#pragma unroll
    for (int c = 0; c < DCH; ++c) {
      const float4 yv = ycur[c];
      nn += dot4(yv, yv);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) dots[r] += dot4(yv, xr[r][c]);
    }
    nn = warp_sum(nn);
    const float invy = 1.f / fmaxf(sqrtf(nn), kEps);
    const float s = warp_multi_sum<ROWS>(dots, lane) * invy;   // cosine(x_mine, y)
    if (xi_ok) {
      const int i = row_side ? xi : y, j = row_side ? y : xi;
      if (row_side && (lane & (32 / ROWS - 1)) == 0) p.scores[(size_t)i * p.Bk + j] = s;
      st.add(p.scale * s, label_at(p.labels, p.Bk, i, j, p.diag));
    }
  }

  // combine the 8 warps' partial statistics
  __shared__ Stat sh[kSimWarps][ROWS];
  if ((lane & (32 / ROWS - 1)) == 0) sh[warp][mine] = st;
  __syncthreads();
  if (threadIdx.x < ROWS && r0 + (int)threadIdx.x < NX) {
    Stat t = sh[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < kSimWarps; ++w) t.merge(sh[w][threadIdx.x]);
    float* out = row_side ? p.row_stats : p.col_stats;
    const int k = r0 + threadIdx.x;
    out[k] = t.m + logf(t.s);
    out[NX + k] = t.sl;
    out[2 * NX + k] = t.slz;
  }
}

// ------------------------------------------------------------------------------------------
// Fused backward: d scores on the fly, dX = sum_y dS(x,y) * yhat, then normalise-backward.
// "a" CTAs own ROWS rows of A (produce dA), "b" CTAs own ROWS rows of B (produce dB).
// ------------------------------------------------------------------------------------------
template <typename T, int ROWS, int DCH>
__global__ void __launch_bounds__(kSimThreads) sim_bwd_kernel(SimParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool a_side = (int)blockIdx.x < p.n_a_blocks;
  const int blk = a_side ? blockIdx.x : blockIdx.x - p.n_a_blocks;
  const T* X = static_cast<const T*>(a_side ? p.a : p.b);
  const T* Y = static_cast<const T*>(a_side ? p.b : p.a);
  const float* inv_x = a_side ? p.inv_a : p.inv_b;
  const float* inv_y = a_side ? p.inv_b : p.inv_a;
  const int NX = a_side ? p.Bq : p.Bk, NY = a_side ? p.Bk : p.Bq;
  const float* xs = a_side ? p.row_stats : p.col_stats;   // statistics of the resident side
  const float* ys = a_side ? p.col_stats : p.row_stats;
  const float* xdiv = a_side ? p.row_div : p.col_div;
  const float* ydiv = a_side ? p.col_div : p.row_div;
  const float x_tot = a_side ? p.inv_rows_total : p.inv_cols_total;
  const float y_tot = a_side ? p.inv_cols_total : p.inv_rows_total;
  const int r0 = blk * ROWS;

  __shared__ float ds_sh[kSimWarps][ROWS][32];
  __shared__ float red[ROWS][DCH * 128];
  for (int k = threadIdx.x; k < ROWS * DCH * 128; k += kSimThreads) (&red[0][0])[k] = 0.f;

  const bool given = p.ds_given != nullptr;
  float x_lse[ROWS], x_sl[ROWS], x_inv_n[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int k = min(r0 + r, NX - 1);
    x_lse[r] = given ? 0.f : xs[k]; x_sl[r] = given ? 0.f : xs[NX + k];
    x_inv_n[r] = given ? 0.f : x_tot / (xdiv ? xdiv[k] : p.num_pos);
  }
  const float go = given ? 1.f : __ldg(p.grad_out) * p.scale;

  float4 acc[ROWS][DCH];
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int c = 0; c < DCH; ++c) acc[r][c] = make_float4(0, 0, 0, 0);

  for (int y0 = warp * 32; y0 < NY; y0 += kSimWarps * 32) {
    const int y = y0 + lane;
    const bool y_ok = y < NY;
    float y_lse = 0.f, y_sl = 0.f, y_inv_n = 0.f, y_inv = 0.f;
    if (y_ok) {
      if (!given) {
        y_lse = ys[y]; y_sl = ys[NY + y];
        y_inv_n = y_tot / (ydiv ? ydiv[y] : p.num_pos);
      }
      y_inv = inv_y[y];
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      float v = 0.f;
      const int x = r0 + r;
      if (y_ok && x < NX) {
        const int i = a_side ? x : y, j = a_side ? y : x;
        if (given) {
          v = __ldg(p.ds_given + (size_t)i * p.Bk + j) * y_inv;
        } else {
          const float z = p.scale * __ldg(p.scores + (size_t)i * p.Bk + j);
          const float lab = label_at(p.labels, p.Bk, i, j, p.diag);
          // dscore is symmetric in (row-stat, col-stat) roles
          v = go * dscore(z, lab, x_lse[r], x_sl[r], x_inv_n[r], y_lse, y_sl, y_inv_n) * y_inv;
        }
      }
      ds_sh[warp][r][lane] = v;
    }
    __syncwarp();
    const int cnt = min(32, NY - y0);
    // rows beyond cnt carry weight 0 in ds_sh (v = 0 above), so the loop runs in groups of 4 with the
    // four streamed vectors fetched up front
    for (int t0 = 0; t0 < cnt; t0 += 4) {
      float4 yv[4][DCH];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int c = 0; c < DCH; ++c) {
          const int col = c * 128 + lane * 4;
          yv[u][c] = (t0 + u < cnt && col < p.D) ? ld4_nc(Y + (size_t)(y0 + t0 + u) * p.D + col) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          const float w = ds_sh[warp][r][t0 + u];
#pragma unroll
          for (int c = 0; c < DCH; ++c) fma4(acc[r][c], w, yv[u][c]);
        }
    }
    __syncwarp();
  }
  __syncthreads();   // red[] zero-fill visible
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int c = 0; c < DCH; ++c) {
      float* dst = &red[r][c * 128 + lane * 4];
      atomicAdd(dst + 0, acc[r][c].x); atomicAdd(dst + 1, acc[r][c].y);
      atomicAdd(dst + 2, acc[r][c].z); atomicAdd(dst + 3, acc[r][c].w);
    }
  __syncthreads();

  // normalise-backward: dx = (g - xhat (xhat.g)) / max(||x||, eps); one warp per resident row
  T* DX = static_cast<T*>(a_side ? p.da : p.db);
  for (int r = warp; r < ROWS; r += kSimWarps) {
    const int x = r0 + r;
    if (x >= NX) continue;
    const float ix = inv_x[x];
    float4 g[DCH], xh[DCH];
    float proj = 0.f;
#pragma unroll
    for (int c = 0; c < DCH; ++c) {
      int col = c * 128 + lane * 4;
      g[c] = *reinterpret_cast<float4*>(&red[r][col]);
      xh[c] = (col < p.D) ? ld4(X + (size_t)x * p.D + col) : make_float4(0, 0, 0, 0);
      xh[c].x *= ix; xh[c].y *= ix; xh[c].z *= ix; xh[c].w *= ix;
      proj += dot4(g[c], xh[c]);
    }
    proj = warp_sum(proj);
    if (ix >= 1.f / kEps) proj = 0.f;      // ||x|| clamped to eps: x/eps is linear in x
#pragma unroll
    for (int c = 0; c < DCH; ++c) {
      int col = c * 128 + lane * 4;
      if (col < p.D) {
        float4 o = make_float4((g[c].x - xh[c].x * proj) * ix, (g[c].y - xh[c].y * proj) * ix,
                               (g[c].z - xh[c].z * proj) * ix, (g[c].w - xh[c].w * proj) * ix);
        st4(DX + (size_t)x * p.D + col, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// cosine_scores only (train_gan.py:85-91): the forward kernel's row side without statistics.
// ------------------------------------------------------------------------------------------
template <typename T, int ROWS, int DCH>
__global__ void __launch_bounds__(kSimThreads) cosine_kernel(SimParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* X = static_cast<const T*>(p.a);
  const T* Y = static_cast<const T*>(p.b);
  const int r0 = blockIdx.x * ROWS;
  float4 xr[ROWS][DCH];
  float inv[ROWS];
  load_resident<T, ROWS, DCH>(X, p.Bq, p.D, r0, lane, xr, inv);
  if (p.inv_a && warp == 0 && lane < ROWS && r0 + lane < p.Bq) {
    float v = inv[0];
#pragma unroll
    for (int r = 1; r < ROWS; ++r) v = (lane == r) ? inv[r] : v;
    p.inv_a[r0 + lane] = v;
  }
  const int mine = owner_row<ROWS>(lane);
  for (int y = warp; y < p.Bk; y += kSimWarps) {
    float dots[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) dots[r] = 0.f;
    float nn = 0.f;
#pragma unroll
    for (int c = 0; c < DCH; ++c) {
      int col = c * 128 + lane * 4;
      float4 yv = (col < p.D) ? ld4(Y + (size_t)y * p.D + col) : make_float4(0, 0, 0, 0);
      nn += dot4(yv, yv);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) dots[r] += dot4(yv, xr[r][c]);
    }
    nn = warp_sum(nn);
    const float invy = 1.f / fmaxf(sqrtf(nn), kEps);
    const float s = warp_multi_sum<ROWS>(dots, lane) * invy;
    if (r0 + mine < p.Bq && (lane & 3) == 0) p.scores[(size_t)(r0 + mine) * p.Bk + y] = s;
    if (p.inv_b && blockIdx.x == 0 && lane == 0) p.inv_b[y] = invy;
  }
}

// ------------------------------------------------------------------------------------------
// Statistics / loss / gradient of the tail over a GIVEN score matrix (word-region scores).
// ------------------------------------------------------------------------------------------
struct TailParams {
  const float* scores; int Bq, Bk;
  const float* labels; int diag; float scale;
  float* row_stats_w; float* col_stats_w;
  const float* row_stats; const float* col_stats;
  const float* row_div; const float* col_div; float num_pos;
  float inv_rows_total, inv_cols_total;
  int col_begin, col_count;
  const float* grad_out;
  float* out;
  int n_row_blocks;
  const int* error_word;   // tail_loss_kernel: non-zero word (a kernel upstream timed out) -> the loss is NaN
  // fused word-score backward (tail_grad_words_kernel)
  const float* rel; const uint8_t* mask; const int* cap_ptr; int T, NQs; float rho2;
};

__global__ void __launch_bounds__(256) tail_stats_kernel(TailParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((int)blockIdx.x < p.n_row_blocks) {            // one warp per row, coalesced along j
    const int i = blockIdx.x * 8 + warp;
    if (i >= p.Bq) return;
    Stat st; st.init();
    const float* row = p.scores + (size_t)i * p.Bk;
    for (int j0 = lane; j0 < p.Bk; j0 += 4 * 32) {   // four loads in flight, consumed in index order
      float z[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) z[u] = (j0 + 32 * u < p.Bk) ? __ldg(row + j0 + 32 * u) : 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j0 + 32 * u < p.Bk) st.add(p.scale * z[u], label_at(p.labels, p.Bk, i, j0 + 32 * u, p.diag));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) st.merge(shfl_xor_stat(st, o));
    if (lane == 0) {
      p.row_stats_w[i] = st.m + logf(st.s);
      p.row_stats_w[p.Bq + i] = st.sl;
      p.row_stats_w[2 * p.Bq + i] = st.slz;
    }
  } else {                                           // 32 columns per CTA, 8 row-strides
    __shared__ Stat sh[8][32];
    const int j = (blockIdx.x - p.n_row_blocks) * 32 + lane;
    Stat st; st.init();
    if (j < p.Bk)
      for (int i0 = warp; i0 < p.Bq; i0 += 4 * 8) {
        float z[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) z[u] = (i0 + 8 * u < p.Bq) ? __ldg(p.scores + (size_t)(i0 + 8 * u) * p.Bk + j) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (i0 + 8 * u < p.Bq) st.add(p.scale * z[u], label_at(p.labels, p.Bk, i0 + 8 * u, j, p.diag));
      }
    sh[warp][lane] = st;
    __syncthreads();
    if (warp == 0 && j < p.Bk) {
      Stat t = sh[0][lane];
#pragma unroll
      for (int w = 1; w < 8; ++w) t.merge(sh[w][lane]);
      p.col_stats_w[j] = t.m + logf(t.s);
      p.col_stats_w[p.Bk + j] = t.sl;
      p.col_stats_w[2 * p.Bk + j] = t.slz;
    }
  }
}

// Column statistics of `world` row shards -> statistics over all rows: log-sum-exp of the shards'
// log-sum-exps, sums of the label sums.  gathered: [world][3][Bk], out: [3][Bk]; one thread per column.
__global__ void __launch_bounds__(256) combine_stats_kernel(const float* __restrict__ g, int world, int Bk,
                                                             float* __restrict__ out) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= Bk) return;
  float m = -INFINITY, sl = 0.f, slz = 0.f;
  for (int w = 0; w < world; ++w) m = fmaxf(m, g[(size_t)w * 3 * Bk + j]);
  float s = 0.f;
  for (int w = 0; w < world; ++w) {
    const float* gw = g + (size_t)w * 3 * Bk;
    s += (m == -INFINITY) ? 0.f : expf(gw[j] - m);
    sl += gw[Bk + j];
    slz += gw[2 * Bk + j];
  }
  out[j] = (m == -INFINITY) ? -INFINITY : m + logf(s);
  out[Bk + j] = sl;
  out[2 * Bk + j] = slz;
}

// Sharded problem, after the packet exchange: gathered[w * stride + ...] = rank w's packet {column statistics over its
// rows [3][Bk], its row-direction partial loss}.  Merges the column statistics (written to col_stats for the backward)
// and evaluates the GLOBAL loss: column direction over all Bk columns from the merged statistics, row direction as the sum
// of the ranks' partials.  Every rank computes the same number, so no all-reduce follows.  Single CTA.
__global__ void __launch_bounds__(1024) combine_loss_kernel(const float* __restrict__ g, int world, int stride, int Bk,
                                                             const float* __restrict__ col_div, float num_pos,
                                                             float inv_cols_total, float* __restrict__ col_stats,
                                                             float* __restrict__ loss_out) {
  float s0 = 0.f;
  for (int j = threadIdx.x; j < Bk; j += 1024) {
    float m = -INFINITY, sl = 0.f, slz = 0.f;
    for (int w = 0; w < world; ++w) m = fmaxf(m, g[(size_t)w * stride + j]);
    float s = 0.f;
    for (int w = 0; w < world; ++w) {
      const float* gw = g + (size_t)w * stride;
      s += (m == -INFINITY) ? 0.f : expf(gw[j] - m);
      sl += gw[Bk + j];
      slz += gw[2 * Bk + j];
    }
    const float lse = (m == -INFINITY) ? -INFINITY : m + logf(s);
    col_stats[j] = lse; col_stats[Bk + j] = sl; col_stats[2 * Bk + j] = slz;
    const float n = col_div ? col_div[j] : num_pos;
    s0 += (lse * sl - slz) / n;
  }
  __shared__ float sh[32];
  s0 = warp_sum(s0);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s0;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 32; ++w) a += sh[w];
    a *= inv_cols_total;
    for (int w = 0; w < world; ++w) b += g[(size_t)w * stride + 3 * Bk];     // NaN if any rank's kernel timed out
    loss_out[0] = a + b; loss_out[1] = a; loss_out[2] = b;
  }
}

// loss_out[0] = s0_part + s1_part, [1] = s0_part (columns), [2] = s1_part (rows); single CTA.
__global__ void __launch_bounds__(256) tail_loss_kernel(TailParams p) {
  float s0 = 0.f, s1 = 0.f;
  for (int i = threadIdx.x; i < p.Bq; i += 256) {
    float n = p.row_div ? p.row_div[i] : p.num_pos;
    s1 += (p.row_stats[i] * p.row_stats[p.Bq + i] - p.row_stats[2 * p.Bq + i]) / n;
  }
  for (int t = threadIdx.x; t < p.col_count; t += 256) {
    int j = p.col_begin + t;
    float n = p.col_div ? p.col_div[j] : p.num_pos;
    s0 += (p.col_stats[j] * p.col_stats[p.Bk + j] - p.col_stats[2 * p.Bk + j]) / n;
  }
  __shared__ float sh0[8], sh1[8];
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) { sh0[threadIdx.x >> 5] = s0; sh1[threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += sh0[w]; b += sh1[w]; }
    a *= p.inv_cols_total; b *= p.inv_rows_total;
    if (p.error_word && *p.error_word != 0) a = b = __int_as_float(0x7fc00000);   // never a plausible number
    p.out[0] = a + b; p.out[1] = a; p.out[2] = b;
  }
}

__global__ void __launch_bounds__(256) tail_grad_kernel(TailParams p) {
  const size_t n = (size_t)p.Bq * p.Bk;
  const float go = __ldg(p.grad_out) * p.scale;
  for (size_t k = (size_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (size_t)gridDim.x * 256) {
    const int i = (int)(k / p.Bk), j = (int)(k % p.Bk);
    const float z = p.scale * p.scores[k];
    const float lab = label_at(p.labels, p.Bk, i, j, p.diag);
    const float inv_nr = p.inv_rows_total / (p.row_div ? p.row_div[i] : p.num_pos);
    const float inv_nc = p.inv_cols_total / (p.col_div ? p.col_div[j] : p.num_pos);
    p.out[k] = go * dscore(z, lab, p.row_stats[i], p.row_stats[p.Bq + i], inv_nr,
                           p.col_stats[j], p.col_stats[p.Bk + j], inv_nc);
  }
}

// tail_grad followed by the backward of the per-caption log-sum-exp, in one pass: thread (i,c) forms
// dS_word(i,c) as tail_grad_kernel does and spreads it over the caption's word rows,
// grel[i, row(c,t)] = dS(i,c) * exp(rho2 * (rel[i,row] - S_word(i,c)))   (0 for padding words).
__global__ void __launch_bounds__(256) tail_grad_words_kernel(TailParams p) {
  const size_t k = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (k >= (size_t)p.Bq * p.Bk) return;
  const int i = (int)(k / p.Bk), c = (int)(k % p.Bk);
  const float sc = p.scores[k];
  const float go = __ldg(p.grad_out) * p.scale;
  const float inv_nr = p.inv_rows_total / (p.row_div ? p.row_div[i] : p.num_pos);
  const float inv_nc = p.inv_cols_total / (p.col_div ? p.col_div[c] : p.num_pos);
  const float ds = go * dscore(p.scale * sc, label_at(p.labels, p.Bk, i, c, p.diag), p.row_stats[i], p.row_stats[p.Bq + i],
                               inv_nr, p.col_stats[c], p.col_stats[p.Bk + c], inv_nc);
  const int lo = p.cap_ptr ? p.cap_ptr[c] : c * p.T, hi = p.cap_ptr ? p.cap_ptr[c + 1] : c * p.T + p.T;
  const uint8_t* mk = (p.mask && !p.cap_ptr) ? p.mask : nullptr;
  const float* r = p.rel + (size_t)i * p.NQs;
  float* g = p.out + (size_t)i * p.NQs;
  for (int q0 = lo; q0 < hi; q0 += 8) {            // eight loads in flight
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (q0 + u < hi) ? __ldg(r + q0 + u) : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (q0 + u >= hi) break;
      const bool pad = mk && mk[q0 + u];
      g[q0 + u] = pad ? 0.f : ds * __expf(p.rho2 * (v[u] - sc));
    }
  }
}

// ------------------------------------------------------------------------------------------
// make_labels soft positives (train_gan.py:72-83).  One warp per row.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) labels_count_kernel(const float* sim, int B, float p, float* count) {
  const int lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= B) return;
  float c = 0.f;
  for (int j = lane; j < B; j += 32) c += (j != i && sim[(size_t)i * B + j] > p) ? 1.f : 0.f;
  c = warp_sum(c);
  if (lane == 0) count[i] = c;
}
__global__ void __launch_bounds__(256) labels_fill_kernel(const float* sim, int B, float p, float smooth,
                                                          const float* count, float* labels, float* row_count) {
  const int lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= B) return;
  float c = 0.f;
  for (int j = lane; j < B; j += 32) {
    const bool pos = (j != i) && sim[(size_t)i * B + j] > p;
    const float w = (smooth != 0.f) ? smooth : 1.f / (fmaxf(count[j], 1.f) + 1.f);   // :79-81, by COLUMN (:82)
    const float v = fminf((j == i ? 1.f : 0.f) + (pos ? w : 0.f), 1.f);
    labels[(size_t)i * B + j] = v;
    c += v > 0.f ? 1.f : 0.f;
  }
  c = warp_sum(c);
  if (lane == 0) row_count[i] = c;
}

// ------------------------------------------------------------------------------------------
// Host-side dispatch
// ------------------------------------------------------------------------------------------
template <typename T, int ROWS, int DCH>
static int launch_sim(int which, const SimParams& p, cudaStream_t st) {
  if (which == 0) {
    int grid = p.n_row_blocks + (p.Bk + ROWS - 1) / ROWS;
    sim_fwd_kernel<T, ROWS, DCH><<<grid, kSimThreads, 0, st>>>(p);
  } else if (which == 1) {
    SimParams q = p;
    q.n_a_blocks = p.da ? (p.Bq + ROWS - 1) / ROWS : 0;
    int grid = q.n_a_blocks + (p.db ? (p.Bk + ROWS - 1) / ROWS : 0);
    if (grid > 0) sim_bwd_kernel<T, ROWS, DCH><<<grid, kSimThreads, 0, st>>>(q);
  } else {
    cosine_kernel<T, ROWS, DCH><<<p.n_row_blocks, kSimThreads, 0, st>>>(p);
  }
  return cuda_fail(cudaGetLastError(), "similarity-loss kernel launch");
}

template <typename T, int DCH>
static int dispatch_rows(int which, SimParams p, int max_rows, cudaStream_t st) {
  // ROWS resident vectors per CTA.  The problem is small and latency-bound: take the largest ROWS
  // (fewest passes over the streamed side) that still gives three quarters of the SMs a CTA, else the smallest.
  auto ctas = [&](int r) { return (p.Bq + r - 1) / r + (p.Bk + r - 1) / r; };
  int rows = 2;
  for (int r = max_rows; r >= 2; r >>= 1)
    if (ctas(r) >= 112) { rows = r; break; }
  p.n_row_blocks = (p.Bq + rows - 1) / rows;
  if (rows == 8) { if constexpr (DCH <= 2) return launch_sim<T, 8, DCH>(which, p, st); }
  if (rows == 4) return launch_sim<T, 4, DCH>(which, p, st);
  return launch_sim<T, 2, DCH>(which, p, st);
}

template <typename T>
static int dispatch_dch(int which, SimParams p, cudaStream_t st) {
  const int dch = (p.D + 127) / 128;
  // at most 8 resident vectors while they fit in 64 registers, else 4
  if (dch == 1) return dispatch_rows<T, 1>(which, p, 8, st);
  if (dch == 2) return dispatch_rows<T, 2>(which, p, 8, st);
  if (dch <= 4) return dispatch_rows<T, 4>(which, p, 4, st);
  if (dch <= 6) return dispatch_rows<T, 6>(which, p, 4, st);
  set_error("D=%d unsupported (max 768)", p.D);
  return XMC_ERR_UNSUPPORTED;
}

static int check_sim_args(const void* a, const void* b, int Bq, int Bk, int D, int dtype) {
  XMC_REQUIRE(a && b, XMC_ERR_INVALID_ARG, "null embedding pointer");
  XMC_REQUIRE(Bq > 0 && Bk > 0 && D > 0, XMC_ERR_INVALID_ARG, "bad shape Bq=%d Bk=%d D=%d", Bq, Bk, D);
  XMC_REQUIRE(D % 4 == 0 && D <= 768, XMC_ERR_UNSUPPORTED, "D=%d must be a multiple of 4 and <= 768", D);
  XMC_REQUIRE(dtype == XMC_F32 || dtype == XMC_BF16, XMC_ERR_UNSUPPORTED, "dtype %d", dtype);
  XMC_REQUIRE(aligned16(a) && aligned16(b), XMC_ERR_ALIGNMENT, "embedding pointers must be 16-byte aligned");
  return XMC_OK;
}

int sim_dispatch(int which, const SimParams& p, int dtype, cudaStream_t st) {
  return dtype == XMC_F32 ? dispatch_dch<float>(which, p, st) : dispatch_dch<__nv_bfloat16>(which, p, st);
}

}  // namespace xmc

using namespace xmc;

extern "C" int xmc_cosine_scores(const void* a, const void* b, int Bq, int Bk, int D, int dtype,
                                 float* scores, float* inv_norm_a, float* inv_norm_b, void* stream) {
  if (int rc = check_sim_args(a, b, Bq, Bk, D, dtype)) return rc;
  XMC_REQUIRE(scores, XMC_ERR_INVALID_ARG, "null scores");
  SimParams p{};
  p.a = a; p.b = b; p.Bq = Bq; p.Bk = Bk; p.D = D;
  p.scores = scores; p.inv_a = inv_norm_a; p.inv_b = inv_norm_b;
  if (sim_tc_eligible(Bq, Bk, D)) return sim_tc_forward(p, dtype, as_stream(stream));
  return sim_dispatch(2, p, dtype, as_stream(stream));
}

extern "C" int xmc_cosine_scores_backward(const void* a, const void* b, int Bq, int Bk, int D, int dtype,
                                          const float* inv_norm_a, const float* inv_norm_b, const float* dscores,
                                          void* da, void* db, void* stream) {
  if (int rc = check_sim_args(a, b, Bq, Bk, D, dtype)) return rc;
  XMC_REQUIRE(inv_norm_a && inv_norm_b && dscores, XMC_ERR_INVALID_ARG, "null pointer");
  SimParams p{};
  p.a = a; p.b = b; p.Bq = Bq; p.Bk = Bk; p.D = D;
  p.inv_a = const_cast<float*>(inv_norm_a); p.inv_b = const_cast<float*>(inv_norm_b);
  p.ds_given = dscores; p.da = da; p.db = db;
  return sim_dispatch(1, p, dtype, as_stream(stream));
}

extern "C" int xmc_simloss_forward(const void* a, const void* b, int Bq, int Bk, int D, int dtype,
                                   const float* labels, int diag_offset, float scale,
                                   float* scores, float* inv_norm_a, float* inv_norm_b,
                                   float* row_stats, float* col_stats, void* stream) {
  if (int rc = check_sim_args(a, b, Bq, Bk, D, dtype)) return rc;
  XMC_REQUIRE(scores && row_stats && col_stats, XMC_ERR_INVALID_ARG, "null output pointer");
  SimParams p{};
  p.a = a; p.b = b; p.Bq = Bq; p.Bk = Bk; p.D = D;
  p.labels = labels; p.diag = diag_offset; p.scale = scale;
  p.scores = scores; p.inv_a = inv_norm_a; p.inv_b = inv_norm_b;
  p.row_stats = row_stats; p.col_stats = col_stats;
  if (sim_tc_eligible(Bq, Bk, D)) {
    // large rectangular problem (global negatives): score tiles on the tensor cores, then one pass over the scores
    // for both directions' statistics (the kernel the word-region scores use)
    if (int rc = sim_tc_forward(p, dtype, as_stream(stream))) return rc;
    return xmc_infonce_stats(scores, Bq, Bk, labels, diag_offset, scale, row_stats, col_stats, stream);
  }
  return sim_dispatch(0, p, dtype, as_stream(stream));
}

extern "C" int xmc_simloss_backward(const void* a, const void* b, int Bq, int Bk, int D, int dtype,
                                    const float* scores, const float* inv_norm_a, const float* inv_norm_b,
                                    const float* labels, int diag_offset, float scale,
                                    const float* row_stats, const float* col_stats,
                                    const float* row_div, const float* col_div, float num_pos,
                                    int rows_total, int cols_total, const float* grad_out,
                                    void* da, void* db, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_sim_args(a, b, Bq, Bk, D, dtype)) return rc;
  XMC_REQUIRE(scores && inv_norm_a && inv_norm_b && row_stats && col_stats && grad_out,
              XMC_ERR_INVALID_ARG, "null saved-state pointer");
  XMC_REQUIRE(rows_total > 0 && cols_total > 0 && num_pos > 0.f, XMC_ERR_INVALID_ARG, "bad totals / num_pos");
  XMC_REQUIRE((!da || aligned16(da)) && (!db || aligned16(db)), XMC_ERR_ALIGNMENT, "gradient pointers must be 16-byte aligned");
  SimParams p{};
  p.a = a; p.b = b; p.Bq = Bq; p.Bk = Bk; p.D = D;
  p.labels = labels; p.diag = diag_offset; p.scale = scale;
  p.scores = const_cast<float*>(scores);
  p.inv_a = const_cast<float*>(inv_norm_a); p.inv_b = const_cast<float*>(inv_norm_b);
  p.row_stats = const_cast<float*>(row_stats); p.col_stats = const_cast<float*>(col_stats);
  p.row_div = row_div; p.col_div = col_div; p.num_pos = num_pos;
  p.inv_rows_total = 1.f / rows_total; p.inv_cols_total = 1.f / cols_total;
  p.grad_out = grad_out; p.da = da; p.db = db;
  if (workspace && sim_tc_eligible(Bq, Bk, D)) return sim_tc_backward(p, dtype, workspace, workspace_bytes, as_stream(stream));
  return sim_dispatch(1, p, dtype, as_stream(stream));
}

extern "C" size_t xmc_simloss_workspace_bytes(int Bq, int Bk, int D) { return sim_tc_workspace_bytes(Bq, Bk, D); }

static int check_tail(const float* scores, int Bq, int Bk) {
  XMC_REQUIRE(scores, XMC_ERR_INVALID_ARG, "null scores");
  XMC_REQUIRE(Bq > 0 && Bk > 0, XMC_ERR_INVALID_ARG, "bad shape Bq=%d Bk=%d", Bq, Bk);
  return XMC_OK;
}

extern "C" int xmc_infonce_stats(const float* scores, int Bq, int Bk, const float* labels, int diag_offset,
                                 float scale, float* row_stats, float* col_stats, void* stream) {
  if (int rc = check_tail(scores, Bq, Bk)) return rc;
  XMC_REQUIRE(row_stats && col_stats, XMC_ERR_INVALID_ARG, "null statistics pointer");
  TailParams p{};
  p.scores = scores; p.Bq = Bq; p.Bk = Bk; p.labels = labels; p.diag = diag_offset; p.scale = scale;
  p.row_stats_w = row_stats; p.col_stats_w = col_stats;
  p.n_row_blocks = (Bq + 7) / 8;
  int grid = p.n_row_blocks + (Bk + 31) / 32;
  tail_stats_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
  return cuda_fail(cudaGetLastError(), "tail_stats_kernel launch");
}

extern "C" int xmc_infonce_combine_stats(const float* gathered, int world, int Bk, float* col_stats, void* stream) {
  XMC_REQUIRE(gathered && col_stats, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(world > 0 && Bk > 0, XMC_ERR_INVALID_ARG, "bad sizes world=%d Bk=%d", world, Bk);
  combine_stats_kernel<<<(Bk + 255) / 256, 256, 0, as_stream(stream)>>>(gathered, world, Bk, col_stats);
  return cuda_fail(cudaGetLastError(), "combine_stats_kernel launch");
}

extern "C" int xmc_infonce_combine_loss(const float* gathered, int world, int stride, int Bk, const float* col_div,
                                        float num_pos, int cols_total, float* col_stats, float* loss_out, void* stream) {
  XMC_REQUIRE(gathered && col_stats && loss_out, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(world > 0 && Bk > 0 && stride >= 3 * Bk + 1 && cols_total > 0 && num_pos > 0.f, XMC_ERR_INVALID_ARG,
              "bad sizes world=%d Bk=%d stride=%d", world, Bk, stride);
  combine_loss_kernel<<<1, 1024, 0, as_stream(stream)>>>(gathered, world, stride, Bk, col_div, num_pos, 1.f / cols_total,
                                                          col_stats, loss_out);
  return cuda_fail(cudaGetLastError(), "combine_loss_kernel launch");
}

extern "C" int xmc_infonce_loss(const float* row_stats, const float* col_stats, int Bq, int Bk,
                                const float* row_div, const float* col_div, float num_pos,
                                int rows_total, int cols_total, int col_begin, int col_count,
                                float* loss_out, const int* error_word, void* stream) {
  XMC_REQUIRE(row_stats && col_stats && loss_out, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(Bq > 0 && Bk > 0 && rows_total > 0 && cols_total > 0 && num_pos > 0.f, XMC_ERR_INVALID_ARG, "bad sizes");
  XMC_REQUIRE(col_begin >= 0 && col_count >= 0 && col_begin + col_count <= Bk, XMC_ERR_INVALID_ARG, "bad column range");
  TailParams p{};
  p.Bq = Bq; p.Bk = Bk; p.row_stats = row_stats; p.col_stats = col_stats;
  p.row_div = row_div; p.col_div = col_div; p.num_pos = num_pos;
  p.inv_rows_total = 1.f / rows_total; p.inv_cols_total = 1.f / cols_total;
  p.col_begin = col_begin; p.col_count = col_count; p.out = loss_out; p.error_word = error_word;
  tail_loss_kernel<<<1, 256, 0, as_stream(stream)>>>(p);
  return cuda_fail(cudaGetLastError(), "tail_loss_kernel launch");
}

extern "C" int xmc_infonce_grad(const float* scores, int Bq, int Bk, const float* labels, int diag_offset,
                                float scale, const float* row_stats, const float* col_stats,
                                const float* row_div, const float* col_div, float num_pos,
                                int rows_total, int cols_total, const float* grad_out,
                                float* dscores, void* stream) {
  if (int rc = check_tail(scores, Bq, Bk)) return rc;
  XMC_REQUIRE(row_stats && col_stats && grad_out && dscores, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(rows_total > 0 && cols_total > 0 && num_pos > 0.f, XMC_ERR_INVALID_ARG, "bad totals / num_pos");
  TailParams p{};
  p.scores = scores; p.Bq = Bq; p.Bk = Bk; p.labels = labels; p.diag = diag_offset; p.scale = scale;
  p.row_stats = row_stats; p.col_stats = col_stats;
  p.row_div = row_div; p.col_div = col_div; p.num_pos = num_pos;
  p.inv_rows_total = 1.f / rows_total; p.inv_cols_total = 1.f / cols_total;
  p.grad_out = grad_out; p.out = dscores;
  size_t n = (size_t)Bq * Bk;
  int grid = (int)((n + 255) / 256); if (grid > 148 * 8) grid = 148 * 8;
  tail_grad_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
  return cuda_fail(cudaGetLastError(), "tail_grad_kernel launch");
}

extern "C" int xmc_word_scores_infonce_backward(const float* rel, const uint8_t* mask, const int* cap_ptr,
                                                const float* scores, int Bi, int Bc, int T, int NQs, float rho2,
                                                const float* labels, int diag_offset, float scale,
                                                const float* row_stats, const float* col_stats,
                                                const float* row_div, const float* col_div, float num_pos,
                                                int rows_total, int cols_total, const float* grad_out,
                                                float* grel, void* stream) {
  if (int rc = check_tail(scores, Bi, Bc)) return rc;
  XMC_REQUIRE(rel && row_stats && col_stats && grad_out && grel, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(T > 0 && NQs >= Bc && rho2 > 0.f, XMC_ERR_INVALID_ARG, "bad shape / rho2");
  XMC_REQUIRE(rows_total > 0 && cols_total > 0 && num_pos > 0.f, XMC_ERR_INVALID_ARG, "bad totals / num_pos");
  TailParams p{};
  p.scores = scores; p.Bq = Bi; p.Bk = Bc; p.labels = labels; p.diag = diag_offset; p.scale = scale;
  p.row_stats = row_stats; p.col_stats = col_stats;
  p.row_div = row_div; p.col_div = col_div; p.num_pos = num_pos;
  p.inv_rows_total = 1.f / rows_total; p.inv_cols_total = 1.f / cols_total;
  p.grad_out = grad_out; p.out = grel;
  p.rel = rel; p.mask = mask; p.cap_ptr = cap_ptr; p.T = T; p.NQs = NQs; p.rho2 = rho2;
  const size_t n = (size_t)Bi * Bc;
  tail_grad_words_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(p);
  return cuda_fail(cudaGetLastError(), "tail_grad_words_kernel launch");
}

extern "C" int xmc_make_labels(const float* sim, int B, float p, float smooth_global,
                               float* labels, float* row_count, float* tmp_count, void* stream) {
  XMC_REQUIRE(sim && labels && row_count && tmp_count, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE(B > 0, XMC_ERR_INVALID_ARG, "bad B=%d", B);
  int grid = (B + 7) / 8;
  labels_count_kernel<<<grid, 256, 0, as_stream(stream)>>>(sim, B, p, tmp_count);
  labels_fill_kernel<<<grid, 256, 0, as_stream(stream)>>>(sim, B, p, smooth_global, tmp_count, labels, row_count);
  return cuda_fail(cudaGetLastError(), "make_labels kernels launch");
}
