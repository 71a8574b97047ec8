// simloss.cuh — pieces shared by the CUDA-core similarity-loss kernels (simloss.cu) and their tcgen05
// counterparts for large, rectangular problems (simloss_tc.cu).
#pragma once
#include "common.cuh"

namespace xmc {

// Label of (row i, col j): explicit matrix or identity with column offset.
__device__ __forceinline__ float label_at(const float* labels, int Bk, int i, int j, int diag) {
  return labels ? __ldg(labels + (size_t)i * Bk + j) : (j == i + diag ? 1.f : 0.f);
}

struct SimParams {
  const void* a; const void* b;
  int Bq, Bk, D;
  const float* labels; int diag; float scale;
  float* scores; float* inv_a; float* inv_b;
  float* row_stats; float* col_stats;
  int n_row_blocks;
  // backward only
  const float* row_div; const float* col_div; float num_pos;
  float inv_rows_total, inv_cols_total;
  const float* grad_out;
  void* da; void* db;
  int n_a_blocks;
  const float* ds_given;   // cosine_scores backward: d loss / d scores comes from the caller instead of the closed form
};

// d loss / d score(i,j) without the grad_out*scale factor.
__device__ __forceinline__ float dscore(float z, float lab, float row_lse, float row_sl, float inv_nr,
                                        float col_lse, float col_sl, float inv_nc) {
  float pr = __expf(z - row_lse), pc = __expf(z - col_lse);
  return (pc * col_sl - lab) * inv_nc + (pr * row_sl - lab) * inv_nr;
}


// tcgen05 path (simloss_tc.cu): taken for large problems (global negatives: Bq x Bk = 256 x 2048), where the
// CUDA-core kernels, built for 256 x 256 latency, take 100-200 us.
bool sim_tc_eligible(int Bq, int Bk, int D);
size_t sim_tc_workspace_bytes(int Bq, int Bk, int D);
int sim_tc_forward(const SimParams& p, int dtype, cudaStream_t st);                       // scores + inverse norms
int sim_tc_backward(const SimParams& p, int dtype, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace xmc
