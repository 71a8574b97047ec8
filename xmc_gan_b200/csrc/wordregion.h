// wordregion.h — internal interface between api.cu and the two word-region paths.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace xmc {

// Operands are unit rows: qn[NQ,D], kn[Bi,Rpad,D]; storage fp32 (SIMT path) or bf16 (tcgen05 path).
struct WrParams {
  const void* qn; const void* kn; const float* rnorm;
  int NQ, Bi, R, Rpad;                     // NQ: rows allocated (row stride of lsum / cnorm / rel / chat)
  const int* nq_dev;                       // device count of valid word rows (compacted captions), nullable: all NQ
  float rho1;
  float* lsum; float* cnorm; float* rel;   // forward outputs / backward inputs, [Bi, NQ]
  void* chat;                              // [Bi, NQ, D] bf16 context sums C = l c_t (tcgen05 paths; the split path: hi plane then lo plane), nullable
  const float* grel;                       // [Bi, NQ]
  float* dqn; float* dkn; float* drnorm;   // fp32, accumulated into
};

int wordregion_f32_forward(const WrParams& p, int D, cudaStream_t st);
int wordregion_f32_backward(const WrParams& p, int D, cudaStream_t st);

size_t wordregion_tc_workspace_bytes(int NQ, int Bi, int R, int Rpad, int D);
int wordregion_tc_forward(const WrParams& p, int D, void* ws, size_t ws_bytes, cudaStream_t st);
int wordregion_tc_backward(const WrParams& p, int D, void* ws, size_t ws_bytes, cudaStream_t st);

// fp32-tolerance path on the tensor cores (wordregion_split.cu): fp32 operands carried as hi/lo bf16 pairs, D = 256
size_t wordregion_split_workspace_bytes(int NQ, int Bi, int R, int Rpad, int D);
int wordregion_split_forward(const WrParams& p, int D, void* ws, size_t ws_bytes, cudaStream_t st);
int wordregion_split_backward(const WrParams& p, int D, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace xmc
