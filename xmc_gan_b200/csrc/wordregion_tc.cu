// wordregion_tc.cu — bf16 tcgen05/TMEM/TMA path of the word-region loss (placeholder until built).
#include "common.cuh"
#include "wordregion.h"
namespace xmc {
size_t wordregion_tc_workspace_bytes(int, int, int, int, int) { return 0; }
int wordregion_tc_forward(const WrParams&, int, void*, size_t, cudaStream_t) {
  set_error("tcgen05 word-region forward not built yet");
  return XMC_ERR_UNSUPPORTED;
}
int wordregion_tc_backward(const WrParams&, int, void*, size_t, cudaStream_t) {
  set_error("tcgen05 word-region backward not built yet");
  return XMC_ERR_UNSUPPORTED;
}
}  // namespace xmc
