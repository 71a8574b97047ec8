// api.cu — C-ABI glue of libxmcloss: error plumbing, version/device checks and the dispatch of the
// word-region entry points to the fp32 CUDA-core path or the bf16 tcgen05 path.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "wordregion.h"

namespace xmc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return XMC_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return XMC_ERR_CUDA;
}

static int check_wr(int path, const void* qn, const void* kn, int NQ, int Bi, int R, int Rpad, int D) {
  XMC_REQUIRE(path == XMC_PATH_FP32_SIMT || path == XMC_PATH_BF16_TCGEN05 || path == XMC_PATH_FP32_TCGEN05, XMC_ERR_UNSUPPORTED,
              "unknown path %d", path);
  XMC_REQUIRE(path != XMC_PATH_FP32_TCGEN05 || D == 256, XMC_ERR_UNSUPPORTED, "the split-bf16 path supports D = 256 (got %d)", D);
  XMC_REQUIRE(qn && kn, XMC_ERR_INVALID_ARG, "null operand pointer");
  XMC_REQUIRE(NQ > 0 && Bi > 0 && R > 0, XMC_ERR_INVALID_ARG, "bad shape NQ=%d Bi=%d R=%d", NQ, Bi, R);
  XMC_REQUIRE(Rpad >= R && Rpad % 16 == 0, XMC_ERR_INVALID_ARG, "Rpad=%d must be a multiple of 16 and >= R=%d", Rpad, R);
  XMC_REQUIRE(D == 64 || D == 128 || D == 256, XMC_ERR_UNSUPPORTED, "word-region D=%d unsupported (64, 128, 256)", D);
  XMC_REQUIRE(aligned16(qn) && aligned16(kn), XMC_ERR_ALIGNMENT, "operand pointers must be 16-byte aligned");
  return XMC_OK;
}

}  // namespace xmc

using namespace xmc;

extern "C" int xmc_version(void) { return XMC_ABI_VERSION; }
extern "C" const char* xmc_last_error(void) { return g_err; }

extern "C" int xmc_check_device(void) {
  int dev = 0;
  XMC_RETURN_IF_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  XMC_RETURN_IF_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  XMC_RETURN_IF_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  XMC_REQUIRE(major == 10, XMC_ERR_UNSUPPORTED, "device is sm_%d%d; libxmcloss is built for sm_100a only", major, minor);
  return XMC_OK;
}

extern "C" size_t xmc_wordregion_workspace_bytes(int path, int NQ, int Bi, int R, int Rpad, int D) {
  if (path == XMC_PATH_BF16_TCGEN05) return wordregion_tc_workspace_bytes(NQ, Bi, R, Rpad, D);
  if (path == XMC_PATH_FP32_TCGEN05) return wordregion_split_workspace_bytes(NQ, Bi, R, Rpad, D);
  return 0;
}

extern "C" int xmc_wordregion_forward(int path, const void* qn, const void* kn, const float* rnorm,
                                      int NQ, int Bi, int R, int Rpad, int D, float rho1,
                                      float* lsum, float* cnorm, float* rel, void* chat, const int* nq_dev,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_wr(path, qn, kn, NQ, Bi, R, Rpad, D)) return rc;
  XMC_REQUIRE(lsum && cnorm && rel, XMC_ERR_INVALID_ARG, "null statistics pointer");
  XMC_REQUIRE(rho1 > 0.f, XMC_ERR_INVALID_ARG, "rho1 must be positive");
  WrParams p{};
  p.qn = static_cast<const float*>(qn); p.kn = static_cast<const float*>(kn); p.rnorm = rnorm;
  p.NQ = NQ; p.Bi = Bi; p.R = R; p.Rpad = Rpad; p.rho1 = rho1;
  p.lsum = lsum; p.cnorm = cnorm; p.rel = rel; p.chat = chat; p.nq_dev = nq_dev;
  XMC_REQUIRE(!chat || aligned16(chat), XMC_ERR_ALIGNMENT, "chat must be 16-byte aligned");
  if (path == XMC_PATH_FP32_SIMT) return wordregion_f32_forward(p, D, as_stream(stream));
  if (path == XMC_PATH_FP32_TCGEN05) return wordregion_split_forward(p, D, workspace, workspace_bytes, as_stream(stream));
  return wordregion_tc_forward(p, D, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int xmc_wordregion_backward(int path, const void* qn, const void* kn, const float* rnorm,
                                       int NQ, int Bi, int R, int Rpad, int D, float rho1,
                                       const float* lsum, const float* cnorm, const float* rel, const void* chat,
                                       const float* grel, float* dqn, float* dkn, float* drnorm, const int* nq_dev,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_wr(path, qn, kn, NQ, Bi, R, Rpad, D)) return rc;
  XMC_REQUIRE(lsum && cnorm && rel && grel && dqn && dkn, XMC_ERR_INVALID_ARG, "null pointer");
  XMC_REQUIRE((rnorm == nullptr) == (drnorm == nullptr), XMC_ERR_INVALID_ARG, "drnorm must be given iff rnorm is");
  XMC_REQUIRE(rho1 > 0.f, XMC_ERR_INVALID_ARG, "rho1 must be positive");
  WrParams p{};
  p.qn = static_cast<const float*>(qn); p.kn = static_cast<const float*>(kn); p.rnorm = rnorm;
  p.NQ = NQ; p.Bi = Bi; p.R = R; p.Rpad = Rpad; p.rho1 = rho1;
  p.lsum = const_cast<float*>(lsum); p.cnorm = const_cast<float*>(cnorm); p.rel = const_cast<float*>(rel);
  p.grel = grel; p.dqn = dqn; p.dkn = dkn; p.drnorm = drnorm; p.chat = const_cast<void*>(chat); p.nq_dev = nq_dev;
  if (path == XMC_PATH_FP32_SIMT) return wordregion_f32_backward(p, D, as_stream(stream));
  if (path == XMC_PATH_FP32_TCGEN05) return wordregion_split_backward(p, D, workspace, workspace_bytes, as_stream(stream));
  return wordregion_tc_backward(p, D, workspace, workspace_bytes, as_stream(stream));
}
