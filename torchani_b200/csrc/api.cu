// Small C-ABI utilities: version, error strings, last CUDA error.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace ani {
static thread_local char g_last_cuda_error[256] = "no error";
void set_cuda_error(cudaError_t e) {
  const char* name = cudaGetErrorName(e);
  const char* msg = cudaGetErrorString(e);
  snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s", name ? name : "?", msg ? msg : "?");
}
}  // namespace ani

extern "C" int ani_b200_abi_version(void) { return ANI_B200_ABI_VERSION; }

extern "C" int ani_b200_operand_format(int32_t* parts, float* value_scale, float* grad_scale) {
  if (parts) *parts = ani::OPND_PARTS;
  if (value_scale) *value_scale = ani::OPND_SCALE_VALUE;
  if (grad_scale) *grad_scale = ani::OPND_SCALE_GRAD;
  return ANI_OK;
}

extern "C" const char* ani_b200_error_string(int code) {
  switch (code) {
    case ANI_OK: return "ok";
    case ANI_ERR_BAD_ARG: return "bad argument (null pointer, size or alignment)";
    case ANI_ERR_UNSUPPORTED: return "configuration not supported by the B200 kernels";
    case ANI_ERR_CUDA: return "CUDA error (see ani_b200_last_cuda_error)";
    default: return "unknown error code";
  }
}

extern "C" const char* ani_b200_last_cuda_error(void) { return ani::g_last_cuda_error; }
