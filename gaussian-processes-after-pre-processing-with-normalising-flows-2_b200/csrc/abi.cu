// ABI bookkeeping for libflowk.so: version, status strings, workspace sizing.
#include "common.cuh"

extern "C" int flowk_abi_version(void) { return 3; }   // 2: flowk_conv_gemm_args grew (operand_format, acc_scale, dilation, acc_scale2); 3: + acc_scale_ptr, flowk_wn_job.fwd_f16

extern "C" size_t flowk_ldj_workspace_bytes(int B) {
  if (B < 1) B = 1;
  return (size_t)B * (flowk::kMaxParts + 1) * sizeof(float);
}

extern "C" const char* flowk_error_string(int status) {
  switch (status) {
    case FLOWK_OK: return "ok";
    case FLOWK_ERR_SHAPE: return "bad shape";
    case FLOWK_ERR_ALIGN: return "misaligned pointer";
    case FLOWK_ERR_ARG: return "bad argument";
    default: break;
  }
  if (status >= FLOWK_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(status - FLOWK_ERR_CUDA_BASE));
  return "unknown flowk status";
}
