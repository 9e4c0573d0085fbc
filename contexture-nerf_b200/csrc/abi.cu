// ABI bookkeeping entry points of libctxnerf.so (see include/ctxnerf.h).
#include "ctx_common.cuh"
#include "ctxnerf.h"

extern "C" int ctx_abi_version(void) { return CTXNERF_ABI_VERSION; }

extern "C" const char* ctx_error_string(int code) {
  if (code == 0) return "success";
  if (code == CTX_ERR_BAD_ARG) return "ctxnerf: bad argument (null pointer, negative size or size out of range)";
  if (code == CTX_ERR_UNSUPPORTED) return "ctxnerf: unsupported configuration for this kernel";
  if (code == CTX_ERR_NO_NCCL) return "ctxnerf: NCCL not loaded or an NCCL call failed (see ctx_comm_last_error)";
  if (code < 0) return "ctxnerf: unknown argument error";
  return cudaGetErrorString((cudaError_t)code);
}
