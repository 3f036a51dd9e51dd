// Status strings, error bookkeeping, device queries.
#include <string.h>
#include <stdio.h>
#include "common.cuh"

namespace pe {
static thread_local char g_last_error[512] = "";

void set_last_error(cudaError_t e, const char* where) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

int sm_count() {
  static int cached[128] = {0};  // per device ordinal; a benign race writes the same value twice
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) return 148;
  if (cached[dev] == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached[dev] = n;
    else return 148;
  }
  return cached[dev];
}
}  // namespace pe

extern "C" PE_API const char* pe_status_string(int status) {
  switch (status) {
    case PE_OK: return "PE_OK";
    case PE_ERR_INVALID_ARGUMENT: return "PE_ERR_INVALID_ARGUMENT";
    case PE_ERR_UNSUPPORTED: return "PE_ERR_UNSUPPORTED";
    case PE_ERR_CUDA: return "PE_ERR_CUDA";
    case PE_ERR_WORKSPACE_TOO_SMALL: return "PE_ERR_WORKSPACE_TOO_SMALL";
    case PE_ERR_NOT_INITIALISED: return "PE_ERR_NOT_INITIALISED";
    default: return "PE_ERR_UNKNOWN";
  }
}

extern "C" PE_API int pe_abi_version(void) { return 1; }

extern "C" PE_API const char* pe_last_error_string(void) { return pe::g_last_error; }
