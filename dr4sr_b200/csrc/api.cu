// api.cu -- ABI version and error plumbing of libdr4sr.
#include <stdio.h>
#include <string.h>
#include "common.cuh"

namespace dr4sr {
static thread_local char g_err[256] = "";
void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}
}  // namespace dr4sr

extern "C" int dr4sr_abi_version(void) { return DR4SR_ABI_VERSION; }
extern "C" const char* dr4sr_last_cuda_error(void) { return dr4sr::g_err; }
