// Error reporting and version of the apnerf C-ABI library.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void apnerf_set_error(const char* where, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}
void apnerf_set_error_msg(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }

APNERF_API const char* apnerf_last_error(void) { return g_err; }
APNERF_API int apnerf_abi_version(void) { return 1; }
