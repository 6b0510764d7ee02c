// Error reporting + device probing for libcaro_b200.so.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/caro_b200.h"
#include "common_host.h"

static thread_local char g_err[512] = "";

int caro_fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "error");
  return code;
}

int caro_check_launch(const char* what) {
  const cudaError_t ce = cudaGetLastError();
  if (ce == cudaSuccess) return CARO_OK;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(ce));
  return CARO_E_CUDA;
}

extern "C" {

int caro_abi_version(void) { return CARO_ABI_VERSION; }

const char* caro_last_error(void) { return g_err; }

int caro_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();  // clear the sticky "no device" error
    return 0;
  }
  return n;
}

}  // extern "C"
