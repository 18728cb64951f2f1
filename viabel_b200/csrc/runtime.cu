// Error reporting and device queries shared by every translation unit.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace vb {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int set_cuda_error(cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d: %s", (int)e, cudaGetErrorString(e));
  return VB_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace vb

extern "C" const char* vb_last_error(void) { return vb::g_err; }
extern "C" int vb_version(void) { return 100; }
extern "C" int vb_device_sm_count(void) { return vb::sm_count(); }
