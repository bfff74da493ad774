// Version / error / device-info entry points of libdamsm_b200.so.
#include <stdarg.h>
#include "common.cuh"

namespace damsm {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 3;
  }
  return 0;
}
}  // namespace damsm

extern "C" int damsm_version(void) { return DAMSM_ABI_VERSION; }
extern "C" const char *damsm_last_error(void) { return damsm::g_err; }

extern "C" int damsm_device_info(int *sm_count, int *cc_major, int *cc_minor, int *max_smem_optin) {
  int dev = 0;
  DAMSM_CUDA(cudaGetDevice(&dev));
  if (sm_count) DAMSM_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) DAMSM_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) DAMSM_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (max_smem_optin) DAMSM_CUDA(cudaDeviceGetAttribute(max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return 0;
}
