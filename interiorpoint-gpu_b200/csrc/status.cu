// Library status / error plumbing.
#include <string.h>

#include "common.cuh"

static thread_local cudaError_t g_last_cuda = cudaSuccess;

extern "C" int ipm_set_cuda_error(cudaError_t e) {
  g_last_cuda = e;
  return IPM_ERR_CUDA;
}
extern "C" const char* ipm_last_cuda_error(void) { return cudaGetErrorString(g_last_cuda); }
extern "C" int ipm_abi_version(void) { return 1; }
extern "C" int ipm_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return IPM_ERR_NO_DEVICE;
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return IPM_ERR_NO_DEVICE;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  return major == 10 ? IPM_OK : IPM_ERR_NO_DEVICE;
}

// Number of kernels this library has launched in this process (bench.py reports the delta over the timed region).
static unsigned long long g_launches = 0;
extern "C" void ipm_count_launch(void) { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
extern "C" unsigned long long ipm_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// L2 residency for iteration state that is re-read every launch (the ADMM state of the Lasso batch: bA, alpha, u, z
// are 84 MB at K = 4096 and the B200 has 126 MB of L2).  Marks [base, base + bytes) as "persisting" for kernels
// launched on `stream`; misses outside keep streaming.  bytes == 0 clears the window and resets persisting lines.
// Returns the hit ratio actually configured through *ratio_out (window larger than the carve-out -> partial).
extern "C" int ipm_l2_persist(const void* base, unsigned long long bytes, double* ratio_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof(attr));
  if (bytes == 0 || !base) {
    attr.accessPolicyWindow.num_bytes = 0;
    IPM_CUDA_CHECK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
    IPM_CUDA_CHECK(cudaCtxResetPersistingL2Cache());
    if (ratio_out) *ratio_out = 0.0;
    return IPM_OK;
  }
  int dev = 0, max_persist = 0, max_window = 0;
  IPM_CUDA_CHECK(cudaGetDevice(&dev));
  IPM_CUDA_CHECK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
  IPM_CUDA_CHECK(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev));
  if (max_persist <= 0 || max_window <= 0) {
    if (ratio_out) *ratio_out = 0.0;
    return IPM_OK;  // no carve-out on this device: nothing to do
  }
  unsigned long long win = bytes < (unsigned long long)max_window ? bytes : (unsigned long long)max_window;
  unsigned long long carve = win < (unsigned long long)max_persist ? win : (unsigned long long)max_persist;
  IPM_CUDA_CHECK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
  const double ratio = (double)carve / (double)win;
  attr.accessPolicyWindow.base_ptr = const_cast<void*>(base);
  attr.accessPolicyWindow.num_bytes = win;
  attr.accessPolicyWindow.hitRatio = (float)ratio;
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  IPM_CUDA_CHECK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
  if (ratio_out) *ratio_out = ratio;
  return IPM_OK;
}
