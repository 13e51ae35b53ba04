// Library status / error plumbing.
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
