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

// Watchdog word (see common.cuh): pinned + mapped + portable, so one allocation serves every device of the process.
#include <stdlib.h>

#include <mutex>
static unsigned int* g_fault = nullptr;
static std::once_flag g_fault_once;
extern "C" unsigned int* ipm_internal_fault_word(void) {
  std::call_once(g_fault_once, [] {
    unsigned int* p = nullptr;
    if (cudaHostAlloc((void**)&p, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    // default limit: 4096 * 2^20 cycles ~ 2.2 s at 1.96 GHz (the longest legitimate wait on this path is < 60 ms);
    // IPM_SPIN_LIMIT_MCYCLES overrides it (tests use a small value to exercise the abort path)
    unsigned int limit = 4096;
    if (const char* e = getenv("IPM_SPIN_LIMIT_MCYCLES")) {
      const long v = atol(e);
      if (v > 0) limit = (unsigned int)v;
    }
    p[0] = 0u;
    p[1] = limit;
    g_fault = p;
  });
  return g_fault;
}
// Fault code recorded by a device-side wait that gave up (0 = none).  Valid after the stream has been synchronised.
extern "C" unsigned int ipm_device_fault(void) { return g_fault ? *(volatile unsigned int*)g_fault : 0u; }
extern "C" void ipm_clear_device_fault(void) {
  if (g_fault) *(volatile unsigned int*)g_fault = 0u;
}
// Test hook: spin limit in units of 2^20 cycles (returns the previous value).
extern "C" unsigned int ipm_set_spin_limit(unsigned int mcycles) {
  unsigned int* p = ipm_internal_fault_word();
  if (!p) return 0u;
  const unsigned int old = p[1];
  if (mcycles > 0) p[1] = mcycles;
  return old;
}

// Self-test of the watchdog (tests/test_kernels_gpu.py): one thread waits for a flag nobody sets.  With a small spin
// limit the kernel must return and leave `code` in the fault word.
__global__ void watchdog_selftest_kernel(const unsigned int* never_set, unsigned int* fault, unsigned int code,
                                         int* gave_up) {
  const bool ok = ipm::spin_wait([&] { return *(volatile const unsigned int*)never_set == 0xdeadbeefu; }, fault, code);
  *gave_up = ok ? 0 : 1;
}
extern "C" int ipm_internal_watchdog_selftest(unsigned int* flag_dev, int* gave_up_dev, unsigned int code, void* stream) {
  if (!flag_dev || !gave_up_dev || !ipm_internal_fault_word()) return IPM_ERR_ARG;
  watchdog_selftest_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag_dev, ipm_internal_fault_word(), code, gave_up_dev);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
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
