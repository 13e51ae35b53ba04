// Shared device/host helpers for libipm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define IPM_OK 0
#define IPM_ERR_ARG (-1)       // bad argument (size / alignment / null pointer)
#define IPM_ERR_CUDA (-2)      // a CUDA runtime call failed (see ipm_last_cuda_error)
#define IPM_ERR_NO_DEVICE (-3) // no sm_100 device / driver entry point missing

extern "C" int ipm_set_cuda_error(cudaError_t e);

#define IPM_CUDA_CHECK(expr)                                   \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ipm_set_cuda_error(_e);      \
  } while (0)

extern "C" void ipm_count_launch(void);
// one call per kernel launch: checks the launch and bumps the library's launch counter (ipm_launch_count)
#define IPM_LAUNCH_CHECK()                 \
  do {                                     \
    ipm_count_launch();                    \
    IPM_CUDA_CHECK(cudaGetLastError());    \
  } while (0)

// Watchdog word of the device-side spin waits: ONE pinned, mapped host allocation per process (device-visible through
// UVA).  word[0] = fault code (0 = none; written by the first wait that gives up, read by every other wait so that a
// stuck grid drains instead of hanging), word[1] = spin limit in units of 2^20 SM cycles (host-written).  The host
// reads word[0] without a synchronisation through ipm_device_fault().  nullptr if the allocation failed (waits are then
// unbounded, as before).
extern "C" unsigned int* ipm_internal_fault_word(void);

#define IPM_FAULT_POTRF_DAG 1u    // tile-DAG Cholesky: a column-progress counter never arrived
#define IPM_FAULT_STREAMK 2u      // persistent GEMM: a stream-K partial never arrived
#define IPM_FAULT_PEER_REDUCE 3u  // row-sharded Hessian: a peer's partial tile never arrived
#define IPM_FAULT_PEER_WAIT 4u    // row-sharded Hessian: the owners' final tiles never arrived
#define IPM_FAULT_TRSV 5u         // triangular solve: a solution block was never published
#define IPM_FAULT_LASSO 6u        // persistent ADMM kernel: a neighbour panel never arrived
#define IPM_FAULT_POTRF_PEER 7u   // distributed tile-DAG Cholesky: a peer's tiles never arrived
#define IPM_FAULT_HESS_I8 8u      // INT8 Hessian kernel: an mbarrier of its TMA / MMA / epilogue pipeline never completed

namespace ipm {

constexpr int kWarp = 32;

// Bounded spin: polls `ready()` until it holds.  Every 1024 polls it looks at the watchdog word: if somebody else has
// already given up it returns false at once, and after word[1] * 2^20 cycles of its own it records `code` and returns
// false.  Callers then carry on as if the wait had succeeded -- the results are garbage, but every kernel of the process
// terminates in bounded time and the host sees ipm_device_fault() != 0 (and potrf reports info = -1).
template <class Ready>
__device__ __forceinline__ bool spin_wait(Ready ready, unsigned int* fault, unsigned int code) {
  if (ready()) return true;
  const long long t0 = clock64();
  for (unsigned int spins = 1;; ++spins) {
    if (ready()) return true;
    if ((spins & 1023u) == 0 && fault) {
      volatile unsigned int* f = fault;
      if (f[0] != 0u) return false;
      if (((clock64() - t0) >> 20) > (long long)f[1]) {
        f[0] = code;
        __threadfence_system();
        return false;
      }
    }
  }
}

// Programmatic dependent launch (sm_90+).  A kernel launched with launch_pdl() may be scheduled while the previous
// kernel of the stream is still draining; it must execute pdl_wait() before it touches anything that kernel wrote
// (a no-op when the launch did not carry the attribute).  Used on the chains of short dependent kernels (Cholesky
// panel chain, ADMM iterations), where the ~3 us launch latency between kernels is a visible share of the chain.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: opt in once per (kernel, device).  `done` is a
// zero-initialised static array owned by the call site.
constexpr int kMaxDevices = 16;
template <class Kernel>
static inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, bool (&done)[kMaxDevices]) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < kMaxDevices && done[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev >= 0 && dev < kMaxDevices) done[dev] = true;
  return e;
}

// Cooperative launch: the driver guarantees that every CTA of the grid is resident at the same time (or fails the
// launch) -- what the persistent kernels whose CTAs wait for each other rely on.
template <class... KArgs, class... Args>
static inline cudaError_t launch_cooperative(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                             cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum in a fixed (deterministic) order; result valid in thread 0. `red` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = lane < nw ? red[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}
__device__ __forceinline__ double block_min(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_min(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = INFINITY;
  if (w == 0) {
    r = lane < nw ? red[lane] : INFINITY;
    r = warp_min(r);
  }
  return r;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace ipm
