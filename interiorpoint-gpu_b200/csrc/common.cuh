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

namespace ipm {

constexpr int kWarp = 32;

// Programmatic dependent launch (sm_90+).  A kernel launched with launch_pdl() may be scheduled while the previous
// kernel of the stream is still draining; it must execute pdl_wait() before it touches anything that kernel wrote
// (a no-op when the launch did not carry the attribute).  Used on the chains of short dependent kernels (Cholesky
// panel chain, ADMM iterations), where the ~3 us launch latency between kernels is a visible share of the chain.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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
