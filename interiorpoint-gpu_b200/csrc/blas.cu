// HBM-streaming level-1/2 kernels (FP64): row GEMV, transposed GEMV with deterministic two-stage column
// sums, batched dot products, device-scalar AXPY.  All reductions use a fixed summation order.
#include "common.cuh"

using namespace ipm;

// ------------------------------------------------------------------------------------------------
// y[r] = alpha * dot(M[r, :], x) + beta * y[r]          (one warp per row, 16-byte loads)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemv_n_kernel(const double* __restrict__ Mx, long long ld, int rows, int cols,
                                                     const double* __restrict__ x, double* __restrict__ y,
                                                     double alpha, double beta) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const bool vec = ((ld & 1) == 0) && ((((uintptr_t)Mx) & 15) == 0) && ((((uintptr_t)x) & 15) == 0);
  for (int r = warp_global; r < rows; r += nwarps) {
    const double* row = Mx + (long long)r * ld;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    if (vec) {
      const double2* row2 = reinterpret_cast<const double2*>(row);
      const double2* x2 = reinterpret_cast<const double2*>(x);
      const int n2 = cols >> 1;
      int j = lane;
      for (; j + 96 < n2; j += 128) {
        double2 a0 = __ldcs(row2 + j), a1 = __ldcs(row2 + j + 32), a2 = __ldcs(row2 + j + 64), a3 = __ldcs(row2 + j + 96);
        double2 b0 = x2[j], b1 = x2[j + 32], b2 = x2[j + 64], b3 = x2[j + 96];
        acc0 = fma(a0.x, b0.x, acc0); acc0 = fma(a0.y, b0.y, acc0);
        acc1 = fma(a1.x, b1.x, acc1); acc1 = fma(a1.y, b1.y, acc1);
        acc2 = fma(a2.x, b2.x, acc2); acc2 = fma(a2.y, b2.y, acc2);
        acc3 = fma(a3.x, b3.x, acc3); acc3 = fma(a3.y, b3.y, acc3);
      }
      for (; j < n2; j += 32) {
        double2 a0 = __ldcs(row2 + j);
        double2 b0 = x2[j];
        acc0 = fma(a0.x, b0.x, acc0); acc0 = fma(a0.y, b0.y, acc0);
      }
      if ((cols & 1) && lane == 0) acc1 = fma(row[cols - 1], x[cols - 1], acc1);
    } else {
      for (int j = lane; j < cols; j += 32) acc0 = fma(row[j], x[j], acc0);
    }
    double acc = warp_sum((acc0 + acc1) + (acc2 + acc3));
    if (lane == 0) y[r] = (beta == 0.0) ? alpha * acc : fma(alpha, acc, beta * y[r]);
  }
}

extern "C" int ipm_gemv_n_f64(const double* Mx, int ld, int rows, int cols, const double* x, double* y, double alpha,
                              double beta, void* stream) {
  if (rows < 0 || cols < 0 || ld < cols) return IPM_ERR_ARG;
  if (rows == 0) return IPM_OK;
  if (!Mx || !x || !y) return IPM_ERR_ARG;
  int blocks = ceil_div(rows, 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  gemv_n_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(Mx, ld, rows, cols, x, y, alpha, beta);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// ------------------------------------------------------------------------------------------------
// Y[v][j] = sum_r V[v][r] * M[r][j]     v < NV   (thread = 2 adjacent columns, rows split in chunks;
// stage 1 writes per-chunk partials, stage 2 adds them in chunk order -> deterministic)
// ------------------------------------------------------------------------------------------------
constexpr int GT_ROWS = 256;  // rows per chunk
constexpr int GT_COLS = 256;  // columns per CTA (128 threads x 2)

template <int NV>
__global__ void __launch_bounds__(128) gemv_t_partial_kernel(const double* __restrict__ Mx, long long ld, int rows,
                                                             int cols, const double* __restrict__ V, long long ldv,
                                                             double* __restrict__ part /*[chunks][NV][cols]*/) {
  __shared__ double vs[NV][GT_ROWS];
  const int r0 = blockIdx.y * GT_ROWS;
  const int nr = min(GT_ROWS, rows - r0);
  for (int i = threadIdx.x; i < GT_ROWS; i += blockDim.x)
#pragma unroll
    for (int v = 0; v < NV; ++v) vs[v][i] = i < nr ? V[v * ldv + r0 + i] : 0.0;
  __syncthreads();
  const int j = blockIdx.x * GT_COLS + 2 * threadIdx.x;
  if (j >= cols) return;
  const bool pair = (j + 1 < cols);
  const bool vec = pair && ((ld & 1) == 0) && ((((uintptr_t)Mx) & 15) == 0);
  double a0[NV], a1[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) a0[v] = a1[v] = 0.0;
  const double* p = Mx + (long long)r0 * ld + j;
  if (vec) {
    int i = 0;
    for (; i + 3 < nr; i += 4) {
      double2 m0 = __ldcs(reinterpret_cast<const double2*>(p + (long long)i * ld));
      double2 m1 = __ldcs(reinterpret_cast<const double2*>(p + (long long)(i + 1) * ld));
      double2 m2 = __ldcs(reinterpret_cast<const double2*>(p + (long long)(i + 2) * ld));
      double2 m3 = __ldcs(reinterpret_cast<const double2*>(p + (long long)(i + 3) * ld));
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        a0[v] = fma(vs[v][i], m0.x, a0[v]); a1[v] = fma(vs[v][i], m0.y, a1[v]);
        a0[v] = fma(vs[v][i + 1], m1.x, a0[v]); a1[v] = fma(vs[v][i + 1], m1.y, a1[v]);
        a0[v] = fma(vs[v][i + 2], m2.x, a0[v]); a1[v] = fma(vs[v][i + 2], m2.y, a1[v]);
        a0[v] = fma(vs[v][i + 3], m3.x, a0[v]); a1[v] = fma(vs[v][i + 3], m3.y, a1[v]);
      }
    }
    for (; i < nr; ++i) {
      double2 m0 = __ldcs(reinterpret_cast<const double2*>(p + (long long)i * ld));
#pragma unroll
      for (int v = 0; v < NV; ++v) { a0[v] = fma(vs[v][i], m0.x, a0[v]); a1[v] = fma(vs[v][i], m0.y, a1[v]); }
    }
  } else {
    for (int i = 0; i < nr; ++i) {
      const double m0 = p[(long long)i * ld];
      const double m1 = pair ? p[(long long)i * ld + 1] : 0.0;
#pragma unroll
      for (int v = 0; v < NV; ++v) { a0[v] = fma(vs[v][i], m0, a0[v]); a1[v] = fma(vs[v][i], m1, a1[v]); }
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double* o = part + ((long long)blockIdx.y * NV + v) * cols + j;
    o[0] = a0[v];
    if (pair) o[1] = a1[v];
  }
}

__global__ void gemv_t_reduce_kernel(const double* __restrict__ part, int chunks, int nv, int cols,
                                     double* __restrict__ Y, long long ldy, double alpha, double beta) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  if (j >= cols) return;
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += part[((long long)c * nv + v) * cols + j];
  double* o = Y + v * ldy + j;
  *o = (beta == 0.0) ? alpha * s : fma(alpha, s, beta * *o);
}

extern "C" long long ipm_gemv_t_ws_doubles(int rows, int cols, int nv) {
  return (long long)ceil_div(rows > 0 ? rows : 1, GT_ROWS) * nv * cols;
}

extern "C" int ipm_gemv_t_f64(const double* Mx, int ld, int rows, int cols, const double* V, int nv, int ldv,
                              double* Y, int ldy, double alpha, double beta, double* ws, long long ws_doubles,
                              void* stream) {
  if (rows < 0 || cols <= 0 || ld < cols || nv < 1 || nv > 2 || !Y) return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = rows > 0 ? ceil_div(rows, GT_ROWS) : 0;
  if (chunks > 0) {
    if (!Mx || !V || !ws || ws_doubles < ipm_gemv_t_ws_doubles(rows, cols, nv)) return IPM_ERR_ARG;
    dim3 grid(ceil_div(cols, GT_COLS), chunks);
    if (nv == 1)
      gemv_t_partial_kernel<1><<<grid, 128, 0, st>>>(Mx, ld, rows, cols, V, ldv, ws);
    else
      gemv_t_partial_kernel<2><<<grid, 128, 0, st>>>(Mx, ld, rows, cols, V, ldv, ws);
    IPM_LAUNCH_CHECK();
  }
  dim3 g2(ceil_div(cols, 256), nv);
  gemv_t_reduce_kernel<<<g2, 256, 0, st>>>(ws, chunks, nv, cols, Y, ldy, alpha, beta);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// ------------------------------------------------------------------------------------------------
// out[k] = dot(a_k, b_k), k < npairs <= 8   (one 1024-thread CTA per pair, fixed order)
// ------------------------------------------------------------------------------------------------
struct DotArgs {
  const double* a[8];
  const double* b[8];
  int n[8];
};

__global__ void __launch_bounds__(1024) dots_kernel(DotArgs args, double* __restrict__ out) {
  __shared__ double red[32];
  const int k = blockIdx.x;
  const double* a = args.a[k];
  const double* b = args.b[k];
  const int n = args.n[k];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc = fma(a[i], b[i], acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[k] = acc;
}

extern "C" int ipm_dots_f64(int npairs, const double* const* a, const double* const* b, const int* n, double* out,
                            void* stream) {
  if (npairs < 1 || npairs > 8 || !a || !b || !n || !out) return IPM_ERR_ARG;
  DotArgs args;
  for (int k = 0; k < npairs; ++k) {
    args.a[k] = a[k];
    args.b[k] = b[k];
    args.n[k] = n[k];
  }
  dots_kernel<<<npairs, 1024, 0, (cudaStream_t)stream>>>(args, out);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// ------------------------------------------------------------------------------------------------
// y += (*a_dev) * x   with the scalar on the device (no host round trip inside the Newton step)
// ------------------------------------------------------------------------------------------------
__global__ void axpy_dev_kernel(int n, const double* __restrict__ a_dev, const double* __restrict__ x,
                                double* __restrict__ y) {
  const double a = *a_dev;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    y[i] = __dadd_rn(y[i], __dmul_rn(a, x[i]));  // NumPy rounding: multiply, then add
}

extern "C" int ipm_axpy_dev_f64(int n, const double* a_dev, const double* x, double* y, void* stream) {
  if (n < 0 || !a_dev || (n > 0 && (!x || !y))) return IPM_ERR_ARG;
  if (n == 0) return IPM_OK;
  int blocks = ceil_div(n, 256);
  if (blocks > 1184) blocks = 1184;
  axpy_dev_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, a_dev, x, y);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}
