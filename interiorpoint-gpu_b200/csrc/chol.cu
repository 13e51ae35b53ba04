// Dense SPD factorisation and triangular solves (FP64), upper / row-major:   H = U^T U.
//
// Storing the UPPER factor row-major makes every bulk update of the factorisation the same "TN"
// contraction as the Hessian itself (k = row index of the panel, see gemm_tn_core.cuh):
//     trailing update   A22 -= U12^T U12        -> ipm_gemm_tn_f64(A = B = U12, alpha = -1, beta = 1, upper)
//     TRSM update       B2  -= U12^T Y1         -> ipm_gemm_tn_f64(A = U12, B = Y1)
// so the DMMA/TMA core does n^3/3 of the n^3/3 flops; the per-panel pieces below (128x128 diagonal
// factor, 128-row triangular panel solve) are the serial O(n^2 NB) remainder.
#include "common.cuh"

using namespace ipm;

extern "C" int ipm_gemm_tn_f64(const double* A, int lda, const double* B, int ldb, const double* w, double alpha,
                               double beta, double* D, int ldd, int M, int N, int K, int upper, void* stream);

constexpr int NB = 128;  // panel height == GEMM tile

// ------------------------------------------------------------------------------------------------
// Diagonal block: unblocked right-looking Cholesky of an nb x nb (nb <= 128) upper block held in smem.
// info (1-based global index of the first non-positive pivot) is written once; 0 means success.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) potf2_kernel(double* __restrict__ A, long long ld, int nb, int k0,
                                                       int* __restrict__ info) {
  extern __shared__ double S[];  // nb x LDS_
  const int LDS_ = NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int r = warp; r < nb; r += nwarps)
    for (int c = r + lane; c < nb; c += 32) S[r * LDS_ + c] = A[(long long)r * ld + c];
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    if (tid == 0) {
      const double d = S[j * LDS_ + j];
      if (!(d > 0.0) && atomicCAS(info, 0, k0 + j + 1) == 0) { /* first failure recorded */ }
      S[j * LDS_ + j] = sqrt(d);
    }
    __syncthreads();
    const double dinv = 1.0 / S[j * LDS_ + j];
    for (int c = j + 1 + tid; c < nb; c += blockDim.x) S[j * LDS_ + c] *= dinv;
    __syncthreads();
    for (int r = j + 1 + warp; r < nb; r += nwarps) {
      const double ujr = S[j * LDS_ + r];
      for (int c = r + lane; c < nb; c += 32) S[r * LDS_ + c] = fma(-ujr, S[j * LDS_ + c], S[r * LDS_ + c]);
    }
    // the next pivot S[j+1][j+1] is final after this update; the barrier at the top of the next step orders it
    __syncthreads();
  }
  for (int r = warp; r < nb; r += nwarps)
    for (int c = r + lane; c < nb; c += 32) A[(long long)r * ld + c] = S[r * LDS_ + c];
}

// ------------------------------------------------------------------------------------------------
// Panel solve  X = U11^{-T} P   (U11: nb x nb upper, P: nb x ncols row-major, in place).
// CTA = 64 columns; U11 and the CTA's panel slice live in shared memory (192 KiB); substitution runs in
// 32-row blocks (x kept in registers, true divisions -- no explicit inverse, same backward stability as
// LAPACK dtrsm), the rows below each block are updated by all 256 threads.
// ------------------------------------------------------------------------------------------------
constexpr int TP_COLS = 64;

__global__ void __launch_bounds__(256, 1) trsm_panel_kernel(const double* __restrict__ U11, long long ldu, int nb,
                                                            double* __restrict__ P, long long ldp, int ncols) {
  extern __shared__ double sm[];
  double* Us = sm;             // NB x NB
  double* Ps = sm + NB * NB;   // NB x TP_COLS
  const int tid = threadIdx.x;
  const int c = tid & (TP_COLS - 1), tr = tid >> 6;  // 4 row phases
  const int col0 = blockIdx.x * TP_COLS;
  const int ncl = min(TP_COLS, ncols - col0);
  for (int idx = tid; idx < nb * NB; idx += blockDim.x) {
    const int r = idx / NB, cc = idx - r * NB;
    Us[idx] = (cc >= r && cc < nb) ? U11[(long long)r * ldu + cc] : 0.0;
  }
  for (int idx = tid; idx < nb * TP_COLS; idx += blockDim.x) {
    const int r = idx / TP_COLS, cc = idx - r * TP_COLS;
    Ps[idx] = cc < ncl ? P[(long long)r * ldp + col0 + cc] : 0.0;
  }
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += 32) {
    const int bl = min(32, nb - b0);
    if (tr == 0) {
      double x[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        if (r < bl) {
          double v = Ps[(b0 + r) * TP_COLS + c];
#pragma unroll
          for (int l = 0; l < r; ++l) v = fma(-Us[(b0 + l) * NB + b0 + r], x[l], v);
          x[r] = v / Us[(b0 + r) * NB + b0 + r];
          Ps[(b0 + r) * TP_COLS + c] = x[r];
        }
      }
    }
    __syncthreads();
    const int rest0 = b0 + bl;
    for (int rb = rest0 + 4 * tr; rb < nb; rb += 16) {
      double acc[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = (rb + q < nb) ? Ps[(rb + q) * TP_COLS + c] : 0.0;
      for (int l = 0; l < bl; ++l) {
        const double xl = Ps[(b0 + l) * TP_COLS + c];
        const double* urow = Us + (b0 + l) * NB + rb;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = fma(-(rb + q < nb ? urow[q] : 0.0), xl, acc[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (rb + q < nb) Ps[(rb + q) * TP_COLS + c] = acc[q];
    }
    __syncthreads();
  }
  for (int idx = tid; idx < nb * TP_COLS; idx += blockDim.x) {
    const int r = idx / TP_COLS, cc = idx - r * TP_COLS;
    if (cc < ncl) P[(long long)r * ldp + col0 + cc] = Ps[idx];
  }
}

static int launch_trsm_panel(const double* U11, long long ldu, int nb, double* P, long long ldp, int ncols,
                             cudaStream_t st) {
  if (ncols <= 0) return IPM_OK;
  const int smem = (NB * NB + NB * TP_COLS) * 8;
  IPM_CUDA_CHECK(cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  trsm_panel_kernel<<<ceil_div(ncols, TP_COLS), 256, smem, st>>>(U11, ldu, nb, P, ldp, ncols);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// ------------------------------------------------------------------------------------------------
// ipm_potrf_upper_f64: in-place blocked right-looking Cholesky, H = U^T U (upper triangle of H in/out; the
// strict lower triangle is never read or written).  *info_dev = 0 on success, else the 1-based index of the
// first non-positive pivot (LAPACK dpotrf convention); it is written on the device, never synchronised here.
// ------------------------------------------------------------------------------------------------
extern "C" int ipm_potrf_upper_f64(double* H, int ld, int n, int* info_dev, void* stream) {
  if (!H || !info_dev || n < 0 || ld < n || (ld & 1)) return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  IPM_CUDA_CHECK(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
  IPM_CUDA_CHECK(cudaFuncSetAttribute(potf2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NB * NB * 8));
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int nb = n - k0 < NB ? n - k0 : NB;
    double* Akk = H + (long long)k0 * ld + k0;
    potf2_kernel<<<1, 512, NB * NB * 8, st>>>(Akk, ld, nb, k0, info_dev);
    IPM_LAUNCH_CHECK();
    const int rest = n - k0 - nb;
    if (rest > 0) {
      double* A12 = Akk + nb;
      int rc = launch_trsm_panel(Akk, ld, nb, A12, ld, rest, st);
      if (rc) return rc;
      double* A22 = H + (long long)(k0 + nb) * ld + (k0 + nb);
      rc = ipm_gemm_tn_f64(A12, ld, A12, ld, nullptr, -1.0, 1.0, A22, ld, rest, rest, nb, 1, stream);
      if (rc) return rc;
    }
  }
  return IPM_OK;
}

// ------------------------------------------------------------------------------------------------
// ipm_trsm_upper_t_f64:  B <- U^{-T} B   (U: n x n upper row-major, B: n x p row-major), blocked forward
// substitution: 128-row panel solve + DMMA update of the rows below.
// ------------------------------------------------------------------------------------------------
extern "C" int ipm_trsm_upper_t_f64(const double* U, int ldu, int n, double* B, int ldb, int p, void* stream) {
  if (!U || !B || n < 0 || p < 0 || ldu < n || ldb < p || (ldu & 1) || (ldb & 1)) return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  for (int k0 = 0; k0 < n && p > 0; k0 += NB) {
    const int nb = n - k0 < NB ? n - k0 : NB;
    const double* Ukk = U + (long long)k0 * ldu + k0;
    double* Bk = B + (long long)k0 * ldb;
    int rc = launch_trsm_panel(Ukk, ldu, nb, Bk, ldb, p, st);
    if (rc) return rc;
    const int rest = n - k0 - nb;
    if (rest > 0) {
      rc = ipm_gemm_tn_f64(Ukk + nb, ldu, Bk, ldb, nullptr, -1.0, 1.0, B + (long long)(k0 + nb) * ldb, ldb, rest, p,
                           nb, 0, stream);
      if (rc) return rc;
    }
  }
  return IPM_OK;
}

// ------------------------------------------------------------------------------------------------
// Vector triangular solves (HBM-bound: U is read once).  Right-looking in 128-wide strips:
//   trans = 1:  solve U^T y = b  (top -> bottom)      trans = 0:  solve U x = b  (bottom -> top)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) trsv_diag_kernel(const double* __restrict__ Ukk, long long ld, int nb,
                                                           double* __restrict__ b, int trans) {
  extern __shared__ double S[];  // nb x (NB + 1)
  __shared__ double xs[NB];
  const int LDS_ = NB + 1;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < nb * NB; idx += blockDim.x) {
    const int r = idx / NB, c = idx - r * NB;
    if (c < nb) S[r * LDS_ + c] = c >= r ? Ukk[(long long)r * ld + c] : 0.0;
  }
  if (tid < nb) xs[tid] = b[tid];
  __syncthreads();
  if (trans) {
    // forward: y_r = (b_r - sum_{l<r} U[l][r] y_l) / U[r][r]; right-looking update of the later entries
    for (int r = 0; r < nb; ++r) {
      if (tid == r) xs[r] = xs[r] / S[r * LDS_ + r];
      __syncthreads();
      if (tid > r && tid < nb) xs[tid] = fma(-S[r * LDS_ + tid], xs[r], xs[tid]);
      __syncthreads();
    }
  } else {
    for (int r = nb - 1; r >= 0; --r) {
      if (tid == r) xs[r] = xs[r] / S[r * LDS_ + r];
      __syncthreads();
      if (tid < r) xs[tid] = fma(-S[tid * LDS_ + r], xs[r], xs[tid]);
      __syncthreads();
    }
  }
  if (tid < nb) b[tid] = xs[tid];
}

// b[j] -= sum_{i<nb} U[i][j] * y[i]   for j in [0, ncols): rows i are the strip just solved.
__global__ void __launch_bounds__(128) trsv_update_fwd_kernel(const double* __restrict__ Us, long long ld, int nb,
                                                              const double* __restrict__ y, double* __restrict__ b,
                                                              int ncols) {
  __shared__ double ys[NB];
  for (int i = threadIdx.x; i < NB; i += blockDim.x) ys[i] = i < nb ? y[i] : 0.0;
  __syncthreads();
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (j >= ncols) return;
  const bool pair = j + 1 < ncols;
  const bool vec = pair && !(ld & 1) && !(((uintptr_t)(Us + j)) & 15);
  double a0 = 0.0, a1 = 0.0;
  if (vec) {
    int i = 0;
    for (; i + 7 < nb; i += 8) {
      double2 m[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) m[q] = *reinterpret_cast<const double2*>(Us + (long long)(i + q) * ld + j);
#pragma unroll
      for (int q = 0; q < 8; ++q) { a0 = fma(m[q].x, ys[i + q], a0); a1 = fma(m[q].y, ys[i + q], a1); }
    }
    for (; i < nb; ++i) {
      double2 m = *reinterpret_cast<const double2*>(Us + (long long)i * ld + j);
      a0 = fma(m.x, ys[i], a0); a1 = fma(m.y, ys[i], a1);
    }
  } else {
    for (int i = 0; i < nb; ++i) {
      a0 = fma(Us[(long long)i * ld + j], ys[i], a0);
      if (pair) a1 = fma(Us[(long long)i * ld + j + 1], ys[i], a1);
    }
  }
  b[j] -= a0;
  if (pair) b[j + 1] -= a1;
}

// b[i] -= sum_{j<nb} U[i][j] * x[j]   for rows i in [0, nrows): one warp per row, strip columns contiguous.
__global__ void __launch_bounds__(256) trsv_update_bwd_kernel(const double* __restrict__ Us, long long ld, int nb,
                                                              const double* __restrict__ x, double* __restrict__ b,
                                                              int nrows) {
  __shared__ double xs[NB];
  for (int i = threadIdx.x; i < NB; i += blockDim.x) xs[i] = i < nb ? x[i] : 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp_global; r < nrows; r += nwarps) {
    const double* row = Us + (long long)r * ld;
    double acc = 0.0;
    for (int j = lane; j < nb; j += 32) acc = fma(row[j], xs[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) b[r] -= acc;
  }
}

extern "C" int ipm_trsv_upper_f64(const double* U, int ld, int n, double* b, int trans, void* stream) {
  if (!U || !b || n < 0 || ld < n) return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int smem = NB * (NB + 1) * 8;
  IPM_CUDA_CHECK(cudaFuncSetAttribute(trsv_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (trans) {
    for (int k0 = 0; k0 < n; k0 += NB) {
      const int nb = n - k0 < NB ? n - k0 : NB;
      const double* Ukk = U + (long long)k0 * ld + k0;
      trsv_diag_kernel<<<1, 128, smem, st>>>(Ukk, ld, nb, b + k0, 1);
      IPM_LAUNCH_CHECK();
      const int rest = n - k0 - nb;
      if (rest > 0) {
        trsv_update_fwd_kernel<<<ceil_div(rest, 256), 128, 0, st>>>(Ukk + nb, ld, nb, b + k0, b + k0 + nb, rest);
        IPM_LAUNCH_CHECK();
      }
    }
  } else {
    const int nblk = ceil_div(n, NB);
    for (int kb = nblk - 1; kb >= 0; --kb) {
      const int k0 = kb * NB;
      const int nb = n - k0 < NB ? n - k0 : NB;
      const double* Ukk = U + (long long)k0 * ld + k0;
      trsv_diag_kernel<<<1, 128, smem, st>>>(Ukk, ld, nb, b + k0, 0);
      IPM_LAUNCH_CHECK();
      if (k0 > 0) {
        int blocks = ceil_div(k0, 8);
        if (blocks > 148 * 4) blocks = 148 * 4;
        trsv_update_bwd_kernel<<<blocks, 256, 0, st>>>(U + k0, ld, nb, b + k0, b, k0);
        IPM_LAUNCH_CHECK();
      }
    }
  }
  return IPM_OK;
}
