// Sparse-aware pieces for LPs whose inequality matrix is stored dense but is almost entirely zero (the MIPLIB
// `.npy` problems of testSolver.py:278-300: aflow40b has ~0.17 % non-zeros, so a dense SYRK wastes 99.8 % of its
// flops -- SURVEY.md 8(f)-1, and the reference's poster: "Implementing sparse matrix handling would greatly improve
// performance").  The Hessian itself stays dense (it is factorised by the dense Cholesky); what becomes sparse is
// everything that touches C:
//   * ipm_csr_gemv_f64        y = alpha * S x + beta * y   for S in CSR (used with CSR(C) for slacks and C dx, and
//                             with CSR(C^T) for the gradient, so no atomics and a fixed summation order);
//   * ipm_sparse_syrk_f64     H[i][j] += sum_{r in seg(i,j)} w[r] * p_r   for the precomputed list of output
//                             entries (i <= j) of C^T diag(w) C: one thread per output entry walks its segment of
//                             (row, product c_ri * c_rj) pairs in ascending row order -- deterministic, no atomics.
#include "common.cuh"

using namespace ipm;

// one thread per row (rows of these matrices hold a handful of entries)
__global__ void __launch_bounds__(256) csr_gemv_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                                       const double* __restrict__ val, int rows,
                                                       const double* __restrict__ x, double* __restrict__ y,
                                                       double alpha, double beta) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double acc = 0.0;
  for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) acc = fma(val[k], x[col[k]], acc);
  y[r] = (beta == 0.0) ? alpha * acc : fma(alpha, acc, beta * y[r]);
}

extern "C" int ipm_csr_gemv_f64(const int* rowptr, const int* col, const double* val, int rows, const double* x,
                                double* y, double alpha, double beta, void* stream) {
  if (rows < 0) return IPM_ERR_ARG;
  if (rows == 0) return IPM_OK;
  if (!rowptr || !x || !y) return IPM_ERR_ARG;
  csr_gemv_kernel<<<ceil_div(rows, 256), 256, 0, (cudaStream_t)stream>>>(rowptr, col, val, rows, x, y, alpha, beta);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

__global__ void __launch_bounds__(256) sparse_syrk_kernel(int nout, const int* __restrict__ segptr,
                                                          const int* __restrict__ seg_row,
                                                          const double* __restrict__ seg_prod,
                                                          const int* __restrict__ out_i,
                                                          const int* __restrict__ out_j, long long ld,
                                                          const double* __restrict__ w, double* __restrict__ H) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nout) return;
  double acc = 0.0;
  for (int k = segptr[e]; k < segptr[e + 1]; ++k) acc = fma(w[seg_row[k]], seg_prod[k], acc);
  H[(long long)out_i[e] * ld + out_j[e]] += acc;
}

// H (upper triangle, row-major, leading dimension ld) += C^T diag(w) C for the `nout` structurally non-zero
// entries (out_i[e], out_j[e]), out_i <= out_j.
extern "C" int ipm_sparse_syrk_f64(int nout, const int* segptr, const int* seg_row, const double* seg_prod,
                                   const int* out_i, const int* out_j, const double* w, double* H, int ld,
                                   void* stream) {
  if (nout < 0) return IPM_ERR_ARG;
  if (nout == 0) return IPM_OK;
  if (!segptr || !seg_row || !seg_prod || !out_i || !out_j || !w || !H) return IPM_ERR_ARG;
  sparse_syrk_kernel<<<ceil_div(nout, 256), 256, 0, (cudaStream_t)stream>>>(nout, segptr, seg_row, seg_prod, out_i,
                                                                          out_j, ld, w, H);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}
