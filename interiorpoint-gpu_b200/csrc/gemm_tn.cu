// ipm_gemm_tn_f64: D = beta*D + alpha * A^T diag(w) B   (FP64, TMA + DMMA, see gemm_tn_core.cuh)
#include "common.cuh"
#include "gemm_tn_core.cuh"
#include "tensormap.cuh"

namespace ipm {

struct PlainEpilogue {
  double* D;
  long long ldd;
  int M, N;
  double alpha, beta;
  int upper;
  __device__ __forceinline__ void put(int row, int col, double v) const {
    if (row >= M || col >= N) return;
    if (upper && col < row) return;
    double* p = D + (long long)row * ldd + col;
    *p = (beta == 0.0) ? alpha * v : fma(alpha, v, beta * *p);
  }
  __device__ __forceinline__ void operator()(int row, int col, double v0, double v1) const {
    put(row, col, v0);
    put(row, col + 1, v1);
  }
};

}  // namespace ipm

using namespace ipm;

extern "C" int ipm_gemm_tn_f64(const double* A, int lda, const double* B, int ldb, const double* w, double alpha,
                               double beta, double* D, int ldd, int M, int N, int K, int upper, void* stream) {
  if (!A || !B || !D || M <= 0 || N <= 0 || K < 0 || lda < M || ldb < N || ldd < N) return IPM_ERR_ARG;
  if (upper && M != N) return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmA, tmB;
  int rc = make_operand_map(&tmA, A, lda, K, M);
  if (rc) return rc;
  rc = make_operand_map(&tmB, B, ldb, K, N);
  if (rc) return rc;
  const int tm = ceil_div(M, gemm::BM), tn = ceil_div(N, gemm::BN);
  const int tiles = upper ? tn * (tn + 1) / 2 : tm * tn;
  PlainEpilogue epi{D, ldd, M, N, alpha, beta, upper};
  if (w) {
    auto kern = gemm::gemm_tn_kernel<true, PlainEpilogue>;
    IPM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES));
    kern<<<tiles, gemm::THREADS, gemm::SMEM_BYTES, st>>>(tmA, tmB, M, N, K, w, upper, epi);
  } else {
    auto kern = gemm::gemm_tn_kernel<false, PlainEpilogue>;
    IPM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES));
    kern<<<tiles, gemm::THREADS, gemm::SMEM_BYTES, st>>>(tmA, tmB, M, N, K, nullptr, upper, epi);
  }
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}
