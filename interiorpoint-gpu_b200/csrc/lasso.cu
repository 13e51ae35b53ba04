// Batched ADMM Lasso (LassoSolver.py:240-337): one kernel per iteration.
//
//   x      = bA + Q~ (u - alpha)                 Q~ = -m*rho*(A'A + m*rho*I)^{-1}   (n x n, symmetric)
//   alpha+ = prox(x + u, eta_c)                  soft threshold per column c, bias row exempt (:517-543)
//   u+     = u + x - alpha+
//
// The state is n x K (K problems as columns).  The GEMM  Q~ z  (z = u - alpha, kept as a third state array so
// the operand is a plain TMA tile) runs on the DMMA core of gemm_tn_core.cuh; everything else is its epilogue:
// each output tile reads bA / u / alpha once, writes alpha+, u+ and z+ = u+ - alpha+ (into the OTHER z buffer:
// other CTAs are still reading z), and, on stop-check iterations, leaves four partial sums of squares per warp
// (|x - alpha+|^2, |rho (alpha+ - alpha)|^2, |alpha+|^2, |u+|^2) for the batch-coupled stop test (:273-298).
#include "common.cuh"
#include "gemm_tn_core.cuh"
#include "tensormap.cuh"

namespace ipm {

struct LassoEpilogue {
  const double* bA;
  const double* eta;   // per column
  double* alpha;
  double* u;
  double* z_out;
  long long ld;        // common leading dimension of bA / alpha / u / z
  int n, K;
  double rho;
  int add_bias, positive, want_norms;
  double* partials;    // [gridDim.x][8 warps][4]

  __device__ __forceinline__ void tile(const double (&acc)[8][4][2], int m_base, int n_base, int g8, int l4) const {
    double s_r = 0.0, s_d = 0.0, s_a = 0.0, s_u = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = m_base + i * 8 + g8;
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = n_base + jn * 8 + 2 * l4 + e;
          if (row < n && col < K) {
            const long long idx = (long long)row * ld + col;
            const double x = bA[idx] + acc[i][jn][e];
            const double uo = u[idx], ao = alpha[idx];
            const double v = x + uo;
            const double et = eta[col];
            double an = fmax(v - et, 0.0);
            if (!positive) an -= fmax(-v - et, 0.0);
            if (add_bias && row == 0) an = v;
            const double un = uo + x - an;
            alpha[idx] = an;
            u[idx] = un;
            z_out[idx] = un - an;
            if (want_norms) {
              const double r = x - an, dd = rho * (an - ao);
              s_r = fma(r, r, s_r);
              s_d = fma(dd, dd, s_d);
              s_a = fma(an, an, s_a);
              s_u = fma(un, un, s_u);
            }
          }
        }
      }
    }
    if (want_norms) {
      s_r = warp_sum(s_r);
      s_d = warp_sum(s_d);
      s_a = warp_sum(s_a);
      s_u = warp_sum(s_u);
      if ((threadIdx.x & 31) == 0) {
        double* p = partials + ((long long)blockIdx.x * gemm::CONSUMER_WARPS + (threadIdx.x >> 5)) * 4;
        p[0] = s_r; p[1] = s_d; p[2] = s_a; p[3] = s_u;
      }
    }
  }
};

// fixed-order sum of the per-warp partials -> out[4] = squared norms
__global__ void __launch_bounds__(256) lasso_norms_kernel(const double* __restrict__ partials, int count,
                                                          double* __restrict__ out) {
  __shared__ double red[32];
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < count; i += blockDim.x)
#pragma unroll
    for (int q = 0; q < 4; ++q) a[q] += partials[(long long)i * 4 + q];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double s = block_sum(a[q], red);
    if (threadIdx.x == 0) out[q] = s;
  }
}

// f_c = 1/(2m) sum_r R[r][c]^2 + reg_c * sum_{j >= add_bias} |alpha[j][c]|     (LassoSolver.py:314-325)
__global__ void __launch_bounds__(256)
lasso_objective_kernel(const double* __restrict__ R, long long ldr, int m, const double* __restrict__ alpha,
                       long long lda, int n, int K, const double* __restrict__ reg, int add_bias, int positive,
                       double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= K) return;
  double sq = 0.0, l1 = 0.0;
  for (int r = 0; r < m; ++r) {
    const double v = R[(long long)r * ldr + c];
    sq = fma(v, v, sq);
  }
  for (int j = add_bias ? 1 : 0; j < n; ++j) {
    const double a = alpha[(long long)j * lda + c];
    l1 += positive ? a : fabs(a);
  }
  out[c] = sq / (2.0 * m) + reg[c] * l1;
}

}  // namespace ipm

using namespace ipm;

extern "C" long long ipm_lasso_partials_doubles(int n, int K) {
  return (long long)ceil_div(n, gemm::BM) * ceil_div(K, gemm::BN) * gemm::CONSUMER_WARPS * 4;
}

// One ADMM iteration for all K problems.  Qt: n x n (ldq), z_in/z_out/bA/alpha/u: n x K (ld).  z_out != z_in.
// If want_norms, norms_out[4] receives the squared Frobenius norms {|x-alpha+|, |rho(alpha+-alpha)|, |alpha+|, |u+|}.
extern "C" int ipm_lasso_admm_step_f64(const double* Qt, int ldq, int n, int K, const double* bA, const double* eta,
                                       double rho, double* alpha, double* u, const double* z_in, double* z_out,
                                       int ld, int add_bias, int positive, int want_norms, double* partials,
                                       double* norms_out, void* stream) {
  if (!Qt || !bA || !eta || !alpha || !u || !z_in || !z_out || z_in == z_out || n <= 0 || K <= 0 || ldq < n ||
      ld < K || (want_norms && (!partials || !norms_out)))
    return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmA, tmB;
  int rc = make_operand_map(&tmA, Qt, ldq, n, n);  // Q~ symmetric: Q~[k][i] is the A operand (k rows, i columns)
  if (rc) return rc;
  rc = make_operand_map(&tmB, z_in, ld, n, K);
  if (rc) return rc;
  const int tiles = ceil_div(n, gemm::BM) * ceil_div(K, gemm::BN);
  LassoEpilogue epi{bA, eta, alpha, u, z_out, ld, n, K, rho, add_bias, positive, want_norms, partials};
  auto kern = gemm::gemm_tn_kernel<false, LassoEpilogue>;
  static bool attr_set = false;
  if (!attr_set) {
    IPM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES));
    attr_set = true;
  }
  if (want_norms)
    IPM_CUDA_CHECK(cudaMemsetAsync(partials, 0, sizeof(double) * tiles * gemm::CONSUMER_WARPS * 4, st));
  kern<<<tiles, gemm::THREADS, gemm::SMEM_BYTES, st>>>(tmA, tmB, n, K, n, nullptr, 0, epi);
  IPM_LAUNCH_CHECK();
  if (want_norms) {
    lasso_norms_kernel<<<1, 256, 0, st>>>(partials, tiles * gemm::CONSUMER_WARPS, norms_out);
    IPM_LAUNCH_CHECK();
  }
  return IPM_OK;
}

extern "C" int ipm_lasso_objective_f64(const double* R, int ldr, int m, const double* alpha, int lda, int n, int K,
                                       const double* reg, int add_bias, int positive, double* out, void* stream) {
  if (!R || !alpha || !reg || !out || m <= 0 || n <= 0 || K <= 0) return IPM_ERR_ARG;
  lasso_objective_kernel<<<ceil_div(K, 256), 256, 0, (cudaStream_t)stream>>>(R, ldr, m, alpha, lda, n, K, reg,
                                                                            add_bias, positive, out);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// out[i][j] = s * in[i][j]  (rows x cols, row-major) and optional diagonal shift: out[i][i] += dshift
__global__ void scale_shift_kernel(const double* __restrict__ in, long long ldi, double* __restrict__ out,
                                   long long ldo, int rows, int cols, double s, double dshift) {
  const int i = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < cols; j += gridDim.x * blockDim.x) {
    double v = s * in[(long long)i * ldi + j];
    if (i == j) v += dshift;
    out[(long long)i * ldo + j] = v;
  }
}

extern "C" int ipm_scale_shift_f64(const double* in, int ldi, double* out, int ldo, int rows, int cols, double s,
                                   double dshift, void* stream) {
  if (!in || !out || rows <= 0 || cols <= 0 || ldi < cols || ldo < cols) return IPM_ERR_ARG;
  dim3 grid(ceil_div(cols, 256) < 8 ? ceil_div(cols, 256) : 8, rows);
  scale_shift_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, ldi, out, ldo, rows, cols, s, dshift);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}
