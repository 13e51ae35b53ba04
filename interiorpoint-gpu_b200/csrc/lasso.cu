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
//
// Tail rows: n = 512 + 1 (bias) would cost a fifth tile row of which 1/128 is useful and a second wave on the
// 148 SMs.  Rows beyond the last full 128-row tile (when there are at most 8 of them) are instead computed by a
// few extra CTAs of the same launch as plain dot products (thread = column), so cfg-5 is ONE wave: 128 DMMA
// tiles + 16 tail CTAs.
#include "common.cuh"
#include "gemm_tn_core.cuh"
#include "tensormap.cuh"

namespace ipm {

constexpr int TAIL_MAX = 8;      // tail rows handled by the dot-product CTAs
constexpr int TAIL_COLS = 256;   // columns per tail CTA
constexpr int WARPS_PER_CTA = gemm::THREADS / 32;

struct LassoEpilogue {
  const double* bA;
  const double* eta;   // per column
  double* alpha;
  double* u;
  double* z_out;
  const double* Qt;    // tail rows only
  const double* z_in;  // tail rows only
  long long ld;        // common leading dimension of bA / alpha / u / z
  long long ldq;
  int n, K, n_main;    // rows [0, n_main) by DMMA tiles, [n_main, n) by tail CTAs
  int k_main;          // contraction rows [0, k_main) by DMMA k-tiles, [k_main, n) as rank-1 terms in the epilogue
  double rho;
  int add_bias, positive, want_norms;
  double* partials;    // [gridDim.x][WARPS_PER_CTA][4]

  struct Sums {
    double r, d, a, u;
  };

  __device__ __forceinline__ void update(int row, int col, double xacc, double b, double uo, double ao, double& an,
                                         double& un, Sums& s) const {
    const double x = b + xacc;
    const double v = x + uo;
    const double et = eta[col];
    an = fmax(v - et, 0.0);
    if (!positive) an -= fmax(-v - et, 0.0);
    if (add_bias && row == 0) an = v;
    un = uo + x - an;
    if (want_norms) {
      const double r = x - an, dd = rho * (an - ao);
      s.r = fma(r, r, s.r);
      s.d = fma(dd, dd, s.d);
      s.a = fma(an, an, s.a);
      s.u = fma(un, un, s.u);
    }
  }

  __device__ __forceinline__ void flush(Sums s) const {
    if (!want_norms) return;
    s.r = warp_sum(s.r);
    s.d = warp_sum(s.d);
    s.a = warp_sum(s.a);
    s.u = warp_sum(s.u);
    if ((threadIdx.x & 31) == 0) {
      double* p = partials + ((long long)blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5)) * 4;
      // += : a persistent CTA handles several tiles (the slots are zeroed before every launch that wants norms)
      p[0] += s.r; p[1] += s.d; p[2] += s.a; p[3] += s.u;
    }
  }

  template <int MI, int NI>
  __device__ __forceinline__ void tile(double (&acc)[MI][NI][2], int m_base, int n_base, int g8, int l4) const {
    static_assert(MI == 8 && NI == 4, "LassoEpilogue is written for the 128x128 CTA tile");
    Sums s{0.0, 0.0, 0.0, 0.0};
    const bool interior = (m_base + 64 <= n_main) && (n_base + 32 <= K);
    // contraction rows beyond the last full k-tile (n = 513: the bias row, which would otherwise cost a 33rd k-tile
    // of 16 rows for one): acc[i][jn][e] += sum_{k >= k_main} Q~[k][row_i] * z[k][col]   (in place: no registers)
    for (int k = k_main; k < n; ++k) {
      double qk[MI], zk[NI][2];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int row = m_base + i * 8 + g8;
        qk[i] = row < n_main ? __ldg(Qt + (long long)k * ldq + row) : 0.0;
      }
#pragma unroll
      for (int jn = 0; jn < NI; ++jn)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = n_base + jn * 8 + 2 * l4 + e;
          zk[jn][e] = col < K ? z_in[(long long)k * ld + col] : 0.0;
        }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int jn = 0; jn < NI; ++jn) {
          acc[i][jn][0] = fma(qk[i], zk[jn][0], acc[i][jn][0]);
          acc[i][jn][1] = fma(qk[i], zk[jn][1], acc[i][jn][1]);
        }
    }
    if (interior) {
      // batched 16-byte loads: all inputs of two 8-row groups first, then the math, then the stores
#pragma unroll
      for (int ib = 0; ib < 8; ib += 2) {
        double2 vb[2][4], vu[2][4], va[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int jn = 0; jn < 4; ++jn) {
            const long long idx = (long long)(m_base + (ib + i) * 8 + g8) * ld + n_base + jn * 8 + 2 * l4;
            vb[i][jn] = *reinterpret_cast<const double2*>(bA + idx);
            vu[i][jn] = *reinterpret_cast<const double2*>(u + idx);
            va[i][jn] = *reinterpret_cast<const double2*>(alpha + idx);
          }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int jn = 0; jn < 4; ++jn) {
            const int row = m_base + (ib + i) * 8 + g8, col = n_base + jn * 8 + 2 * l4;
            const long long idx = (long long)row * ld + col;
            double2 an, un;
            update(row, col, acc[ib + i][jn][0], vb[i][jn].x, vu[i][jn].x, va[i][jn].x, an.x, un.x, s);
            update(row, col + 1, acc[ib + i][jn][1], vb[i][jn].y, vu[i][jn].y, va[i][jn].y, an.y, un.y, s);
            *reinterpret_cast<double2*>(alpha + idx) = an;
            *reinterpret_cast<double2*>(u + idx) = un;
            *reinterpret_cast<double2*>(z_out + idx) = make_double2(un.x - an.x, un.y - an.y);
          }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = m_base + i * 8 + g8;
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = n_base + jn * 8 + 2 * l4 + e;
            if (row < n_main && col < K) {
              const long long idx = (long long)row * ld + col;
              double an, un;
              update(row, col, acc[i][jn][e], bA[idx], u[idx], alpha[idx], an, un, s);
              alpha[idx] = an;
              u[idx] = un;
              z_out[idx] = un - an;
            }
          }
        }
      }
    }
    flush(s);
  }

  // Hooks of the one-CTA-per-tile kernel.  Both were tried for this epilogue on B200 and measured slower than doing
  // nothing: an L2 prefetch of bA / u / alpha at CTA start (K = 512 shard: 85 -> 91 us per iteration), and folding
  // the tail row into the tile CTAs instead of the 16 extra CTAs below (104 -> 118 us: it serialises behind the tile).
  __device__ __forceinline__ void prefetch(int, int, int, int) const {}
  __device__ __forceinline__ void after_tile(int, int, int, int) const {}

  // tail rows [n_main, n): x[r][c] = bA + sum_k Q~[k][r] z[k][c]   (thread = column, coalesced z reads)
  __device__ __forceinline__ void extra(int bid) const {
    Sums s{0.0, 0.0, 0.0, 0.0};
    const int c = bid * TAIL_COLS + threadIdx.x;
    const int nt = n - n_main;
    if (threadIdx.x < TAIL_COLS && c < K) {
      double acc[TAIL_MAX];
#pragma unroll
      for (int q = 0; q < TAIL_MAX; ++q) acc[q] = 0.0;
      int k = 0;
      for (; k + 15 < n; k += 16) {  // 16 independent loads in flight per thread
        double zk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) zk[j] = z_in[(long long)(k + j) * ld + c];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const double* qrow = Qt + (long long)(k + j) * ldq + n_main;
#pragma unroll
          for (int q = 0; q < TAIL_MAX; ++q)
            if (q < nt) acc[q] = fma(__ldg(qrow + q), zk[j], acc[q]);
        }
      }
      for (; k < n; ++k) {
        const double zk = z_in[(long long)k * ld + c];
        const double* qrow = Qt + (long long)k * ldq + n_main;
#pragma unroll
        for (int q = 0; q < TAIL_MAX; ++q)
          if (q < nt) acc[q] = fma(__ldg(qrow + q), zk, acc[q]);
      }
#pragma unroll
      for (int q = 0; q < TAIL_MAX; ++q) {
        if (q < nt) {
          const int row = n_main + q;
          const long long idx = (long long)row * ld + c;
          double an, un;
          update(row, c, acc[q], bA[idx], u[idx], alpha[idx], an, un, s);
          alpha[idx] = an;
          u[idx] = un;
          z_out[idx] = un - an;
        }
      }
    }
    flush(s);
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
  double sq0 = 0.0, sq1 = 0.0, sq2 = 0.0, sq3 = 0.0, l1 = 0.0;
  int r = 0;
  for (; r + 3 < m; r += 4) {
    const double v0 = R[(long long)r * ldr + c], v1 = R[(long long)(r + 1) * ldr + c];
    const double v2 = R[(long long)(r + 2) * ldr + c], v3 = R[(long long)(r + 3) * ldr + c];
    sq0 = fma(v0, v0, sq0); sq1 = fma(v1, v1, sq1); sq2 = fma(v2, v2, sq2); sq3 = fma(v3, v3, sq3);
  }
  for (; r < m; ++r) {
    const double v = R[(long long)r * ldr + c];
    sq0 = fma(v, v, sq0);
  }
  for (int j = add_bias ? 1 : 0; j < n; ++j) {
    const double a = alpha[(long long)j * lda + c];
    l1 += positive ? a : fabs(a);
  }
  out[c] = ((sq0 + sq1) + (sq2 + sq3)) / (2.0 * m) + reg[c] * l1;
}

static inline int lasso_n_main(int n) {
  const int tail = n % gemm::BM;
  return (n >= gemm::BM && tail > 0 && tail <= TAIL_MAX) ? n - tail : n;
}

}  // namespace ipm

using namespace ipm;

extern "C" int ipm_internal_sk_slot(void* stream, double** partials, unsigned int** flags, unsigned int* epoch,
                                    int* num_sms);

extern "C" long long ipm_lasso_partials_doubles(int n, int K) {
  const int n_main = lasso_n_main(n);
  long long ctas = (long long)ceil_div(n_main, gemm::BM) * ceil_div(K, gemm::BN);
  if (ctas < 256) ctas = 256;  // the stream-K path launches up to one CTA per SM however few tiles there are
  ctas += n_main < n ? ceil_div(K, TAIL_COLS) : 0;
  return ctas * WARPS_PER_CTA * 4;
}

// One ADMM iteration for all K problems.  Qt: n x n (ldq), z_in/z_out/bA/alpha/u: n x K (ld).  z_out != z_in.
// If want_norms, norms_out[4] receives the squared Frobenius norms {|x-alpha+|, |rho(alpha+-alpha)|, |alpha+|, |u+|}.
extern "C" int ipm_lasso_admm_step_f64(const double* Qt, int ldq, int n, int K, const double* bA, const double* eta,
                                       double rho, double* alpha, double* u, const double* z_in, double* z_out,
                                       int ld, int add_bias, int positive, int want_norms, double* partials,
                                       double* norms_out, void* stream) {
  if (!Qt || !bA || !eta || !alpha || !u || !z_in || !z_out || z_in == z_out || n <= 0 || K <= 0 || ldq < n ||
      ld < K || (ld & 1) || (want_norms && (!partials || !norms_out)))
    return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int n_main = lasso_n_main(n);
  CUtensorMap tmA, tmB;
  int rc = make_operand_map(&tmA, Qt, ldq, n, n);  // Q~ symmetric: Q~[k][i] is the A operand (k rows, i columns)
  if (rc) return rc;
  rc = make_operand_map(&tmB, z_in, ld, n, K);
  if (rc) return rc;
  const int tiles = ceil_div(n_main, gemm::BM) * ceil_div(K, gemm::BN);
  const int tail_ctas = n_main < n ? ceil_div(K, TAIL_COLS) : 0;
  // a k-tile holding <= 4 contraction rows is cheaper as rank-1 terms in the epilogue than as 128 DMMAs per warp
  const int k_tail = n % gemm::BK;
  const int k_main = (n > gemm::BK && k_tail > 0 && k_tail <= 4) ? n - k_tail : n;
  LassoEpilogue epi{bA, eta, alpha, u, z_out, Qt, z_in, ld, ldq, n, K, n_main, k_main, rho, add_bias, positive, want_norms,
                    partials};
  auto kern = gemm::gemm_tn_kernel<false, LassoEpilogue>;
  auto pkern = gemm::gemm_tn_persistent_kernel<false, LassoEpilogue>;
  static bool attr_set[kMaxDevices], pattr_set[kMaxDevices];
  IPM_CUDA_CHECK(ensure_dynamic_smem(kern, gemm::SMEM_BYTES, attr_set));
  IPM_CUDA_CHECK(ensure_dynamic_smem(pkern, gemm::SMEM_BYTES, pattr_set));
  // Small batches (a per-GPU shard of K = 512 is 16 tiles on 148 SMs): spread the k-tiles of every tile over all SMs
  // with the persistent stream-K kernel; the tail-row CTAs stay co-resident behind the tile CTAs.
  const int ktiles = ceil_div(k_main, gemm::BK);
  double* sk_partials = nullptr;
  unsigned int *sk_flags = nullptr, sk_epoch = 0;
  int sms = 0;
  rc = ipm_internal_sk_slot(stream, &sk_partials, &sk_flags, &sk_epoch, &sms);
  if (rc) return rc;
  const int G = sms - tail_ctas;
  const bool few = sk_partials && G > 0 && 4 * tiles <= 3 * G && ktiles >= 2;
  const int grid_ctas = few ? (G < tiles * ktiles ? G : tiles * ktiles) : tiles;
  const long long nparts = (long long)(grid_ctas + tail_ctas) * WARPS_PER_CTA * 4;
  if (want_norms) IPM_CUDA_CHECK(cudaMemsetAsync(partials, 0, sizeof(double) * nparts, st));
  // M = n_main rows through the DMMA tiles, contraction rows [0, k_main) through the k-tiles
  // programmatic dependent launch: iteration i+1's CTAs are scheduled while iteration i drains (they wait in
  // pdl_wait() before touching z / alpha / u)
  if (few) {
    gemm::StreamK sk{sk_partials, sk_flags, sk_epoch, grid_ctas, grid_ctas, ipm_internal_fault_word()};
    IPM_CUDA_CHECK(launch_pdl(pkern, dim3(grid_ctas + tail_ctas), dim3(gemm::THREADS), gemm::SMEM_BYTES, st, tmA, tmB,
                              n_main, K, k_main, (const double*)nullptr, 0, epi, sk));
  } else {
    IPM_CUDA_CHECK(launch_pdl(kern, dim3(tiles + tail_ctas), dim3(gemm::THREADS), gemm::SMEM_BYTES, st, tmA, tmB, n_main,
                              K, k_main, (const double*)nullptr, 0, epi));
  }
  IPM_LAUNCH_CHECK();
  if (want_norms) {
    lasso_norms_kernel<<<1, 256, 0, st>>>(partials, (int)(nparts / 4), norms_out);
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
