// Second-order-cone barrier pieces (FunctionManagerSOCP / FunctionManagerSOCPPhase1, FunctionManager.py:834-1460)
// in the stacked form of SURVEY.md K2:  all cone rows A_i live in one row-major matrix (cone i = rows
// cone_off[i] .. cone_off[i+1]), so slacks are one GEMV, the per-cone gradient rows g_i are one segmented
// transposed GEMV and the Hessian is ONE weighted SYRK over  W = [A rows ; c_i rows ; g_i rows]  with weights
// [2/(s_i+eps) per A row ; 2/(s_i+eps) ; 1]  -- no per-cone n x n cache (the reference keeps A_i^T A_i per cone).
// The reference's quirks are kept: "+ c c^T" (Q6), eps = 1e-12 in gradient/Hessian of the main phase but 1e-15
// in the objective and in phase-I (Q5).
#include "common.cuh"

using namespace ipm;

constexpr double LOG_GUARD = 1e-15;

// One CTA per cone.  slack layout: [M cone slacks | bound slacks (written elsewhere) | M right-hand sides].
//   s_i = rhs_i^2 - |lhs_i|^2 (+ s in phase-I);   inv_i = 1/(s_i + guard)
//   wts   : SYRK weights for the A rows and the c row of cone i (2*inv_i); the g rows get weight 1
//   coefA : 2*inv_i*lhs_r  (per A row)      coefC : -2*inv_i*rhs_i      =>  g_i = sum coefA_r A_r + coefC_i c_i
//   per-cone reduction inputs: plog[i] = log(s_i + 1e-15), pinv[i] = 1/(s_i + 1e-15)
__global__ void __launch_bounds__(128)
cone_slack_kernel(int M, const int* __restrict__ cone_off, const double* __restrict__ lhs,
                  const double* __restrict__ rhs, const double* __restrict__ s_ptr, double guard, int tail_off,
                  int ktot, double* __restrict__ slacks, double* __restrict__ inv_c, double* __restrict__ wts,
                  double* __restrict__ coefA, double* __restrict__ coefC, double* __restrict__ plog,
                  double* __restrict__ pinv) {
  __shared__ double red[32];
  __shared__ double bc;
  const int i = blockIdx.x;
  const int r0 = cone_off[i], r1 = cone_off[i + 1];
  double acc = 0.0;
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) acc = fma(lhs[r], lhs[r], acc);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const double rh = rhs[i];
    double sl = rh * rh - acc;
    if (s_ptr) sl += *s_ptr;
    const double iv = 1.0 / (sl + guard);
    slacks[i] = sl;
    slacks[tail_off + i] = rh;
    inv_c[i] = iv;
    coefC[i] = -2.0 * iv * rh;
    wts[ktot + i] = 2.0 * iv;
    wts[ktot + M + i] = 1.0;
    plog[i] = log(sl + LOG_GUARD);
    pinv[i] = 1.0 / (sl + LOG_GUARD);
    bc = iv;
  }
  __syncthreads();
  const double iv = bc;
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    wts[r] = 2.0 * iv;
    coefA[r] = 2.0 * iv * lhs[r];
  }
}

// Single CTA: folds the per-cone values and the bound-part reductions (from ipm_lin_barrier_eval_f64 with m = 0,
// or NULL) into red_out = [sum log, min slack (incl. the rhs entries), sum inv, sum inv^2, #negative].
__global__ void __launch_bounds__(256)
cone_reduce_kernel(int M, int tail_off, const double* __restrict__ slacks, const double* __restrict__ plog,
                   const double* __restrict__ pinv, const double* __restrict__ red_bounds,
                   double* __restrict__ red_out) {
  __shared__ double red[32];
  double sl = 0.0, mn = INFINITY, si = 0.0, si2 = 0.0, ng = 0.0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    sl += plog[i];
    si += pinv[i];
    si2 += pinv[i] * pinv[i];
    const double a = slacks[i], b = slacks[tail_off + i];
    mn = fmin(mn, fmin(a, b));
    ng += (a < 0.0 ? 1.0 : 0.0) + (b < 0.0 ? 1.0 : 0.0);
  }
  sl = block_sum(sl, red);
  si = block_sum(si, red);
  si2 = block_sum(si2, red);
  ng = block_sum(ng, red);
  mn = block_min(mn, red);
  if (threadIdx.x == 0) {
    if (red_bounds) {
      sl += red_bounds[0];
      mn = fmin(mn, red_bounds[1]);
      si += red_bounds[2];
      si2 += red_bounds[3];
      ng += red_bounds[4];
    }
    red_out[0] = sl; red_out[1] = mn; red_out[2] = si; red_out[3] = si2; red_out[4] = ng;
  }
}

extern "C" int ipm_cone_eval_f64(int M, const int* cone_off, int ktot, const double* lhs, const double* rhs,
                                 const double* s_ptr, double guard, int tail_off, double* slacks, double* inv_c,
                                 double* wts, double* coefA, double* coefC, double* plog, double* pinv,
                                 const double* red_bounds, double* red_out, void* stream) {
  if (M <= 0 || !cone_off || !lhs || !rhs || !slacks || !inv_c || !wts || !coefA || !coefC || !plog || !pinv ||
      !red_out)
    return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cone_slack_kernel<<<M, 128, 0, st>>>(M, cone_off, lhs, rhs, s_ptr, guard, tail_off, ktot, slacks, inv_c, wts,
                                       coefA, coefC, plog, pinv);
  IPM_LAUNCH_CHECK();
  cone_reduce_kernel<<<1, 256, 0, st>>>(M, tail_off, slacks, plog, pinv, red_bounds, red_out);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// G[i][j] = sum_{r in cone i} coefA[r] * A[r][j] + coefC[i] * Cc[i][j]      (one pass over the stacked A)
__global__ void __launch_bounds__(128)
cone_grad_rows_kernel(int n, const int* __restrict__ cone_off, const double* __restrict__ A, long long lda,
                      const double* __restrict__ Cc, long long ldc, const double* __restrict__ coefA,
                      const double* __restrict__ coefC, double* __restrict__ G, long long ldg) {
  const int i = blockIdx.y;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (j >= n) return;
  const int r0 = cone_off[i], r1 = cone_off[i + 1];
  const bool pair = j + 1 < n;
  const bool vec = pair && !(lda & 1) && !(((uintptr_t)A) & 15);
  double a0 = 0.0, a1 = 0.0;
  if (vec) {
    int r = r0;
    for (; r + 3 < r1; r += 4) {
      const double2 m0 = __ldcs(reinterpret_cast<const double2*>(A + (long long)r * lda + j));
      const double2 m1 = __ldcs(reinterpret_cast<const double2*>(A + (long long)(r + 1) * lda + j));
      const double2 m2 = __ldcs(reinterpret_cast<const double2*>(A + (long long)(r + 2) * lda + j));
      const double2 m3 = __ldcs(reinterpret_cast<const double2*>(A + (long long)(r + 3) * lda + j));
      const double c0 = coefA[r], c1 = coefA[r + 1], c2 = coefA[r + 2], c3 = coefA[r + 3];
      a0 = fma(c0, m0.x, a0); a1 = fma(c0, m0.y, a1);
      a0 = fma(c1, m1.x, a0); a1 = fma(c1, m1.y, a1);
      a0 = fma(c2, m2.x, a0); a1 = fma(c2, m2.y, a1);
      a0 = fma(c3, m3.x, a0); a1 = fma(c3, m3.y, a1);
    }
    for (; r < r1; ++r) {
      const double2 m0 = __ldcs(reinterpret_cast<const double2*>(A + (long long)r * lda + j));
      a0 = fma(coefA[r], m0.x, a0); a1 = fma(coefA[r], m0.y, a1);
    }
  } else {
    for (int r = r0; r < r1; ++r) {
      a0 = fma(coefA[r], A[(long long)r * lda + j], a0);
      if (pair) a1 = fma(coefA[r], A[(long long)r * lda + j + 1], a1);
    }
  }
  const double cc = coefC[i];
  G[(long long)i * ldg + j] = fma(cc, Cc[(long long)i * ldc + j], a0);
  if (pair) G[(long long)i * ldg + j + 1] = fma(cc, Cc[(long long)i * ldc + j + 1], a1);
}

extern "C" int ipm_cone_grad_rows_f64(int M, int n, const int* cone_off, const double* A, int lda, const double* Cc,
                                      int ldc, const double* coefA, const double* coefC, double* G, int ldg,
                                      void* stream) {
  if (M <= 0 || n <= 0 || !cone_off || !A || !Cc || !coefA || !coefC || !G) return IPM_ERR_ARG;
  dim3 grid(ceil_div(n, 256), M);
  cone_grad_rows_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(n, cone_off, A, lda, Cc, ldc, coefA, coefC, G, ldg);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// Line-search polynomial of the cone slacks along x + a*dx:
//   s_i(a) = (rhs_i + a*drhs_i)^2 - |lhs_i + a*dlhs_i|^2 (+ s + a*ds)  =  s_i + a*p1_i + a^2*p2_i
//   p1_i = 2*(rhs_i*drhs_i - lhs_i.dlhs_i) (+ ds),   p2_i = drhs_i^2 - |dlhs_i|^2
// and of the right-hand sides (feasibility only):  rhs_i + a*drhs_i.
__global__ void __launch_bounds__(128)
cone_ls_coeffs_kernel(int M, const int* __restrict__ cone_off, const double* __restrict__ lhs,
                      const double* __restrict__ rhs, const double* __restrict__ dlhs,
                      const double* __restrict__ drhs, const double* __restrict__ ds_ptr, int tail_off,
                      double* __restrict__ p1, double* __restrict__ p2) {
  __shared__ double red[32];
  const int i = blockIdx.x;
  const int r0 = cone_off[i], r1 = cone_off[i + 1];
  double ld = 0.0, dd = 0.0;
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
    ld = fma(lhs[r], dlhs[r], ld);
    dd = fma(dlhs[r], dlhs[r], dd);
  }
  ld = block_sum(ld, red);
  dd = block_sum(dd, red);
  if (threadIdx.x == 0) {
    const double rh = rhs[i], dr = drhs[i];
    p1[i] = 2.0 * (rh * dr - ld) + (ds_ptr ? *ds_ptr : 0.0);
    p2[i] = dr * dr - dd;
    p1[tail_off + i] = dr;
    p2[tail_off + i] = 0.0;
  }
}

extern "C" int ipm_cone_ls_coeffs_f64(int M, const int* cone_off, const double* lhs, const double* rhs,
                                      const double* dlhs, const double* drhs, const double* ds_ptr, int tail_off,
                                      double* p1, double* p2, void* stream) {
  if (M <= 0 || !cone_off || !lhs || !rhs || !dlhs || !drhs || !p1 || !p2) return IPM_ERR_ARG;
  cone_ls_coeffs_kernel<<<M, 128, 0, (cudaStream_t)stream>>>(M, cone_off, lhs, rhs, dlhs, drhs, ds_ptr, tail_off, p1,
                                                           p2);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}
