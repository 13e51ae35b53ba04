// Conjugate-gradient Newton solves (the reference's secondary ``linear_solve_method="cg"``):
//   NewtonSolverCG.newton_linear_solve           NewtonSolver.py:365-400     cg(-H, gradf, x0, maxiter=max_cg_iters)
//   PhaseOne.PhaseOneSolver(linear_solver="cg")  PhaseOne.py:137-150         cg(hess, -grad, x0=[x, s], maxiter)
// both through scipy.sparse.linalg.cg / cupyx...cg (un-vendored SciPy; its published algorithm is restated here):
//
//     atol = rtol * ||b||  (rtol = 1e-5);  r = b - A x0
//     for it in range(maxiter):
//         if ||r|| < atol: return x
//         rho = r.r;  p = r + (rho / rho_prev) p   (p = r at it = 0)
//         q = A p;  alpha = rho / (p.q);  x += alpha p;  r -= alpha q
//     return x
//
// A = sign * H with H symmetric, given DENSE (both triangles; ipm_symmetrize_upper_f64 mirrors the engine's upper-stored
// Hessian first), so the product q = H p is the HBM-streaming row GEMV of blas.cu.  Everything else -- dot products,
// vector updates, the convergence test -- is two single-CTA kernels per iteration that keep their scalars on the
// device; once the test holds the remaining launches are no-ops, so the host enqueues all `maxiter` iterations
// without a single read-back.  Fixed summation order: bit-reproducible run to run.
#include "common.cuh"

using namespace ipm;

extern "C" int ipm_gemv_n_f64(const double* Mx, int ld, int rows, int cols, const double* x, double* y, double alpha,
                              double beta, void* stream);

namespace {
constexpr int CG_THREADS = 1024;
// state: [0] rho_prev  [1] atol  [2] done (0/1)  [3] iterations performed  [4] rho_cur

__global__ void __launch_bounds__(256) symmetrize_upper_kernel(double* __restrict__ H, long long ld, int n) {
  // 32 x 32 tiles strictly below the diagonal block row: H[j][i] = H[i][j] for j > i
  __shared__ double tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;  // source tile (rows bi, cols bj), bj >= bi
  if (bj < bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + r, j = bj * 32 + tx;
    tile[r][tx] = (i < n && j < n) ? H[(long long)i * ld + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int j = bj * 32 + r, i = bi * 32 + tx;  // destination (row j, col i)
    if (j < n && i < n && j > i) H[(long long)j * ld + i] = tile[tx][r];
  }
}

// r = b - sign * (H x0) ;  atol = rtol * ||b|| ;  done = (||b|| == 0)
__global__ void __launch_bounds__(CG_THREADS, 1)
cg_init_kernel(int n, const double* __restrict__ b, const double* __restrict__ Hx0, double sign, double rtol,
               double* __restrict__ r, double* __restrict__ state) {
  __shared__ double red[32];
  double bb = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double bi = b[i];
    bb = fma(bi, bi, bb);
    r[i] = Hx0 ? bi - sign * Hx0[i] : bi;
  }
  bb = block_sum(bb, red);
  if (threadIdx.x == 0) {
    const double bn = sqrt(bb);
    state[0] = 0.0;
    state[1] = rtol * bn;
    state[2] = bn == 0.0 ? 1.0 : 0.0;
    state[3] = 0.0;
    state[4] = 0.0;
  }
}

// top of an iteration: convergence test, rho, search direction
__global__ void __launch_bounds__(CG_THREADS, 1)
cg_direction_kernel(int n, int it, const double* __restrict__ r, double* __restrict__ p, double* __restrict__ state) {
  __shared__ double red[32];
  __shared__ double s_beta;
  __shared__ int s_done;
  if (state[2] != 0.0) return;
  double rr = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) rr = fma(r[i], r[i], rr);
  rr = block_sum(rr, red);
  if (threadIdx.x == 0) {
    s_done = sqrt(rr) < state[1];
    s_beta = it > 0 ? rr / state[0] : 0.0;
    if (s_done) state[2] = 1.0;
    state[4] = rr;
  }
  __syncthreads();
  if (s_done) return;
  const double beta = s_beta;
  for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = it > 0 ? fma(beta, p[i], r[i]) : r[i];
}

// after q = H p:  alpha = rho / (p . sign q);  x += alpha p;  r -= alpha sign q
__global__ void __launch_bounds__(CG_THREADS, 1)
cg_update_kernel(int n, double sign, const double* __restrict__ p, const double* __restrict__ q, double* __restrict__ x,
                 double* __restrict__ r, double* __restrict__ state) {
  __shared__ double red[32];
  __shared__ double s_alpha;
  if (state[2] != 0.0) return;
  double pq = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) pq = fma(p[i], sign * q[i], pq);
  pq = block_sum(pq, red);
  if (threadIdx.x == 0) {
    s_alpha = state[4] / pq;
    state[0] = state[4];
    state[3] += 1.0;
  }
  __syncthreads();
  const double alpha = s_alpha;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    x[i] = fma(alpha, p[i], x[i]);
    r[i] = fma(-alpha, sign * q[i], r[i]);
  }
}

// x0 of NewtonSolverCG (NewtonSolver.py:379-383): dc = x.g;  x0 = dc < 0 ? -dc * x / (x . H x) : 0;  also Hx0 = scale * Hx
__global__ void __launch_bounds__(CG_THREADS, 1)
cg_descent_x0_kernel(int n, const double* __restrict__ x, const double* __restrict__ g, const double* __restrict__ Hx,
                     double* __restrict__ x0, double* __restrict__ Hx0) {
  __shared__ double red[32];
  __shared__ double s_scale;
  double dc = 0.0, xhx = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    dc = fma(x[i], g[i], dc);
    xhx = fma(x[i], Hx[i], xhx);
  }
  dc = block_sum(dc, red);
  xhx = block_sum(xhx, red);
  if (threadIdx.x == 0) s_scale = dc < 0.0 ? -dc / xhx : 0.0;
  __syncthreads();
  const double sc = s_scale;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    x0[i] = sc * x[i];
    Hx0[i] = sc * Hx[i];
  }
}
}  // namespace

// H[j][i] = H[i][j] for j > i: the engine stores only the upper triangle of its Hessians.
extern "C" int ipm_symmetrize_upper_f64(double* H, int ld, int n, void* stream) {
  if (!H || n <= 0 || ld < n) return IPM_ERR_ARG;
  const int t = ceil_div(n, 32);
  symmetrize_upper_kernel<<<dim3(t, t), 256, 0, (cudaStream_t)stream>>>(H, ld, n);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

extern "C" long long ipm_cg_ws_doubles(int n) { return n <= 0 ? 0 : 3ll * n + 8; }

// x0 of the reference's CG Newton step (see cg_descent_x0_kernel).  Hx = H x (caller: ipm_gemv_n_f64).
extern "C" int ipm_cg_descent_x0_f64(int n, const double* x, const double* g, const double* Hx, double* x0, double* Hx0,
                                     void* stream) {
  if (n <= 0 || !x || !g || !Hx || !x0 || !Hx0) return IPM_ERR_ARG;
  cg_descent_x0_kernel<<<1, CG_THREADS, 0, (cudaStream_t)stream>>>(n, x, g, Hx, x0, Hx0);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// Conjugate gradients for (sign * H) x = b, scipy.sparse.linalg.cg semantics (rtol on ||b||, at most maxiter steps,
// the last iterate is returned whether or not it converged).  H: dense symmetric n x n (ld);  x: x0 in, solution out;
// Hx0: H x0 or NULL when x0 = 0;  ws: ipm_cg_ws_doubles(n) doubles -- ws[3n + 2] != 0 if converged, ws[3n + 3] =
// iterations performed.  No host synchronisation.
extern "C" int ipm_cg_solve_f64(const double* H, int ld, int n, const double* b, double* x, const double* Hx0, double sign,
                                int maxiter, double rtol, double* ws, void* stream) {
  if (!H || !b || !x || !ws || n <= 0 || ld < n || maxiter < 0 || (sign != 1.0 && sign != -1.0)) return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  double *r = ws, *p = ws + n, *q = ws + 2ll * n, *state = ws + 3ll * n;
  cg_init_kernel<<<1, CG_THREADS, 0, st>>>(n, b, Hx0, sign, rtol, r, state);
  IPM_LAUNCH_CHECK();
  for (int it = 0; it < maxiter; ++it) {
    cg_direction_kernel<<<1, CG_THREADS, 0, st>>>(n, it, r, p, state);
    IPM_LAUNCH_CHECK();
    int rc = ipm_gemv_n_f64(H, ld, n, n, p, q, 1.0, 0.0, stream);
    if (rc) return rc;
    cg_update_kernel<<<1, CG_THREADS, 0, st>>>(n, sign, p, q, x, r, state);
    IPM_LAUNCH_CHECK();
  }
  return IPM_OK;
}
