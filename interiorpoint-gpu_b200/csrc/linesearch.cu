// Backtracking line searches, entirely on the device (no host round trip per trial).
//
// The reference walks  a <- beta * a  in a Python loop and recomputes the slack GEMV for every trial
// (NewtonSolver.py:157-206, NewtonSolverInfeasibleStart.py:170-273) -- 60-96 % of its GPU wall time.  Here the
// slacks along the ray are a polynomial per entry,  s_i(a) = s0_i + a*p1_i + a^2*p2_i  (p2 = 0 for linear
// constraints, quadratic for second-order cones), so once C*dx is known every trial costs O(#slacks):
//   1. ls_feas   : k_i = first table index whose step keeps entry i feasible; kmax = max_i k_i (atomicMax --
//                  feasibility is monotone in a, so the max reproduces the sequential "shrink until feasible")
//   2. ls_armijo : one CTA; freezes L = sum log(s(a_kmax) + 1e-15) and replays the reference's Armijo loop with
//                  its quirks (slope g.x, frozen barrier term, evaluated point lagging one beta)
//   3. ls_residual: one CTA; residual-norm search of the infeasible-start method (Q4)
// `table[k]` holds the reference's step sequence 1, beta, beta*beta, ... built by repeated multiplication on
// the host, up to and including the first entry below 1e-13 (index len-1 = "stuck").
#include "common.cuh"

using namespace ipm;

constexpr double LOG_GUARD = 1e-15;
constexpr double STUCK = 1e-13;

__device__ __forceinline__ double trial_slack(double s0, double p1, double p2, double a) {
  // explicit roundings so the feasibility test and the frozen log-sum see the same number
  double v = __dadd_rn(s0, __dmul_rn(a, p1));
  if (p2 != 0.0) v = __dadd_rn(v, __dmul_rn(__dmul_rn(a, a), p2));
  return v;
}

__device__ __forceinline__ int first_feasible(double s0, double p1, double p2, const double* __restrict__ table,
                                              int len) {
  int k = 0;
  while (k < len - 1 && trial_slack(s0, p1, p2, table[k]) < 0.0) ++k;
  return k;
}

// Linear family: builds p1 from C*dx and dz, then searches.  Layout [m | ub | lb] as in barrier.cu.
__global__ void __launch_bounds__(256)
ls_feas_lin_kernel(int m, int n, const double* __restrict__ slacks, const double* __restrict__ Cdx,
                   const double* __restrict__ dz, int has_ub, int has_lb, int phase1,
                   const double* __restrict__ table, int len, double* __restrict__ p1_out, int* __restrict__ kmax) {
  const double dzs = phase1 ? dz[n] : 0.0;
  const int total = m + (has_ub ? n : 0) + (has_lb ? n : 0);
  int kloc = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    double p1;
    if (i < m) {
      p1 = dzs - Cdx[i];
    } else if (has_ub && i < m + n) {
      p1 = dzs - dz[i - m];
    } else {
      p1 = dzs + dz[i - m - (has_ub ? n : 0)];
    }
    p1_out[i] = p1;
    kloc = max(kloc, first_feasible(slacks[i], p1, 0.0, table, len));
  }
  kloc = warp_max_i(kloc);
  if ((threadIdx.x & 31) == 0 && kloc > 0) atomicMax(kmax, kloc);
}

extern "C" int ipm_ls_feas_lin_f64(int m, int n, const double* slacks, const double* Cdx, const double* dz,
                                   int has_ub, int has_lb, int phase1, const double* table, int len, double* p1_out,
                                   int* kmax, void* stream) {
  if (m < 0 || n <= 0 || !slacks || !dz || !table || len < 2 || !p1_out || !kmax || (m > 0 && !Cdx))
    return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  IPM_CUDA_CHECK(cudaMemsetAsync(kmax, 0, sizeof(int), st));
  const int total = m + (has_ub ? n : 0) + (has_lb ? n : 0);
  if (total == 0) return IPM_OK;
  int blocks = ceil_div(total, 256);
  if (blocks > 592) blocks = 592;
  ls_feas_lin_kernel<<<blocks, 256, 0, st>>>(m, n, slacks, Cdx, dz, has_ub, has_lb, phase1, table, len, p1_out,
                                             kmax);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// Generic polynomial entries (second-order cones): kmax = max(kmax, ...) -- does NOT reset kmax, so it can be
// chained after ipm_ls_feas_lin_f64 for problems with both cones and bounds.
__global__ void __launch_bounds__(256)
ls_feas_poly_kernel(int count, const double* __restrict__ s0, const double* __restrict__ p1,
                    const double* __restrict__ p2, const double* __restrict__ table, int len,
                    int* __restrict__ kmax) {
  int kloc = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    kloc = max(kloc, first_feasible(s0[i], p1[i], p2 ? p2[i] : 0.0, table, len));
  kloc = warp_max_i(kloc);
  if ((threadIdx.x & 31) == 0 && kloc > 0) atomicMax(kmax, kloc);
}

extern "C" int ipm_ls_feas_poly_f64(int count, const double* s0, const double* p1, const double* p2,
                                    const double* table, int len, int* kmax, int reset, void* stream) {
  if (count < 0 || !table || len < 2 || !kmax || (count > 0 && (!s0 || !p1))) return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (reset) IPM_CUDA_CHECK(cudaMemsetAsync(kmax, 0, sizeof(int), st));
  if (count == 0) return IPM_OK;
  int blocks = ceil_div(count, 256);
  if (blocks > 592) blocks = 592;
  ls_feas_poly_kernel<<<blocks, 256, 0, st>>>(count, s0, p1, p2, table, len, kmax);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// ------------------------------------------------------------------------------------------------
// Armijo search of the feasible-start Newton method (single CTA).
//   sumlog  : sum log(s(x)+1e-15) at the current point
//   terms   : [0] obj(x)   [1] d obj . dx (linear part)   [2] dx' P dx (0 if none)   [3] g . x
//   out     : [0] step  [1] stuck (0/1; 3 = trial point infeasible, raise kmax and retry)  [2] final table index
//             [3] frozen log-sum  [4] Armijo trials
// obj(x + a dx) = obj + a*terms[1] + 0.5*a*a*terms[2]   (exact for linear / quadratic objectives)
// L_direct / nneg (optional, second-order cones): the barrier log-sum and the count of negative slacks EVALUATED
// AT the trial point x + table[kmax]*dx by a full barrier evaluation.  Cone slacks rhs^2 - |lhs|^2 cancel
// catastrophically near the boundary, so the frozen term must come from the same formula the next iteration will
// use (that is what the reference does, NewtonSolver.py:172-183); the polynomial only proposes kmax.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
ls_armijo_kernel(int nc, const double* __restrict__ s0, const double* __restrict__ p1,
                 const double* __restrict__ p2, const double* __restrict__ table, int len,
                 int* __restrict__ kmax_ptr, const double* __restrict__ sumlog_ptr,
                 const double* __restrict__ terms, double t, double alpha,
                 int update_slacks_every, const double* __restrict__ L_direct,
                 const double* __restrict__ nneg, int textbook, int resume, double* __restrict__ out) {
  __shared__ double red[32];
  __shared__ double bcast;
  const int kstuck = len - 1;
  // resume (direct-evaluation barriers with update_slacks_every > 0): the previous launch stopped at a refresh, left
  // *kmax = index of the lagging point a_eval for the host's barrier evaluation and the attempt count in out[4]
  int k = resume ? *kmax_ptr + 1 : *kmax_ptr;
  if (resume) nneg = nullptr;
  if (k >= kstuck && !resume) {  // feasibility back-off ran out of steps (NewtonSolver.py:176-181)
    if (threadIdx.x == 0) {
      out[0] = table[kstuck]; out[1] = 1.0; out[2] = (double)kstuck; out[3] = NAN; out[4] = 0.0;
    }
    return;
  }
  if (nneg && *nneg > 0.0) {  // the proposed step leaves the domain when evaluated directly
    if (threadIdx.x == 0) {
      out[0] = table[k]; out[1] = 3.0; out[2] = (double)k; out[3] = NAN; out[4] = 0.0;
    }
    return;
  }
  auto logsum = [&](double a) -> double {
    double acc = 0.0;
    for (int i = threadIdx.x; i < nc; i += blockDim.x)
      acc += log(trial_slack(s0[i], p1[i], p2 ? p2[i] : 0.0, a) + LOG_GUARD);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) bcast = acc;
    __syncthreads();
    const double r = bcast;
    __syncthreads();
    return r;
  };
  // textbook != 0 (stand-alone PhaseOne.py:187-218): slope g.dx, barrier term re-evaluated at every trial, no lag
  const double sumlog0 = *sumlog_ptr, obj0 = terms[0], dobj = terms[1], quad = terms[2];
  const double gx = textbook ? terms[4] : terms[3];
  const double fx = t * obj0 - sumlog0;
  double a = table[k], a_eval = resume ? table[k - 1] : a;
  double L = L_direct ? *L_direct : logsum(a_eval);
  const double L_first = resume ? out[3] : L;
  int attempt = resume ? (int)out[4] : 0, stuck = 0;
  __syncthreads();  // out[3] / out[4] read by everybody before thread 0 rewrites them
  while (true) {
    const double objv = obj0 + a_eval * dobj + 0.5 * a_eval * a_eval * quad;
    const double lhs = t * objv - L;
    const double rhs = fx + alpha * a * gx;
    if (!(lhs > rhs)) break;
    ++attempt;
    if (a < STUCK) { stuck = 1; break; }
    a_eval = a;               // the point evaluated next lags the step by one beta (Q3)
    ++k;
    a = table[k];
    if (textbook) {
      a_eval = a;
      L = logsum(a_eval);
    } else if (update_slacks_every > 0 && (attempt % update_slacks_every == update_slacks_every - 1)) {
      if (L_direct) {  // the refreshed barrier term must come from a direct evaluation at x + a_eval dx: hand back
        stuck = 4;
        break;
      }
      L = logsum(a_eval);
    }
  }
  if (threadIdx.x == 0) {
    out[0] = a; out[1] = (double)stuck; out[2] = (double)k; out[3] = L_first; out[4] = (double)attempt;
    if (stuck == 4) *kmax_ptr = k - 1;
  }
}

extern "C" int ipm_ls_armijo_f64(int nc, const double* s0, const double* p1, const double* p2, const double* table,
                                 int len, int* kmax, const double* sumlog, const double* terms, double t,
                                 double alpha, int update_slacks_every, const double* L_direct,
                                 const double* nneg, int textbook, int resume, double* out, void* stream) {
  if (nc < 0 || !table || len < 2 || !kmax || !sumlog || !terms || !out || (nc > 0 && (!s0 || !p1)) ||
      (resume && !L_direct))
    return IPM_ERR_ARG;
  ls_armijo_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(nc, s0, p1, p2, table, len, kmax, sumlog, terms, t, alpha,
                                                        update_slacks_every, L_direct, nneg, textbook, resume, out);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// ------------------------------------------------------------------------------------------------
// Residual-norm search of the infeasible-start Newton method (single CTA).
//   r(a) = || [ u0 + a*u1 ; q0 + a*q1 ] ||     u: dual part (n),  q: primal part (p)
//   r0   = || [ g + A'v ; A x - b ] || = || [ r0d ; q0 ] ||
//   while r(a) > (1 - alpha*a) * r0:  a <- next table entry (stop when it drops below 1e-13)
//   out : [0] step  [1] stuck (0/1/2: 2 = stuck already in the feasibility back-off; 3 = trial point infeasible
//         when evaluated directly (nneg > 0): raise kmax and retry; 4 = update_slacks_every refresh due: *kmax holds the
//         index whose barrier gradient the host must evaluate into u0, then call again with resume = 1)  [2] index
//         [3] r0  [4] r(a) of the last evaluated trial  [5] attempts so far
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
ls_residual_kernel(int n, int p, const double* __restrict__ r0d, const double* __restrict__ u0,
                   const double* __restrict__ u1, const double* __restrict__ q0, const double* __restrict__ q1,
                   const double* __restrict__ table, int len, int* __restrict__ kmax_ptr, double alpha,
                   const double* __restrict__ nneg, int update_slacks_every, int resume, double* __restrict__ out) {
  __shared__ double red[32];
  __shared__ double bcast;
  const int kstuck = len - 1;
  int k = *kmax_ptr;
  int attempt = resume ? (int)out[5] : 0;
  if (resume) nneg = nullptr;  // the refreshed point lies inside the step that was already verified
  __syncthreads();
  if (k >= kstuck && !resume) {
    if (threadIdx.x == 0) {
      out[0] = table[kstuck]; out[1] = 2.0; out[2] = (double)kstuck; out[3] = NAN; out[4] = NAN;
    }
    return;
  }
  if (nneg && *nneg > 0.0) {
    if (threadIdx.x == 0) {
      out[0] = table[k]; out[1] = 3.0; out[2] = (double)k; out[3] = NAN; out[4] = NAN;
    }
    return;
  }
  auto bsum = [&](double acc) -> double {
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) bcast = acc;
    __syncthreads();
    const double r = bcast;
    __syncthreads();
    return r;
  };
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc = fma(r0d[i], r0d[i], acc);
  for (int i = threadIdx.x; i < p; i += blockDim.x) acc = fma(q0[i], q0[i], acc);
  const double r0 = sqrt(bsum(acc));
  auto rnorm = [&](double a) -> double {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const double v = u0[i] + a * u1[i];
      s = fma(v, v, s);
    }
    for (int i = threadIdx.x; i < p; i += blockDim.x) {
      const double v = q0[i] + a * q1[i];
      s = fma(v, v, s);
    }
    return sqrt(bsum(s));
  };
  double a = table[k];
  double rn = rnorm(a);
  int stuck = 0;
  while (rn > (1.0 - alpha * a) * r0) {
    ++attempt;
    ++k;
    a = table[k];
    if (a < STUCK) { stuck = 1; break; }
    if (update_slacks_every > 0 && attempt % update_slacks_every == update_slacks_every - 1) {
      stuck = 4;  // the barrier gradient is re-evaluated at this trial point (NewtonSolverInfeasibleStart.py:249-255)
      break;
    }
    rn = rnorm(a);
  }
  if (threadIdx.x == 0) {
    out[0] = a; out[1] = (double)stuck; out[2] = (double)k; out[3] = r0; out[4] = rn; out[5] = (double)attempt;
    if (stuck == 4) *kmax_ptr = k;
  }
}

extern "C" int ipm_ls_residual_f64(int n, int p, const double* r0d, const double* u0, const double* u1,
                                   const double* q0, const double* q1, const double* table, int len,
                                   int* kmax, double alpha, const double* nneg, int update_slacks_every, int resume,
                                   double* out, void* stream) {
  if (n <= 0 || p < 0 || !r0d || !u0 || !u1 || !table || len < 2 || !kmax || !out || (p > 0 && (!q0 || !q1)))
    return IPM_ERR_ARG;
  ls_residual_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(n, p, r0d, u0, u1, q0, q1, table, len, kmax, alpha, nneg,
                                                          update_slacks_every, resume, out);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// ------------------------------------------------------------------------------------------------
// small vector helpers used around the searches
// ------------------------------------------------------------------------------------------------
// out = x + (*a_dev) * dx   (trial point; rounding as NumPy: multiply, then add)
__global__ void trial_point_kernel(int n, const double* __restrict__ a_dev, const double* __restrict__ x,
                                   const double* __restrict__ dx, double* __restrict__ out) {
  const double a = *a_dev;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = __dadd_rn(x[i], __dmul_rn(a, dx[i]));
}

extern "C" int ipm_trial_point_f64(int n, const double* a_dev, const double* x, const double* dx, double* out,
                                   void* stream) {
  if (n <= 0 || !a_dev || !x || !dx || !out) return IPM_ERR_ARG;
  int blocks = ceil_div(n, 256);
  if (blocks > 1184) blocks = 1184;
  trial_point_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, a_dev, x, dx, out);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// out = ca*a + cb*b + cc*c   (b, c may be NULL)
__global__ void lincomb3_kernel(int n, double ca, const double* __restrict__ a, double cb,
                                const double* __restrict__ b, double cc, const double* __restrict__ c,
                                double* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double v = ca * a[i];
    if (b) v += cb * b[i];
    if (c) v += cc * c[i];
    out[i] = v;
  }
}

extern "C" int ipm_lincomb3_f64(int n, double ca, const double* a, double cb, const double* b, double cc,
                                const double* c, double* out, void* stream) {
  if (n <= 0 || !a || !out) return IPM_ERR_ARG;
  int blocks = ceil_div(n, 256);
  if (blocks > 1184) blocks = 1184;
  lincomb3_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(n, ca, a, cb, b, cc, c, out);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// out[0] = sum_i log(s0_i + a*p1_i + a^2*p2_i + 1e-15) at a = table[min(*kmax, len-1)] over this rank's entries
// (row-sharded problems: the partial sums are all-reduced and handed to ipm_ls_armijo_f64 as L_direct).
__global__ void __launch_bounds__(1024, 1)
ls_logsum_kernel(int nc, const double* __restrict__ s0, const double* __restrict__ p1,
                 const double* __restrict__ p2, const double* __restrict__ table, int len,
                 const int* __restrict__ kmax_ptr, double* __restrict__ out) {
  __shared__ double red[32];
  int k = *kmax_ptr;
  if (k > len - 1) k = len - 1;
  const double a = table[k];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nc; i += blockDim.x)
    acc += log(trial_slack(s0[i], p1[i], p2 ? p2[i] : 0.0, a) + LOG_GUARD);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = acc;
}

extern "C" int ipm_ls_logsum_f64(int nc, const double* s0, const double* p1, const double* p2, const double* table,
                                 int len, const int* kmax, double* out, void* stream) {
  if (nc < 0 || !table || len < 2 || !kmax || !out || (nc > 0 && (!s0 || !p1))) return IPM_ERR_ARG;
  ls_logsum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(nc, s0, p1, p2, table, len, kmax, out);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// out[0] = table[min(*kmax, len - 1)]  (step chosen by the feasibility back-off, kept on the device)
__global__ void table_lookup_kernel(const double* __restrict__ table, int len, const int* __restrict__ kmax,
                                    double* __restrict__ out) {
  int k = *kmax;
  out[0] = table[k < len - 1 ? k : len - 1];
}

extern "C" int ipm_table_lookup_f64(const double* table, int len, const int* kmax, double* out, void* stream) {
  if (!table || len < 2 || !kmax || !out) return IPM_ERR_ARG;
  table_lookup_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(table, len, kmax, out);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// element-wise helpers: op 0: out = alpha*a*b   op 1: out = alpha*a/b   op 2: out = alpha/a
__global__ void vec_op_kernel(int op, int n, const double* __restrict__ a, const double* __restrict__ b,
                              double* __restrict__ out, double alpha) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double v;
    if (op == 0) v = alpha * (a[i] * b[i]);
    else if (op == 1) v = alpha * (a[i] / b[i]);
    else v = alpha / a[i];
    out[i] = v;
  }
}

extern "C" int ipm_vec_op_f64(int op, int n, const double* a, const double* b, double* out, double alpha,
                              void* stream) {
  if (op < 0 || op > 2 || n <= 0 || !a || !out || (op < 2 && !b)) return IPM_ERR_ARG;
  int blocks = ceil_div(n, 256);
  if (blocks > 1184) blocks = 1184;
  vec_op_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(op, n, a, b, out, alpha);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}
