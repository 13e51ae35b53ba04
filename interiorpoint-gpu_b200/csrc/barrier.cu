// Barrier evaluation for the linear-inequality family (LP / QP / phase-I): slacks, reciprocals, SYRK
// weights, diagonal terms and the scalar reductions, in one fused pass over the slack vector; gradient
// assembly; Hessian finishing (diagonal + phase-I border).  Reference semantics: FunctionManager.py
// :118-149 (slacks), :208-230 (objective), :232-265 / :741-781 (gradient), :267-326 / :783-827 (Hessian),
// :429-449, :509-611 (phase-I).  Guards: log / reciprocal use s + 1e-15; the main-phase bound diagonal uses
// the raw slack (FunctionManager.py:320-322).
#include "common.cuh"

using namespace ipm;

constexpr double LOG_GUARD = 1e-15;
constexpr int EV_THREADS = 256;

// Fixed-order grid reduction: every block writes its partials, the last block to finish adds them in block
// order (deterministic run to run).  part: [nblocks][5] doubles followed by one unsigned counter.
struct EvalRed {
  double sumlog, minslack, suminv, suminv2, nneg;
};

__global__ void __launch_bounds__(EV_THREADS)
lin_barrier_eval_kernel(int m, int n, const double* __restrict__ Cx, const double* __restrict__ d,
                        const double* __restrict__ x, const double* __restrict__ ub, const double* __restrict__ lb,
                        const double* __restrict__ s_ptr, int phase1, double hd_guard, double* __restrict__ slacks,
                        double* __restrict__ inv, double* __restrict__ w, double* __restrict__ hdiag,
                        double* __restrict__ red_out, double* __restrict__ part, unsigned* __restrict__ counter) {
  __shared__ double red[32];
  __shared__ bool is_last;
  const double s = s_ptr ? *s_ptr : 0.0;
  const int ub_off = m, lb_off = m + (ub ? n : 0);
  double sumlog = 0.0, mins = INFINITY, suminv = 0.0, suminv2 = 0.0, nneg = 0.0;
  const int total = m + n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < m) {
      const double sl = phase1 ? (s + d[i]) - Cx[i] : d[i] - Cx[i];
      const double iv = 1.0 / (sl + LOG_GUARD);
      slacks[i] = sl;
      inv[i] = iv;
      w[i] = iv * iv;
      sumlog += log(sl + LOG_GUARD);
      mins = fmin(mins, sl);
      suminv += iv;
      suminv2 += iv * iv;
      nneg += sl < 0.0 ? 1.0 : 0.0;
    } else {
      const int j = i - m;
      double hd = 0.0;
      // lower bound first, then upper (FunctionManager.py:319-322, 582-587)
      if (lb) {
        const double sl = phase1 ? (s + x[j]) - lb[j] : x[j] - lb[j];
        const double iv = 1.0 / (sl + LOG_GUARD);
        slacks[lb_off + j] = sl;
        inv[lb_off + j] = iv;
        hd += (hd_guard == 1e-15) ? iv * iv : 1.0 / ((sl + hd_guard) * (sl + hd_guard));
        sumlog += log(sl + LOG_GUARD);
        mins = fmin(mins, sl);
        suminv += iv;
        suminv2 += iv * iv;
        nneg += sl < 0.0 ? 1.0 : 0.0;
      }
      if (ub) {
        const double sl = phase1 ? (s + ub[j]) - x[j] : ub[j] - x[j];
        const double iv = 1.0 / (sl + LOG_GUARD);
        slacks[ub_off + j] = sl;
        inv[ub_off + j] = iv;
        hd += (hd_guard == 1e-15) ? iv * iv : 1.0 / ((sl + hd_guard) * (sl + hd_guard));
        sumlog += log(sl + LOG_GUARD);
        mins = fmin(mins, sl);
        suminv += iv;
        suminv2 += iv * iv;
        nneg += sl < 0.0 ? 1.0 : 0.0;
      }
      hdiag[j] = hd;
    }
  }
  sumlog = block_sum(sumlog, red);
  suminv = block_sum(suminv, red);
  suminv2 = block_sum(suminv2, red);
  nneg = block_sum(nneg, red);
  mins = block_min(mins, red);
  if (threadIdx.x == 0) {
    double* p = part + 5 * blockIdx.x;
    p[0] = sumlog; p[1] = mins; p[2] = suminv; p[3] = suminv2; p[4] = nneg;
    __threadfence();
    const unsigned done = atomicAdd(counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double a0 = 0.0, a1 = INFINITY, a2 = 0.0, a3 = 0.0, a4 = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) {
      const volatile double* p = part + 5 * b;
      a0 += p[0]; a1 = fmin(a1, p[1]); a2 += p[2]; a3 += p[3]; a4 += p[4];
    }
    red_out[0] = a0; red_out[1] = a1; red_out[2] = a2; red_out[3] = a3; red_out[4] = a4;
    *counter = 0u;  // self-reset for the next call
  }
}

extern "C" long long ipm_lin_barrier_ws_doubles(void) { return 5 * 296 + 2; }

// slack layout: [m inequality rows | n upper-bound rows (if ub) | n lower-bound rows (if lb)]
// hdiag_guard g: diagonal terms 1/(s+g)^2 -- g = 1e-15 for phase-I and the diagonal-Hessian LP (FunctionManager.py:283-292,
// 561-587) instead of 1/s^2 (dense main phase, FunctionManager.py:320-322)
extern "C" int ipm_lin_barrier_eval_f64(int m, int n, const double* Cx, const double* d, const double* x,
                                        const double* ub, const double* lb, const double* s_ptr, int phase1,
                                        double hdiag_guard, double* slacks, double* inv, double* w, double* hdiag, double* red_out,
                                        double* ws, void* stream) {
  if (m < 0 || n <= 0 || !x || !slacks || !inv || !hdiag || !red_out || !ws) return IPM_ERR_ARG;
  if (m > 0 && (!Cx || !d || !w)) return IPM_ERR_ARG;
  int blocks = ceil_div(m + n, EV_THREADS);
  if (blocks > 296) blocks = 296;
  unsigned* counter = reinterpret_cast<unsigned*>(ws + 5 * 296);
  lin_barrier_eval_kernel<<<blocks, EV_THREADS, 0, (cudaStream_t)stream>>>(m, n, Cx, d, x, ub, lb, s_ptr, phase1,
                                                                          hdiag_guard, slacks, inv, w, hdiag, red_out, ws,
                                                                          counter);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// ------------------------------------------------------------------------------------------------
// Gradient assembly.
//   main   : g[j] = ((t*lin[j] - inv_lb[j]) + inv_ub[j]) + CtInv[j]          lin = c (LP) or Px+q (QP)
//   phase-I: g[j] = (CtInv[j] - inv_lb[j]) + inv_ub[j],  g[n] = t - sum(inv)
//            hxs[j] = ((-CtW[j]) + inv_lb[j]^2) - inv_ub[j]^2                 (Hessian border)
// ------------------------------------------------------------------------------------------------
__global__ void lin_grad_kernel(int n, double t, const double* __restrict__ lin, const double* __restrict__ CtInv,
                                const double* __restrict__ inv_ub, const double* __restrict__ inv_lb, int phase1,
                                const double* __restrict__ suminv, const double* __restrict__ CtW,
                                double* __restrict__ g, double* __restrict__ hxs) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) {
    if (!phase1) {
      double v = lin ? t * lin[j] : 0.0;
      if (inv_lb) v -= inv_lb[j];
      if (inv_ub) v += inv_ub[j];
      if (CtInv) v += CtInv[j];
      g[j] = v;
    } else {
      double v = CtInv ? CtInv[j] : 0.0;
      if (inv_lb) v -= inv_lb[j];
      if (inv_ub) v += inv_ub[j];
      g[j] = v;
      double h = CtW ? -CtW[j] : 0.0;
      if (inv_lb) h += inv_lb[j] * inv_lb[j];
      if (inv_ub) h -= inv_ub[j] * inv_ub[j];
      hxs[j] = h;
    }
  } else if (j == n && phase1) {
    g[n] = t - *suminv;
  }
}

extern "C" int ipm_lin_grad_f64(int n, double t, const double* lin, const double* CtInv, const double* inv_ub,
                                const double* inv_lb, int phase1, const double* suminv, const double* CtW, double* g,
                                double* hxs, void* stream) {
  if (n <= 0 || !g || (phase1 && (!suminv || !hxs))) return IPM_ERR_ARG;
  lin_grad_kernel<<<ceil_div(n + 1, 256), 256, 0, (cudaStream_t)stream>>>(n, t, lin, CtInv, inv_ub, inv_lb, phase1,
                                                                         suminv, CtW, g, hxs);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// ------------------------------------------------------------------------------------------------
// Hessian helpers (upper triangle, row-major).
// ------------------------------------------------------------------------------------------------
__global__ void hess_finish_kernel(double* __restrict__ H, long long ld, int n, const double* __restrict__ hdiag,
                                   const double* __restrict__ border, const double* __restrict__ hss,
                                   double shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double dg = H[(long long)i * ld + i];
    if (hdiag) dg += hdiag[i];
    H[(long long)i * ld + i] = dg + shift;
    if (border) H[(long long)i * ld + n] = border[i];
  } else if (i == n && border) {
    H[(long long)n * ld + n] = *hss + shift;
  }
}

// H[i][i] += hdiag[i] (+ shift); optional phase-I border: H[i][n] = border[i], H[n][n] = *hss (+ shift)
extern "C" int ipm_hess_finish_f64(double* H, int ld, int n, const double* hdiag, const double* border,
                                   const double* hss, double shift, void* stream) {
  if (!H || n <= 0 || ld < n + (border ? 1 : 0) || (border && !hss)) return IPM_ERR_ARG;
  hess_finish_kernel<<<ceil_div(n + 1, 256), 256, 0, (cudaStream_t)stream>>>(H, ld, n, hdiag, border, hss, shift);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

__global__ void scale_copy_upper_kernel(double* __restrict__ H, long long ldh, const double* __restrict__ P,
                                        long long ldp, int n, double t) {
  const int i = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    if (j >= i) H[(long long)i * ldh + j] = P ? t * P[(long long)i * ldp + j] : 0.0;
}

// H(upper) = t * P(upper)   (P == NULL: zero fill)
extern "C" int ipm_scale_copy_upper_f64(double* H, int ldh, const double* P, int ldp, int n, double t,
                                        void* stream) {
  if (!H || n <= 0 || ldh < n || (P && ldp < n)) return IPM_ERR_ARG;
  dim3 grid(ceil_div(n, 1024) < 8 ? ceil_div(n, 1024) : 8, n);
  scale_copy_upper_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(H, ldh, P, ldp, n, t);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}
