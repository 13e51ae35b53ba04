// Dense SPD factorisation and triangular solves (FP64), upper / row-major:   H = U^T U.
//
// Storing the UPPER factor row-major makes every bulk update of the factorisation the same "TN"
// contraction as the Hessian itself (k = row index of the panel, see gemm_tn_core.cuh):
//     trailing update   A22 -= U12^T U12        -> ipm_gemm_tn_f64(A = B = U12, alpha = -1, beta = 1, upper)
//     TRSM update       B2  -= U12^T Y1         -> ipm_gemm_tn_f64(A = U12, B = Y1)
// so the DMMA/TMA core does the n^3/3 flops; the per-panel pieces below (128x128 diagonal factor, 128-row
// triangular panel solve) are the serial O(n^2 NB) remainder.  Two-level blocking: the trailing update uses
// K = 256 (two 128-row panels) so that each output tile is read/written half as often.
#include "common.cuh"

using namespace ipm;

extern "C" int ipm_gemm_tn_f64(const double* A, int lda, const double* B, int ldb, const double* w, double alpha,
                               double beta, double* D, int ldd, int M, int N, int K, int upper, void* stream);

constexpr int NB = 128;    // panel height == GEMM tile
constexpr int NBO = 256;   // outer block (K of the trailing update)

// ------------------------------------------------------------------------------------------------
// Diagonal block: right-looking Cholesky of an nb x nb (nb <= 128) upper block, REGISTER resident.
// 16 warps; thread (warp w, lane t) owns rows w + 16a (a < 8) and columns t + 32b (b < 4).  Per column j the
// owning warp scales the pivot row (rsqrt, no division on the chain) and publishes it through a
// double-buffered 128-entry shared vector -- one barrier per column; everyone then applies the rank-1 update
// to its registers.  info (1-based global index of the first non-positive pivot) is written once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) potf2_kernel(double* __restrict__ A, long long ld, int nb, int k0,
                                                       int* __restrict__ info) {
  __shared__ double urow[2][NB];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double own[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int r = warp + 16 * a, c = lane + 32 * b;
      own[a][b] = (r < nb && c < nb && c >= r) ? A[(long long)r * ld + c] : 0.0;
    }
  // column j = 16*ja + jw: row j lives in own[ja][*] of warp jw, its pivot in own[ja][ja >> 1] of lane j & 31.
  // ja is unrolled so every register-array index is a compile-time constant (no local-memory spill).
#pragma unroll
  for (int ja = 0; ja < 8; ++ja) {
    for (int jw = 0; jw < 16; ++jw) {
      const int j = 16 * ja + jw;
      if (j >= nb) break;  // uniform
      double* ur = urow[j & 1];
      if (warp == jw) {
        const double d = __shfl_sync(0xffffffffu, own[ja][ja >> 1], j & 31);
        if (lane == 0 && !(d > 0.0)) atomicCAS(info, 0, k0 + j + 1);
        const double dinv = rsqrt(d);
        const double ujj = d * dinv;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int c = lane + 32 * b;
          const double v = (c == j) ? ujj : (c > j ? own[ja][b] * dinv : 0.0);
          ur[c] = v;
          if (c >= j) own[ja][b] = v;
        }
      }
      __syncthreads();
      double uc[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) uc[b] = ur[lane + 32 * b];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int r = warp + 16 * a;
        if (a >= ja && r > j && r < nb) {  // warp-uniform; rows above the pivot are final
          const double ujr = ur[r];
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (lane + 32 * b >= r) own[a][b] = fma(-ujr, uc[b], own[a][b]);
        }
      }
      // urow is double buffered: the next column writes the other half, so one barrier per column suffices
    }
  }
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int r = warp + 16 * a, c = lane + 32 * b;
      if (r < nb && c < nb && c >= r) A[(long long)r * ld + c] = own[a][b];
    }
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
  const bool vecU = (nb == NB) && !(ldu & 1) && !(((uintptr_t)U11) & 15);
  const bool vecP = (nb == NB) && (ncl == TP_COLS) && !(ldp & 1) && !(((uintptr_t)(P + col0)) & 15);
  if (vecU) {
#pragma unroll 8
    for (int q = 0; q < 32; ++q) {
      const int idx = tid + 256 * q;  // double2 index
      const int r = idx >> 6, cc = (idx & 63) * 2;
      double2 v = make_double2(0.0, 0.0);
      if (cc + 1 >= r) v = *reinterpret_cast<const double2*>(U11 + (long long)r * ldu + cc);
      Us[r * NB + cc] = cc >= r ? v.x : 0.0;
      Us[r * NB + cc + 1] = v.y;
    }
  } else {
    for (int idx = tid; idx < NB * NB; idx += 256) {
      const int r = idx >> 7, cc = idx & 127;
      Us[idx] = (r < nb && cc >= r && cc < nb) ? U11[(long long)r * ldu + cc] : 0.0;
    }
  }
  if (vecP) {
#pragma unroll 8
    for (int q = 0; q < 16; ++q) {
      const int idx = tid + 256 * q;  // double2 index: row = idx / 32, col2 = idx % 32
      const int r = idx >> 5, cc = (idx & 31) * 2;
      const double2 v = *reinterpret_cast<const double2*>(P + (long long)r * ldp + col0 + cc);
      Ps[r * TP_COLS + cc] = v.x;
      Ps[r * TP_COLS + cc + 1] = v.y;
    }
  } else {
    for (int idx = tid; idx < NB * TP_COLS; idx += 256) {
      const int r = idx >> 6, cc = idx & 63;
      Ps[idx] = (r < nb && cc < ncl) ? P[(long long)r * ldp + col0 + cc] : 0.0;
    }
  }
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += 32) {
    const int bl = min(32, nb - b0);
    if (tr == 0) {
      double x[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        if (r < bl) {
          // four partial sums shorten the dependent FMA chain
          double v0 = Ps[(b0 + r) * TP_COLS + c], v1 = 0.0, v2 = 0.0, v3 = 0.0;
#pragma unroll
          for (int l = 0; l < r; ++l) {
            const double u = Us[(b0 + l) * NB + b0 + r];
            if ((l & 3) == 0) v0 = fma(-u, x[l], v0);
            else if ((l & 3) == 1) v1 = fma(-u, x[l], v1);
            else if ((l & 3) == 2) v2 = fma(-u, x[l], v2);
            else v3 = fma(-u, x[l], v3);
          }
          x[r] = ((v0 + v1) + (v2 + v3)) / Us[(b0 + r) * NB + b0 + r];
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
#pragma unroll 8
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
  if (vecP) {
#pragma unroll 8
    for (int q = 0; q < 16; ++q) {
      const int idx = tid + 256 * q;
      const int r = idx >> 5, cc = (idx & 31) * 2;
      *reinterpret_cast<double2*>(P + (long long)r * ldp + col0 + cc) =
          make_double2(Ps[r * TP_COLS + cc], Ps[r * TP_COLS + cc + 1]);
    }
  } else {
    for (int idx = tid; idx < NB * TP_COLS; idx += 256) {
      const int r = idx >> 6, cc = idx & 63;
      if (r < nb && cc < ncl) P[(long long)r * ldp + col0 + cc] = Ps[idx];
    }
  }
}

static int launch_trsm_panel(const double* U11, long long ldu, int nb, double* P, long long ldp, int ncols,
                             cudaStream_t st) {
  if (ncols <= 0) return IPM_OK;
  const int smem = (NB * NB + NB * TP_COLS) * 8;
  static bool attr_set = false;
  if (!attr_set) {
    IPM_CUDA_CHECK(cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  trsm_panel_kernel<<<ceil_div(ncols, TP_COLS), 256, smem, st>>>(U11, ldu, nb, P, ldp, ncols);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// Library-owned side stream + events for the look-ahead (one set per device, created on first use).
struct SideStream {
  cudaStream_t stream;
  cudaEvent_t fork, join, chain, u2;
  bool ready;
};
static SideStream g_side[16];

static int get_side_stream(SideStream** out) {
  int dev = 0;
  IPM_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return IPM_ERR_ARG;
  SideStream* s = &g_side[dev];
  if (!s->ready) {
    // highest priority: the single-CTA panel chain must get the next free SM while the bulk trailing update of
    // the previous block still has CTAs queued on the caller's stream
    int prio_lo = 0, prio_hi = 0;
    IPM_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    IPM_CUDA_CHECK(cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, prio_hi));
    IPM_CUDA_CHECK(cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming));
    IPM_CUDA_CHECK(cudaEventCreateWithFlags(&s->join, cudaEventDisableTiming));
    IPM_CUDA_CHECK(cudaEventCreateWithFlags(&s->chain, cudaEventDisableTiming));
    IPM_CUDA_CHECK(cudaEventCreateWithFlags(&s->u2, cudaEventDisableTiming));
    s->ready = true;
  }
  *out = s;
  return IPM_OK;
}

// ------------------------------------------------------------------------------------------------
// ipm_potrf_upper_f64: in-place blocked right-looking Cholesky, H = U^T U (upper triangle of H in/out; the
// strict lower triangle is never read or written).  *info_dev = 0 on success, else the 1-based index of the
// first non-positive pivot (LAPACK dpotrf convention); it is written on the device, never synchronised here.
//
// for each outer block of 256 rows:   [ potf2(128) ; trsm of its 128-row panel ; K=128 update of the second
// 128-row band ]  [ potf2(128) ; trsm of the second panel ]  then ONE K=256 DMMA update of the trailing matrix.
// ------------------------------------------------------------------------------------------------
extern "C" int ipm_potrf_upper_f64(double* H, int ld, int n, int* info_dev, void* stream) {
  if (!H || !info_dev || n < 0 || ld < n || (ld & 1)) return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  IPM_CUDA_CHECK(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
  // Look-ahead: the serial panel chain of outer block o+1 (two potf2 + two panel solves, single-CTA latency
  // bound) runs on a side stream concurrently with the bulk trailing update of block o on the caller's stream.
  //   side : chain(0), U1(0), chain(1), [wait U2(0)] U1(1), chain(2), ...
  //   main :            [wait chain(0)] U2(0), [wait chain(1)] U2(1), ...
  // U1(o) = update of the NEXT block's 256 rows (all the next chain needs); U2(o) = rows below them.
  SideStream* ss = nullptr;
  const bool lookahead = n > 2 * NBO;
  if (lookahead) {
    int rc = get_side_stream(&ss);
    if (rc) return rc;
    IPM_CUDA_CHECK(cudaEventRecord(ss->fork, st));
    IPM_CUDA_CHECK(cudaStreamWaitEvent(ss->stream, ss->fork, 0));
  }
  cudaStream_t cs = lookahead ? ss->stream : st;  // stream of the panel chain
  bool have_u2 = false;
  for (int o0 = 0; o0 < n; o0 += NBO) {
    const int oend = o0 + NBO < n ? o0 + NBO : n;
    for (int k0 = o0; k0 < oend; k0 += NB) {
      const int nb = n - k0 < NB ? n - k0 : NB;
      double* Akk = H + (long long)k0 * ld + k0;
      potf2_kernel<<<1, 512, 0, cs>>>(Akk, ld, nb, k0, info_dev);
      IPM_LAUNCH_CHECK();
      const int rest = n - k0 - nb;
      if (rest <= 0) continue;
      double* A12 = Akk + nb;  // nb x rest row panel
      int rc = launch_trsm_panel(Akk, ld, nb, A12, ld, rest, cs);
      if (rc) return rc;
      const int band = oend - (k0 + nb);  // rows of this outer block still to be factored
      if (band > 0) {
        // A[k0+nb : oend, k0+nb : n] -= U12[:, :band]^T U12      (band x rest, K = nb; store col >= row only)
        double* Aband = H + (long long)(k0 + nb) * ld + (k0 + nb);
        rc = ipm_gemm_tn_f64(A12, ld, A12, ld, nullptr, -1.0, 1.0, Aband, ld, band, rest, nb, 2, (void*)cs);
        if (rc) return rc;
      }
    }
    const int rest = n - oend;
    if (rest <= 0) break;
    const int K = oend - o0;
    const double* Uo = H + (long long)o0 * ld + oend;  // K x rest panel rows of this outer block
    double* A22 = H + (long long)oend * ld + oend;
    if (!lookahead) {
      int rc = ipm_gemm_tn_f64(Uo, ld, Uo, ld, nullptr, -1.0, 1.0, A22, ld, rest, rest, K, 1, stream);
      if (rc) return rc;
      continue;
    }
    IPM_CUDA_CHECK(cudaEventRecord(ss->chain, cs));
    const int band2 = rest < NBO ? rest : NBO;
    if (have_u2) IPM_CUDA_CHECK(cudaStreamWaitEvent(cs, ss->u2, 0));  // U1(o) rewrites rows that U2(o-1) wrote
    int rc = ipm_gemm_tn_f64(Uo, ld, Uo, ld, nullptr, -1.0, 1.0, A22, ld, band2, rest, K, 2, (void*)cs);  // U1(o)
    if (rc) return rc;
    const int rest2 = rest - band2;
    if (rest2 > 0) {
      IPM_CUDA_CHECK(cudaStreamWaitEvent(st, ss->chain, 0));
      rc = ipm_gemm_tn_f64(Uo + band2, ld, Uo + band2, ld, nullptr, -1.0, 1.0,
                           A22 + (long long)band2 * ld + band2, ld, rest2, rest2, K, 1, stream);  // U2(o)
      if (rc) return rc;
      IPM_CUDA_CHECK(cudaEventRecord(ss->u2, st));
      have_u2 = true;
    }
  }
  if (lookahead) {
    IPM_CUDA_CHECK(cudaEventRecord(ss->join, cs));
    IPM_CUDA_CHECK(cudaStreamWaitEvent(st, ss->join, 0));
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
  for (int o0 = 0; o0 < n && p > 0; o0 += NBO) {
    const int oend = o0 + NBO < n ? o0 + NBO : n;
    for (int k0 = o0; k0 < oend; k0 += NB) {
      const int nb = n - k0 < NB ? n - k0 : NB;
      const double* Ukk = U + (long long)k0 * ldu + k0;
      double* Bk = B + (long long)k0 * ldb;
      int rc = launch_trsm_panel(Ukk, ldu, nb, Bk, ldb, p, st);
      if (rc) return rc;
      const int band = oend - (k0 + nb);
      if (band > 0) {  // rows of the same outer block: B[k0+nb : oend] -= U[k0:k0+nb, k0+nb:oend]^T Y_k
        rc = ipm_gemm_tn_f64(Ukk + nb, ldu, Bk, ldb, nullptr, -1.0, 1.0, B + (long long)(k0 + nb) * ldb, ldb, band, p,
                             nb, 0, stream);
        if (rc) return rc;
      }
    }
    const int rest = n - oend;
    if (rest > 0) {  // all rows below the outer block, K = 256
      int rc = ipm_gemm_tn_f64(U + (long long)o0 * ldu + oend, ldu, B + (long long)o0 * ldb, ldb, nullptr, -1.0, 1.0,
                               B + (long long)oend * ldb, ldb, rest, p, oend - o0, 0, stream);
      if (rc) return rc;
    }
  }
  return IPM_OK;
}
