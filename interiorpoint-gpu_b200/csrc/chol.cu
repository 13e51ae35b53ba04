// Dense SPD factorisation and triangular solves (FP64), upper / row-major:   H = U^T U.
//
// Storing the UPPER factor row-major makes every bulk update of the factorisation the same "TN"
// contraction as the Hessian itself (k = row index of the panel, see gemm_tn_core.cuh):
//     trailing update   A22 -= U12^T U12        -> ipm_gemm_tn_f64(A = B = U12, alpha = -1, beta = 1, upper)
//     TRSM update       B2  -= U12^T Y1         -> ipm_gemm_tn_f64(A = U12, B = Y1)
// so the DMMA/TMA core does the n^3/3 flops; the per-panel pieces below (128x128 diagonal factor, 128-row
// triangular panel solve) are the serial O(n^2 NB) remainder.  Two-level blocking: the trailing update uses
// K = 384 (three 128-row panels) so that each output tile is read/written a third as often.
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "common.cuh"
#include "gemm_tn_core.cuh"
#include "tensormap.cuh"

using namespace ipm;

extern "C" int ipm_gemm_tn_f64(const double* A, int lda, const double* B, int ldb, const double* w, double alpha,
                               double beta, double* D, int ldd, int M, int N, int K, int upper, void* stream);

constexpr int NB = 128;    // panel height == GEMM tile
constexpr int NBO = 384;   // outer block (K of the trailing update): three panels; measured 256 / 384 / 512 -> 9.05 / 8.87 / 8.94 ms at n = 8192

// ------------------------------------------------------------------------------------------------
// Diagonal block: right-looking Cholesky of an nb x nb (nb <= 128) upper block held in shared memory, blocked in
// 32-column sub-panels so that the 128-step dependent chain never crosses a CTA barrier:
//   (a) warp 0 factors the 32x32 diagonal sub-block in registers (lane = column; pivot and row entries travel by
//       shuffle; rsqrt, no division on the chain),
//   (b) the 32 x W row panel to its right is solved by one thread per column (right-looking substitution, the
//       chain per row is DMUL -> DFMA),
//   (c) all warps apply the rank-32 update to the W x W trailing block as 8x8 DMMA tiles (upper tiles only).
// info (1-based global index of the first non-positive pivot) is written once.
// ------------------------------------------------------------------------------------------------
#ifdef IPM_PHASE_TIMING
__device__ long long g_phase_t[64];
#define IPM_PHASE_MARK(i)                                  \
  do {                                                     \
    if (threadIdx.x == 0 && blockIdx.x == 0) g_phase_t[i] = clock64(); \
  } while (0)
#else
#define IPM_PHASE_MARK(i) \
  do {                    \
  } while (0)
#endif

constexpr int PF_LD = NB + 4;  // 132: rows 16-byte aligned; (4k + m) mod 16 distinct -> conflict-free DMMA fragments
constexpr int PF_THREADS = 256;  // 8 warps: the register-resident 32x32 pivot block needs > 128 registers/thread
constexpr int PF_SMEM = NB * PF_LD * 8;

// Branch-free 1/sqrt(x): hardware seed (rsqrt.approx.f64, ~2^-22) + one third-order correction
// y = y0 (1 + e/2 + 3 e^2 / 8), e = 1 - x y0^2  -> relative error ~2^-64 before rounding.  The CUDA library
// rsqrt() has a call to a slow path for special operands, which splits the unrolled pivot loop into basic blocks
// and keeps ptxas from overlapping one column's shuffles with the next column's pivot chain.
__device__ __forceinline__ double rsqrt_nobranch(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double e = fma(-x, y0 * y0, 1.0);
  const double c = fma(e, 0.375, 0.5);
  return fma(c, y0 * e, y0);
}

// (c) of potf2_kernel: rank-32 update of the W x W trailing block, T -= P^T P with P the 32 x W row panel, as
// 8x8 DMMA tiles (upper tiles only, round-robin over the warps).  Fragment (k = lane & 3, m = lane >> 2) of
// k-group kk and 8-column block x:  S[(base + 4 kk + k) * PF_LD + t0 + 8 x + m].
__device__ __forceinline__ void potf2_trailing_update(double* __restrict__ S, int base, int W, int warp, int lane) {
  const int t0 = base + 32, wt = W >> 3;
  const int k4 = lane & 3, m8 = lane >> 2;
  const double* frag = S + (base + k4) * PF_LD + t0 + m8;
  int ri = 0, first = 0;  // `first` = linear index of tile (ri, ri)
  for (int t = warp; t < wt * (wt + 1) / 2; t += PF_THREADS / 32) {
    while (t >= first + (wt - ri)) first += wt - ri, ++ri;
    const int ci = ri + (t - first);
    double2* cp = reinterpret_cast<double2*>(S + (t0 + 8 * ri + m8) * PF_LD + t0 + 8 * ci + 2 * k4);
    double2 c = *cp;
    double a[8], b[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      a[kk] = -frag[(4 * kk) * PF_LD + 8 * ri];
      b[kk] = frag[(4 * kk) * PF_LD + 8 * ci];
    }
#pragma unroll
    for (int kk = 0; kk < 8; ++kk)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c.x), "+d"(c.y)
                   : "d"(a[kk]), "d"(b[kk]));
    *cp = c;
  }
}

// Loads of tile data: CG = true reads through L2 only (ld.global.cg) -- for tiles another CTA of the SAME kernel has
// just published (the persistent tile-DAG factorisation); the stand-alone kernels use plain loads.
template <bool CG>
__device__ __forceinline__ double2 tile_ld2(const double* p) {
  return CG ? __ldcg(reinterpret_cast<const double2*>(p)) : *reinterpret_cast<const double2*>(p);
}
template <bool CG>
__device__ __forceinline__ double tile_ld1(const double* p) {
  return CG ? __ldcg(p) : *p;
}

// Body of potf2_kernel for one CTA of PF_THREADS threads: S = NB x PF_LD doubles of shared memory, rs = 32 doubles.
template <bool CG>
__device__ __forceinline__ void potf2_tile(double* __restrict__ S, double* __restrict__ rs, double* __restrict__ A,
                                           long long ld, int nb, int k0, int* __restrict__ info) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool vec = !(ld & 1) && !(((uintptr_t)A) & 15);
  if (vec) {
#pragma unroll 16
    for (int q = 0; q < NB * NB / 2 / PF_THREADS; ++q) {
      const int idx = tid + PF_THREADS * q;  // pair index: row = idx / 64, col = 2 * (idx % 64)
      const int r = idx >> 6, c = (idx & 63) * 2;
      double2 v = make_double2(r == c ? 1.0 : 0.0, r == c + 1 ? 1.0 : 0.0);
      if (r < nb && c + 1 >= r) {
        if (c + 1 < nb) {
          v = tile_ld2<CG>(A + (long long)r * ld + c);
        } else if (c < nb) {
          v.x = tile_ld1<CG>(A + (long long)r * ld + c);
        }
      }
      S[r * PF_LD + c] = v.x;
      S[r * PF_LD + c + 1] = v.y;
    }
  } else {
    for (int idx = tid; idx < NB * NB; idx += PF_THREADS) {
      const int r = idx >> 7, c = idx & 127;
      S[r * PF_LD + c] = (r < nb && c < nb && c >= r) ? tile_ld1<CG>(A + (long long)r * ld + c) : (r == c ? 1.0 : 0.0);
    }
  }
  IPM_PHASE_MARK(0);
  __syncthreads();
  IPM_PHASE_MARK(1);
  for (int base = 0; base < NB; base += 32) {
    if (base >= nb) break;  // uniform: the rest is identity padding
    IPM_PHASE_MARK(2 + (base >> 5) * 4);
    if (warp == 0) {
      // (a) 32x32 diagonal sub-block, lane = column
      double d[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) d[i] = S[(base + i) * PF_LD + base + lane];
      int bad = 0;  // 1-based column of the first non-positive pivot of this sub-block (select, no branch)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const double piv = __shfl_sync(0xffffffffu, d[j], j);
        bad = (bad == 0 && !(piv > 0.0)) ? j + 1 : bad;
        const double r = rsqrt_nobranch(piv);
        const double u = d[j] * r;  // lane c: U[j][c] (c == j: piv * r = sqrt(piv))
        d[j] = u;
        if (lane == j) rs[j] = r;
#pragma unroll
        for (int i = j + 1; i < 32; ++i) {
          const double ui = __shfl_sync(0xffffffffu, u, i);
          d[i] = fma(-ui, u, d[i]);  // meaningful for lanes c >= i
        }
      }
      if (bad && lane == 0) atomicCAS(info, 0, k0 + base + bad);
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (lane >= i) S[(base + i) * PF_LD + base + lane] = d[i];
    }
    __syncthreads();
    IPM_PHASE_MARK(3 + (base >> 5) * 4);
    const int W = NB - base - 32;  // 96, 64, 32, 0
    if (W == 0) break;
    if (tid < W) {
      // (b) row panel: column c of the 32 x W block right of the diagonal sub-block
      const int c = base + 32 + tid;
      double v[32];
#pragma unroll
      for (int l = 0; l < 32; ++l) v[l] = S[(base + l) * PF_LD + c];
#pragma unroll
      for (int l = 0; l < 32; ++l) {
        const double x = v[l] * rs[l];
        v[l] = x;
#pragma unroll
        for (int r = l + 1; r < 32; ++r) v[r] = fma(-S[(base + l) * PF_LD + base + r], x, v[r]);
      }
#pragma unroll
      for (int l = 0; l < 32; ++l) S[(base + l) * PF_LD + c] = v[l];
    }
    __syncthreads();
    IPM_PHASE_MARK(4 + (base >> 5) * 4);
    // (c) trailing W x W block
    potf2_trailing_update(S, base, W, warp, lane);
    __syncthreads();
    IPM_PHASE_MARK(5 + (base >> 5) * 4);
  }
  __syncthreads();
  IPM_PHASE_MARK(20);
  if (vec) {
#pragma unroll 16
    for (int q = 0; q < NB * NB / 2 / PF_THREADS; ++q) {
      const int idx = tid + PF_THREADS * q;
      const int r = idx >> 6, c = (idx & 63) * 2;
      if (r < nb && c + 1 >= r) {
        double* p = A + (long long)r * ld + c;
        if (c >= r && c + 1 < nb) {
          *reinterpret_cast<double2*>(p) = make_double2(S[r * PF_LD + c], S[r * PF_LD + c + 1]);
        } else {
          if (c >= r && c < nb) p[0] = S[r * PF_LD + c];
          if (c + 1 < nb) p[1] = S[r * PF_LD + c + 1];
        }
      }
    }
  } else {
    for (int idx = tid; idx < NB * NB; idx += PF_THREADS) {
      const int r = idx >> 7, c = idx & 127;
      if (r < nb && c < nb && c >= r) A[(long long)r * ld + c] = S[r * PF_LD + c];
    }
  }
  IPM_PHASE_MARK(21);
}


__global__ void __launch_bounds__(PF_THREADS, 1) potf2_kernel(double* __restrict__ A, long long ld, int nb, int k0,
                                                              int* __restrict__ info) {
  extern __shared__ double S[];  // NB x PF_LD; identity beyond nb, garbage-tolerant strict lower triangle
  __shared__ double rs[32];      // rsqrt of the current sub-block's pivots
  pdl_wait();
  potf2_tile<false>(S, rs, A, ld, nb, k0, info);
}

static int launch_potf2(double* Akk, long long ld, int nb, int k0, int* info, cudaStream_t st) {
  static bool attr_set[kMaxDevices];
  IPM_CUDA_CHECK(ensure_dynamic_smem(potf2_kernel, PF_SMEM, attr_set));
  IPM_CUDA_CHECK(launch_pdl(potf2_kernel, dim3(1), dim3(PF_THREADS), PF_SMEM, st, Akk, ld, nb, k0, info));
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// ------------------------------------------------------------------------------------------------
// Panel solve  X = U11^{-T} P   (U11: nb x nb upper, P: nb x ncols row-major, in place).
// CTA = 64 columns; U11 and the CTA's panel slice live in shared memory (192 KiB).  32-row blocks: two warps run
// the substitution of the block right-looking with the 32 values of their column in registers (chain per row:
// DMUL by the reciprocal pivot -> DFMA), then all 256 threads update the rows below (8 rows x 1 column per
// thread, coefficients by 16-byte shared loads).
// ------------------------------------------------------------------------------------------------
constexpr int TP_COLS = 64;
constexpr int US_LD = NB + 4;        // 132 and 68 are 4 (mod 16): the (k = lane & 3, m = lane >> 2) fragment loads of
constexpr int PS_LD = TP_COLS + 4;   // the DMMA update hit 16 distinct 8-byte banks per half-warp, rows stay 16-byte aligned

// Pieces of the panel solve for one CTA of 256 threads (Us: NB x US_LD, Ps: NB x PS_LD, rinv: NB doubles, all shared).
template <bool CG>
__device__ __forceinline__ void trsm_load_u(double* __restrict__ Us, const double* __restrict__ U11, long long ldu,
                                            int nb) {
  const int tid = threadIdx.x;
  const bool vecU = (nb == NB) && !(ldu & 1) && !(((uintptr_t)U11) & 15);
  if (vecU) {
#pragma unroll 16
    for (int q = 0; q < 32; ++q) {
      const int idx = tid + 256 * q;  // double2 index
      const int r = idx >> 6, cc = (idx & 63) * 2;
      double2 v = make_double2(0.0, 0.0);
      if (cc + 1 >= r) v = tile_ld2<CG>(U11 + (long long)r * ldu + cc);
      Us[r * US_LD + cc] = cc >= r ? v.x : 0.0;
      Us[r * US_LD + cc + 1] = v.y;
    }
  } else {
    for (int idx = tid; idx < NB * NB; idx += 256) {
      const int r = idx >> 7, cc = idx & 127;
      Us[r * US_LD + cc] = (r < nb && cc >= r && cc < nb) ? tile_ld1<CG>(U11 + (long long)r * ldu + cc) : 0.0;
    }
  }
}

// vecP: nb == NB, ncl == TP_COLS, even ldp, 16-byte aligned P + col0
template <bool CG>
__device__ __forceinline__ void trsm_load_p(double* __restrict__ Ps, const double* __restrict__ P, long long ldp,
                                            int nb, int col0, int ncl, bool vecP) {
  const int tid = threadIdx.x;
  if (vecP) {
#pragma unroll 16
    for (int q = 0; q < 16; ++q) {
      const int idx = tid + 256 * q;  // double2 index: row = idx / 32, col2 = idx % 32
      const int r = idx >> 5, cc = (idx & 31) * 2;
      const double2 v = tile_ld2<CG>(P + (long long)r * ldp + col0 + cc);
      Ps[r * PS_LD + cc] = v.x;
      Ps[r * PS_LD + cc + 1] = v.y;
    }
  } else {
    for (int idx = tid; idx < NB * TP_COLS; idx += 256) {
      const int r = idx >> 6, cc = idx & 63;
      Ps[r * PS_LD + cc] = (r < nb && cc < ncl) ? tile_ld1<CG>(P + (long long)r * ldp + col0 + cc) : 0.0;
    }
  }
}

__device__ __forceinline__ void trsm_store_p(const double* __restrict__ Ps, double* __restrict__ P, long long ldp,
                                             int nb, int col0, int ncl, bool vecP) {
  const int tid = threadIdx.x;
  if (vecP) {
#pragma unroll 8
    for (int q = 0; q < 16; ++q) {
      const int idx = tid + 256 * q;
      const int r = idx >> 5, cc = (idx & 31) * 2;
      *reinterpret_cast<double2*>(P + (long long)r * ldp + col0 + cc) =
          make_double2(Ps[r * PS_LD + cc], Ps[r * PS_LD + cc + 1]);
    }
  } else {
    for (int idx = tid; idx < NB * TP_COLS; idx += 256) {
      const int r = idx >> 6, cc = idx & 63;
      if (r < nb && cc < ncl) P[(long long)r * ldp + col0 + cc] = Ps[r * PS_LD + cc];
    }
  }
}

#ifdef IPM_DAG_TIMING
namespace dag {
__device__ long long g_dag_t[256 * 16];  // per-CTA cycle counters, see the slot list at DAG_ADD
}
#define TRSM_MARK(slot, t0)                                                              \
  do {                                                                                   \
    if (threadIdx.x == 0) dag::g_dag_t[blockIdx.x * 16 + (slot)] += clock64() - (t0);    \
    (t0) = clock64();                                                                    \
  } while (0)
#else
#define TRSM_MARK(slot, t0) \
  do {                      \
  } while (0)
#endif

// Substitution of the TP_COLS columns held in Ps (in place); ends with a CTA barrier.
__device__ __forceinline__ void trsm_solve(const double* __restrict__ Us, double* __restrict__ Ps,
                                           const double* __restrict__ rinv, int nb) {
  const int tid = threadIdx.x;
  const int c = tid & (TP_COLS - 1), tr = tid >> 6;  // 4 row phases
#ifdef IPM_DAG_TIMING
  long long tm0 = clock64();
#endif
  // one copy of the (fully unrolled, ~14 KB) block substitution: unrolled over b0 -- and over the two column halves of
  // the tile-DAG kernel -- it becomes 8 copies that are each executed once per tile and run ~4x slower
#pragma unroll 1
  for (int b0 = 0; b0 < nb; b0 += 32) {
    if (tr == 0) {
      // rows beyond nb are zero rows of Us / Ps with rinv = 1: harmless
      double v[32];
#pragma unroll
      for (int l = 0; l < 32; ++l) v[l] = Ps[(b0 + l) * PS_LD + c];
#pragma unroll
      for (int l = 0; l < 32; ++l) {
        const double x = v[l] * rinv[b0 + l];
        v[l] = x;
        const double* urow = Us + (b0 + l) * US_LD + b0;
#pragma unroll
        for (int r = l + 1; r < 32; ++r) v[r] = fma(-urow[r], x, v[r]);
      }
#pragma unroll
      for (int l = 0; l < 32; ++l) Ps[(b0 + l) * PS_LD + c] = v[l];
    }
    __syncthreads();
    TRSM_MARK(10, tm0);
    // rows below the block:  Ps[rb.., :] -= U[b0..b0+32, rb..]^T X[b0..b0+32, :]  as 8x8 DMMA tiles, K = 32.
    // Warp w owns the 8 columns 8w .. 8w+7 (its eight B fragments are loaded once), and walks the row blocks four
    // at a time so that four independent accumulator chains are in flight.
    {
      const int lane = tid & 31, wp = tid >> 5, l4 = lane & 3, g8 = lane >> 2;
      const int rest0 = b0 + 32;  // rows beyond nb are zero rows / zero coefficient columns: harmless
      double bf[8];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) bf[kk] = Ps[(b0 + 4 * kk + l4) * PS_LD + 8 * wp + g8];
      for (int r0 = rest0; r0 < NB; r0 += 32) {
        double2 cacc[4];
        double af[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          cacc[i] = *reinterpret_cast<const double2*>(Ps + (r0 + 8 * i + g8) * PS_LD + 8 * wp + 2 * l4);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) af[i][kk] = -Us[(b0 + 4 * kk + l4) * US_LD + r0 + 8 * i + g8];
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(cacc[i].x), "+d"(cacc[i].y)
                         : "d"(af[i][kk]), "d"(bf[kk]));
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<double2*>(Ps + (r0 + 8 * i + g8) * PS_LD + 8 * wp + 2 * l4) = cacc[i];
      }
    }
    __syncthreads();
    TRSM_MARK(11, tm0);
  }
}

__global__ void __launch_bounds__(256, 1) trsm_panel_kernel(const double* __restrict__ U11, long long ldu, int nb,
                                                            double* __restrict__ P, long long ldp, int ncols) {
  extern __shared__ double sm[];
  double* Us = sm;                // NB x NB (ld US_LD)
  double* Ps = sm + NB * US_LD;   // NB x TP_COLS (ld PS_LD)
  __shared__ double rinv[NB];
  const int tid = threadIdx.x;
  const int col0 = blockIdx.x * TP_COLS;
  const int ncl = min(TP_COLS, ncols - col0);
  const bool vecP = (nb == NB) && (ncl == TP_COLS) && !(ldp & 1) && !(((uintptr_t)(P + col0)) & 15);
  pdl_wait();
  trsm_load_u<false>(Us, U11, ldu, nb);
  trsm_load_p<false>(Ps, P, ldp, nb, col0, ncl, vecP);
  __syncthreads();
  if (tid < NB) rinv[tid] = tid < nb ? 1.0 / Us[tid * US_LD + tid] : 1.0;
  __syncthreads();
  trsm_solve(Us, Ps, rinv, nb);
  trsm_store_p(Ps, P, ldp, nb, col0, ncl, vecP);
}

static int launch_trsm_panel(const double* U11, long long ldu, int nb, double* P, long long ldp, int ncols,
                             cudaStream_t st) {
  if (ncols <= 0) return IPM_OK;
  const int smem = (NB * US_LD + NB * PS_LD) * 8;
  static bool attr_set[kMaxDevices];
  IPM_CUDA_CHECK(ensure_dynamic_smem(trsm_panel_kernel, smem, attr_set));
  IPM_CUDA_CHECK(launch_pdl(trsm_panel_kernel, dim3(ceil_div(ncols, TP_COLS)), dim3(256), smem, st, U11, ldu, nb, P, ldp,
                            ncols));
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// ------------------------------------------------------------------------------------------------
// Persistent tile-DAG factorisation (left-looking), ONE launch for the whole matrix.  Opt-in (IPM_POTRF_DAG=1).
//
// Task (i, j), i <= j, owns the 128 x 128 tile U(i, j):
//     acc = sum_{k < i} U(k, i)^T U(k, j)       the TN contraction of gemm_tn_core.cuh over the 128 i rows above the
//                                               tile (TMA ring + DMMA, accumulators in registers for the whole K)
//     B   = A(i, j) - acc                       written back in place
//     i == j:  U(i, i) = chol(B)  (potf2_tile)      i < j:  U(i, j) = U(i, i)^{-T} B  (trsm pieces, two 64-column halves)
// and then publishes the tile.  Tile (i, j) needs tile (i - 1, j), so the tiles of a block column are published top
// down and ONE counter per block column is enough: done[j] = (epoch << 32) | (number of finished row blocks of column
// j), written with a release store (epoch-stamped, never reset).  Tasks are numbered row-major over the upper
// triangle, CTA c runs tasks c, c + G, c + 2G, ... in that order.  Every dependency of a task (the tiles above it in
// its two block columns, and the diagonal tile of its row) has a smaller number, and all G <= #SMs CTAs are resident
// (one per SM by shared memory), so the lowest-numbered unfinished task can always proceed: no deadlock.  The
// producer lanes check done[i] and done[j] right before they issue the TMA loads of row block k -- and remember the
// counts, so a task whose inputs were finished long ago polls once -- hence a task starts accumulating as soon as
// the first rows above it exist and only its LAST row block sits on the critical path (look-ahead falls out of the
// schedule instead of being arranged with streams).
//
// Compared with the stream-ordered right-looking code below, the dependent chain per 128 columns is potf2 -> one
// tile solve -> one K = 128 accumulation, with no kernel boundary in between, and every tile is written once.
// ------------------------------------------------------------------------------------------------
namespace dag {
using namespace ipm::gemm;

constexpr int KT_PER_BLOCK = NB / BK;                                  // k-tiles per 128-row block
constexpr int SCRATCH_BYTES = (NB * US_LD + NB * PS_LD) * 8;            // tile-solve scratch (>= potf2's, >= the ring)
static_assert(SCRATCH_BYTES >= STAGES * STAGE_BYTES && SCRATCH_BYTES >= PF_SMEM, "scratch aliases the TMA ring");
constexpr int SMEM = 1024 /*align slack*/ + SCRATCH_BYTES + 2 * STAGES * 8;
constexpr int MAX_T = 256;                                             // n <= 32768

#ifdef IPM_DAG_TIMING
// per CTA: [0] whole kernel, [1] contraction, [2] wait for the row's diagonal tile, [3] potf2, [4] tile solves,
// [5] publish, [6] tasks, [7] polling of the progress counters by warp 0 (inside [1]); inside [4]: [8] load U,
// [9] load P, [10] substitution (2 warps), [11] DMMA update of the rows below, [12] store, [13] subtract_acc (in [1])
// (clock64 cycles)
#define DAG_MARK(slot, t0) TRSM_MARK(slot, t0)
#define DAG_CLOCK() clock64()
#define DAG_ADD(slot, t0)                                                         \
  do {                                                                            \
    if (threadIdx.x == 0) g_dag_t[blockIdx.x * 16 + (slot)] += clock64() - (t0);   \
  } while (0)
#else
#define DAG_MARK(slot, t0) \
  do {                     \
    (void)(t0);            \
  } while (0)
#define DAG_CLOCK() 0ll
#define DAG_ADD(slot, t0) \
  do {                    \
    (void)(t0);           \
  } while (0)
#endif

// number of published row blocks of a block column (0 while the counter still carries another call's epoch)
__device__ __forceinline__ int column_progress(const unsigned long long* p, unsigned epoch) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return (unsigned)(v >> 32) == epoch ? (int)(unsigned)v : 0;
}
// Bounded (common.cuh: spin_wait).  When the watchdog fires, info becomes -1 and the wait reports the column as
// complete, so that the task -- and with it the whole grid -- finishes on garbage instead of hanging.
__device__ __forceinline__ int wait_column(const unsigned long long* p, unsigned epoch, int need, unsigned int* fault,
                                           int* info) {
  int r = 0;
  if (!spin_wait([&] { return (r = column_progress(p, epoch)) >= need; }, fault, IPM_FAULT_POTRF_DAG)) {
    atomicCAS(info, 0, -1);
    r = MAX_T;
  }
  return r;
}

// Producer cursor of one task (every warp keeps a copy; lane 0 issues column chunk `wp` of both operands).
struct Producer {
  const CUtensorMap* tm;
  Ring ring;
  const unsigned long long* done;
  unsigned epoch;
  unsigned int* fault;
  int* info;
  int ti, tj, kt, k1, ready_i, ready_j;  // ready_*: row blocks of columns ti / tj known to be published
  uint32_t it;
  __device__ __forceinline__ void begin(int i, int j) {
    ti = i, tj = j, kt = 0, k1 = i * KT_PER_BLOCK, ready_i = ready_j = 0;
  }
  __device__ __forceinline__ void issue(int wp, int lane) {
    if (kt >= k1) return;
    if (lane == 0) {
      const int need = kt / KT_PER_BLOCK + 1;
      if ((kt % KT_PER_BLOCK) == 0 && (ready_i < need || ready_j < need)) {
#ifdef IPM_DAG_TIMING
        const long long t0 = clock64();
#endif
        if (ready_i < need) ready_i = wait_column(done + ti, epoch, need, fault, info);
        if (tj == ti) ready_j = ready_i;
        if (ready_j < need) ready_j = wait_column(done + tj, epoch, need, fault, info);
        asm volatile("fence.proxy.async;" ::: "memory");  // the tiles were written through the generic proxy
#ifdef IPM_DAG_TIMING
        if (wp == 0) g_dag_t[blockIdx.x * 16 + 7] += clock64() - t0;
#endif
      }
      const uint32_t s = it % STAGES;
      if (it >= STAGES) mbar_wait(ring.empty0 + 8 * s, ((it / STAGES) - 1) & 1);
      const uint32_t full = ring.full0 + 8 * s;
      mbar_expect_tx(full, 2 * CHUNK_BYTES);
      const uint32_t dstA = ring.tiles0 + s * STAGE_BYTES + wp * CHUNK_BYTES;
      tma_load_2d(dstA, tm, ti * BM + wp * 16, kt * BK, full);
      tma_load_2d(dstA + OPERAND_BYTES, tm, tj * BN + wp * 16, kt * BK, full);
    }
    __syncwarp();
    ++it;
    ++kt;
  }
};

// H(tile) -= acc for the warp's 64 x 32 piece (rows / columns beyond n and, on diagonal tiles, col < row skipped).
__device__ __forceinline__ void subtract_acc(double* __restrict__ H, long long ld, int n, bool diag,
                                             const double (&acc)[8][4][2], int m_base, int n_base, int g8, int l4) {
  const bool interior = !(ld & 1) && !(((uintptr_t)H) & 15) && (m_base + 64 <= n) && (n_base + 32 <= n) &&
                        (!diag || n_base >= m_base + 63);
  if (interior) {
#pragma unroll
    for (int ib = 0; ib < 8; ib += 2) {
      double2 old[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int jn = 0; jn < 4; ++jn)
          old[i][jn] = __ldcg(reinterpret_cast<const double2*>(H + (long long)(m_base + (ib + i) * 8 + g8) * ld +
                                                               n_base + jn * 8 + 2 * l4));
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int jn = 0; jn < 4; ++jn)
          *reinterpret_cast<double2*>(H + (long long)(m_base + (ib + i) * 8 + g8) * ld + n_base + jn * 8 + 2 * l4) =
              make_double2(old[i][jn].x - acc[ib + i][jn][0], old[i][jn].y - acc[ib + i][jn][1]);
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m_base + i * 8 + g8;
#pragma unroll
    for (int jn = 0; jn < 4; ++jn)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = n_base + jn * 8 + 2 * l4 + e;
        if (row < n && col < n && !(diag && col < row)) {
          double* p = H + (long long)row * ld + col;
          *p = __ldcg(p) - acc[i][jn][e];
        }
      }
  }
}

__global__ void __launch_bounds__(THREADS, 1)
potrf_dag_kernel(const __grid_constant__ CUtensorMap tm, double* __restrict__ H, long long ld, int n, int T,
                 int* __restrict__ info, unsigned long long* __restrict__ done, unsigned epoch,
                 unsigned int* __restrict__ fault) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ double rinv[NB];
  __shared__ double rs[32];
  using S = Shape128x128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  // 1024-byte alignment (128B swizzle) as an OFFSET into the shared array: arithmetic on the integer value of the
  // pointer would make every scratch access a generic LD/ST instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  double* scratch = reinterpret_cast<double*>(smem);  // the TMA ring during the contraction, tile scratch after it
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SCRATCH_BYTES);
  const Ring ring{smem_u32(bars), smem_u32(bars + STAGES), smem_u32(smem)};
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(ring.full0 + 8 * s, CONSUMER_WARPS);
      mbar_init(ring.empty0 + 8 * s, CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();
  const LaneMap lm = make_lane_map<S>(warp, lane);
  Producer prod{&tm, ring, done, epoch, fault, info, 0, 0, 0, 0, 0, 0, 0u};
  uint32_t it = 0;
  const int ntasks = T * (T + 1) / 2;
  const long long t_kernel = DAG_CLOCK();
  for (int lin = blockIdx.x; lin < ntasks; lin += gridDim.x) {
    int ti, tj;
    decode_tile(lin, T, T, true, ti, tj);
    const bool diag = ti == tj;
    double acc[S::MI][S::NI][2];
    zero_acc(acc);
    [[maybe_unused]] long long t0 = DAG_CLOCK();
    if (ti > 0) {
      prod.begin(ti, tj);
      for (int p = 0; p < PREFETCH; ++p) prod.issue(warp, lane);
      consume_ktiles<false, S, ISSUE_AT_ONE_TILE>(acc, ring, lm, nullptr, ti * NB, 0, ti * KT_PER_BLOCK, it, warp, lane,
                                                  prod);
      [[maybe_unused]] long long t1 = DAG_CLOCK();
      subtract_acc(H, ld, n, diag, acc, ti * BM + lm.wm * 64, tj * BN + lm.wn * 32, lm.g8, lm.l4);
      DAG_MARK(13, t1);
    }
    DAG_ADD(1, t0);
    t0 = DAG_CLOCK();
    if (!diag && tid == 0) wait_column(done + ti, epoch, ti + 1, fault, info);
    __syncthreads();  // ring drained by every warp, B = A - acc visible to the CTA, U(i, i) published
    DAG_ADD(2, t0);
    t0 = DAG_CLOCK();
    const int k0 = ti * NB, nb = min(NB, n - k0);
    if (diag) {
      potf2_tile<true>(scratch, rs, H + (long long)k0 * ld + k0, ld, nb, k0, info);
      __syncthreads();
      DAG_ADD(3, t0);
    } else {
      const int c0 = tj * NB, ncols = min(NB, n - c0);
      double* Us = scratch;
      double* Ps = scratch + NB * US_LD;
      double* P = H + (long long)k0 * ld + c0;
      [[maybe_unused]] long long t1 = DAG_CLOCK();
      trsm_load_u<true>(Us, H + (long long)k0 * ld + k0, ld, nb);
      DAG_MARK(8, t1);
#pragma unroll 1
      for (int col0 = 0; col0 < ncols; col0 += TP_COLS) {
        const int ncl = min(TP_COLS, ncols - col0);
        const bool vecP = (nb == NB) && (ncl == TP_COLS) && !(ld & 1) && !(((uintptr_t)(P + col0)) & 15);
        t1 = DAG_CLOCK();
        trsm_load_p<true>(Ps, P, ld, nb, col0, ncl, vecP);
        __syncthreads();
        if (col0 == 0) {
          if (tid < NB) rinv[tid] = tid < nb ? 1.0 / Us[tid * US_LD + tid] : 1.0;
          __syncthreads();
        }
        DAG_MARK(9, t1);
        trsm_solve(Us, Ps, rinv, nb);
        t1 = DAG_CLOCK();
        trsm_store_p(Ps, P, ld, nb, col0, ncl, vecP);
        __syncthreads();  // Ps is reloaded by the next half
        DAG_MARK(12, t1);
      }
      DAG_ADD(4, t0);
    }
    t0 = DAG_CLOCK();
    // publish: tile data -> device scope (and the async proxy of the CTAs that will TMA-load it), then the flag;
    // the same fence orders this task's generic accesses to the scratch before the next task's TMA writes to it
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      const unsigned long long v = ((unsigned long long)epoch << 32) | (unsigned)(ti + 1);
      asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(done + tj), "l"(v) : "memory");
    }
    DAG_ADD(5, t0);
#ifdef IPM_DAG_TIMING
    if (tid == 0) g_dag_t[blockIdx.x * 16 + 6] += 1;
#endif
  }
  DAG_ADD(0, t_kernel);
}

#ifdef IPM_DAG_TIMING
}  // namespace dag
// variant builds only: copies the per-CTA cycle counters to the host and clears them
extern "C" int ipm_internal_dag_timing(long long* out, int count) {
  static long long zero[256 * 16];
  if (count > 256 * 16) count = 256 * 16;
  IPM_CUDA_CHECK(cudaDeviceSynchronize());
  IPM_CUDA_CHECK(cudaMemcpyFromSymbol(out, dag::g_dag_t, count * sizeof(long long)));
  IPM_CUDA_CHECK(cudaMemcpyToSymbol(dag::g_dag_t, zero, sizeof(zero)));
  return IPM_OK;
}
namespace dag {
#endif

// ------------------------------------------------------------------------------------------------
// Pipelined variant (potrf_dag2_kernel): same tasks, same schedule, dependencies tracked per 32 ROWS instead of per
// 128-row tile, so that the three stages of the dependent chain of a block row
//     potf2 of U(i, i)   ->   tile solve U(i, i+1) = U(i, i)^{-T} B   ->   last 128 rows of the contraction of (i+1, i+1)
// overlap: the solve consumes U(i, i) one 32-row step behind the factorisation, and the contraction consumes
// U(i, i+1) one k-tile (32 rows) behind the solve.  Chain per block row: potf2 + one solve step + one k-tile + three
// flag hops, instead of potf2 + whole solve + four k-tiles.
//   prog[c]          (c < MAX_COLS) : 32-row groups of the OFF-diagonal tiles of block column c that are final
//                                     (4 per tile, published top-down: tile (k, c) step g -> 4 k + g + 1)
//   prog[MAX_COLS + c]  (c < MAX_T) : 32-row steps of the diagonal tile (c, c) that are final (0 .. 4)
//
// Fused triangular solve (ipm_potrf_trsm_upper_f64): the right-hand sides B (n x p) of  Y = U^{-T} B  are simply TB =
// ceil(p / 128) more block columns of the same DAG -- U(i, j) = U(i, i)^{-T} (A(i, j) - sum_{k<i} U(k, i)^T U(k, j)) is
// the forward substitution when column j belongs to B -- so block row i has the tasks (i, i) .. (i, T - 1) of H followed by
// (i, T) .. (i, T + TB - 1) of B, and ONE launch factors H and solves for all right-hand sides (the stream-ordered TRSM
// took ~150 launches per solve).  The extra columns also keep every SM busy through the last block rows of H.
// The tile solve handles all 128 columns at once (B tile in shared memory, U(i, i) streamed through a 32-row slab),
// and B = A - acc goes from the accumulators straight to shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int MAX_COLS = 2 * MAX_T;  // block columns of H plus block columns of the right-hand sides
constexpr int TILE_LD = NB + 4;  // 132: the potf2 / solve tile in shared memory
constexpr int SLAB_ROWS = 32;
constexpr int SCRATCH2_BYTES = STAGES * STAGE_BYTES;  // ring (192 KiB) >= tile (132 KiB) + slab (33 KiB)
static_assert(SCRATCH2_BYTES >= (NB + SLAB_ROWS) * TILE_LD * 8, "tile + slab alias the TMA ring");
constexpr int SMEM2 = 1024 + SCRATCH2_BYTES + 2 * STAGES * 8;

// Distributed factorisation (ipm_potrf_upper_peer_f64): R GPUs, block column j of H belongs to rank j % R, which runs the
// tasks (i, j) of its columns on its own CTAs.  Every rank keeps the WHOLE matrix (it needs all of U for the triangular
// solves that follow) in peer-mapped memory.  PUSH model: the owner of a finished 32-row step stores it into all R copies
// and then release-stores the column's progress counter of every rank (system scope), so every TMA load and every
// dependency wait stays local and the consumer side of the kernel is the single-GPU one.  R == 1: plain pointers.
constexpr int kMaxRanks = 8;
struct Peers {
  double* H[kMaxRanks];                 // the matrix on every rank (H[me] = local)
  unsigned long long* prog[kMaxRanks];  // the progress counters on every rank
  int* info[kMaxRanks];
  int R, me;
};

struct EmuMaps {
  CUtensorMap m[kMaxRanks];  // tensor map of every emulated rank's matrix copy (emulation mode only)
};

__device__ __forceinline__ int column_progress_sys(const unsigned long long* p, unsigned epoch) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return (unsigned)(v >> 32) == epoch ? (int)(unsigned)v : 0;
}
__device__ __forceinline__ void set_info(const Peers& pr, int* info, int value) {
  if (pr.R == 1) {
    atomicCAS(info, 0, value);
  } else {
    for (int r = 0; r < pr.R; ++r) atomicCAS_system(pr.info[r], 0, value);
  }
}

// Spins until prog[idx] >= need (bounded by the watchdog, common.cuh: spin_wait; on a fault info becomes -1 and the wait
// reports success so that the kernel drains on garbage instead of hanging).
__device__ __forceinline__ int wait_progress(unsigned long long* prog, int idx, unsigned epoch, int need, int* info,
                                             unsigned int* fault, bool sys = false) {
  int r = 0;
  const bool ok = sys ? spin_wait([&] { return (r = column_progress_sys(prog + idx, epoch)) >= need; }, fault,
                                  IPM_FAULT_POTRF_PEER)
                      : spin_wait([&] { return (r = column_progress(prog + idx, epoch)) >= need; }, fault,
                                  IPM_FAULT_POTRF_DAG);
  if (!ok) {
    atomicCAS(info, 0, -1);
    r = 4 * MAX_COLS;
  }
  return r;
}
__device__ __forceinline__ void publish(unsigned long long* p, unsigned epoch, int count) {
  const unsigned long long v = ((unsigned long long)epoch << 32) | (unsigned)count;
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// counter `idx` of every rank (the caller has fenced at system scope when R > 1)
__device__ __forceinline__ void publish_all(const Peers& pr, unsigned long long* prog, int idx, unsigned epoch, int count) {
  if (pr.R == 1) {
    publish(prog + idx, epoch, count);
    return;
  }
  const unsigned long long v = ((unsigned long long)epoch << 32) | (unsigned)count;
  for (int r = 0; r < pr.R; ++r)
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pr.prog[r] + idx), "l"(v) : "memory");
}

struct Producer2 {
  const CUtensorMap* tm;
  const CUtensorMap* tmB;  // right-hand sides (block columns >= T), or null
  int T;
  Ring ring;
  unsigned long long* prog;
  int* info;
  unsigned int* fault;
  unsigned epoch;
  bool sys;  // counters are written by other GPUs: system-scope acquire
  int ti, tj, kt, k1, ready_i, ready_j;  // ready_*: 32-row groups of columns ti / tj known to be final
  uint32_t it;
  __device__ __forceinline__ void begin(int i, int j) {
    ti = i, tj = j, kt = 0, k1 = i * KT_PER_BLOCK, ready_i = ready_j = 0;
  }
  __device__ __forceinline__ void issue(int wp, int lane) {
    if (kt >= k1) return;
    if (lane == 0) {
      const int need = kt + 1;  // BK == 32 rows == one group
      if (ready_i < need || ready_j < need) {
        if (ready_i < need) ready_i = wait_progress(prog, ti, epoch, need, info, fault, sys);
        if (tj == ti) ready_j = ready_i;
        if (ready_j < need) ready_j = wait_progress(prog, tj, epoch, need, info, fault, sys);
        asm volatile("fence.proxy.async;" ::: "memory");  // the rows were written through the generic proxy
      }
      const uint32_t s = it % STAGES;
      if (it >= STAGES) mbar_wait(ring.empty0 + 8 * s, ((it / STAGES) - 1) & 1);
      const uint32_t full = ring.full0 + 8 * s;
      mbar_expect_tx(full, 2 * CHUNK_BYTES);
      const uint32_t dstA = ring.tiles0 + s * STAGE_BYTES + wp * CHUNK_BYTES;
      tma_load_2d(dstA, tm, ti * BM + wp * 16, kt * BK, full);
      if (tj < T) {
        tma_load_2d(dstA + OPERAND_BYTES, tm, tj * BN + wp * 16, kt * BK, full);
      } else {
        tma_load_2d(dstA + OPERAND_BYTES, tmB, (tj - T) * BN + wp * 16, kt * BK, full);
      }
    }
    __syncwarp();
    ++it;
    ++kt;
  }
};
static_assert(BK == SLAB_ROWS, "one k-tile == one 32-row group");

// Tile (row0.., col0..) of B = A - acc from the accumulators into shared memory S (ld TILE_LD): entries outside the
// matrix are 0 (1 on the diagonal of a diagonal tile: identity padding), the strict lower triangle of a diagonal tile 0.
__device__ __forceinline__ void stage_tile(double* __restrict__ S, const double* __restrict__ H, long long ld, int n,
                                           int ncols, int row0, int col0, bool diag, const double (&acc)[8][4][2],
                                           const LaneMap& lm) {
  const int lr0 = lm.wm * 64, lc0 = lm.wn * 32;
  const bool interior = !(ld & 1) && !(((uintptr_t)H) & 15) && (row0 + lr0 + 64 <= n) && (col0 + lc0 + 32 <= ncols) &&
                        (!diag || lc0 >= lr0 + 63);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int lr = lr0 + i * 8 + lm.g8, row = row0 + lr;
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) {
      const int lc = lc0 + jn * 8 + 2 * lm.l4, col = col0 + lc;
      double2 v;
      if (interior) {
        const double2 a = __ldcg(reinterpret_cast<const double2*>(H + (long long)row * ld + col));
        v = make_double2(a.x - acc[i][jn][0], a.y - acc[i][jn][1]);
      } else {
        const bool in0 = row < n && col < ncols && !(diag && col < row);
        const bool in1 = row < n && col + 1 < ncols && !(diag && col + 1 < row);
        v.x = in0 ? __ldcg(H + (long long)row * ld + col) - acc[i][jn][0] : ((diag && lr == lc && row >= n) ? 1.0 : 0.0);
        v.y = in1 ? __ldcg(H + (long long)row * ld + col + 1) - acc[i][jn][1]
                  : ((diag && lr == lc + 1 && row >= n) ? 1.0 : 0.0);
      }
      *reinterpret_cast<double2*>(S + lr * TILE_LD + lc) = v;
    }
  }
}

// Rows [r_begin, r_begin + 32) of the shared tile S -> global tile G (rows < nrows, columns [cmin(row), ncols)).
template <bool UPPER>
__device__ __forceinline__ void store_rows(const double* __restrict__ S, double* __restrict__ G, long long ld,
                                           int r_begin, int nrows, int ncols) {
  const bool vec = !(ld & 1) && !(((uintptr_t)G) & 15);
  for (int idx = threadIdx.x; idx < SLAB_ROWS * (NB / 2); idx += THREADS) {
    const int r = r_begin + (idx >> 6), c = (idx & 63) * 2;
    if (r >= nrows) continue;
    const int cmin = UPPER ? r : 0;
    double* p = G + (long long)r * ld + c;
    if (vec && c >= cmin && c + 1 < ncols) {
      *reinterpret_cast<double2*>(p) = make_double2(S[r * TILE_LD + c], S[r * TILE_LD + c + 1]);
    } else {
      if (c >= cmin && c < ncols) p[0] = S[r * TILE_LD + c];
      if (c + 1 >= cmin && c + 1 < ncols) p[1] = S[r * TILE_LD + c + 1];
    }
  }
}

// Distributed factorisation: the finished rows go into every rank's copy of the matrix.  One SM stores to a peer at only
// ~15 GB/s (measured: a 32-row step to 7 peers, 224 KB, took ~15 us and put ~200 us per block row on the dependent chain
// when all copies were stored before anything was published), so the copies are served in order of urgency, in two
// groups with one system fence each:  group 1 = the local copy and the copy of rank `first` -- the owner of the task
// that consumes these rows next on the dependent chain -- published at once;  group 2 = the other ranks, after the
// caller's own next piece of work.
template <bool UPPER>
__device__ __forceinline__ void store_rows_urgent(const Peers& pr, int me, int first, const double* __restrict__ S,
                                                  double* __restrict__ G, long long ld, int r_begin, int nrows,
                                                  int ncols) {
  store_rows<UPPER>(S, G, ld, r_begin, nrows, ncols);
  if (pr.R > 1 && first != me) store_rows<UPPER>(S, pr.H[first] + (G - pr.H[me]), ld, r_begin, nrows, ncols);
}
// after the caller's fence + CTA barrier (thread 0 only)
__device__ __forceinline__ void publish_urgent(const Peers& pr, int me, int first, unsigned long long* prog, int idx,
                                               unsigned epoch, int count) {
  publish(prog + idx, epoch, count);
  if (pr.R > 1 && first != me) {
    const unsigned long long v = ((unsigned long long)epoch << 32) | (unsigned)count;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pr.prog[first] + idx), "l"(v) : "memory");
  }
}
// group 2: every rank except this one and `first`; ends with a CTA barrier (all threads call it)
template <bool UPPER>
__device__ __forceinline__ void push_rest(const Peers& pr, int me, int first, const double* __restrict__ S,
                                          double* __restrict__ G, long long ld, int r_begin, int nrows, int ncols, int idx,
                                          unsigned epoch, int count) {
  if (pr.R <= 2 && (pr.R == 1 || first != me)) return;  // nobody left
  const long long off = G - pr.H[me];
  for (int r = 1; r < pr.R; ++r) {
    const int dest = (me + r) % pr.R;
    if (dest != first) store_rows<UPPER>(S, pr.H[dest] + off, ld, r_begin, nrows, ncols);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long v = ((unsigned long long)epoch << 32) | (unsigned)count;
    for (int r = 1; r < pr.R; ++r) {
      const int dest = (me + r) % pr.R;
      if (dest != first)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pr.prog[dest] + idx), "l"(v) : "memory");
    }
  }
}

// Diagonal task: potf2 of the tile in S, publishing every 32-row step (rows final after the pivot block and its row
// panel) before the trailing update of the step.
__device__ __forceinline__ void potf2_pipelined(const Peers& pr, int me, double* __restrict__ S, double* __restrict__ rs,
                                                double* __restrict__ G, long long ld, int nb, int k0,
                                                int* __restrict__ info, unsigned long long* prog, int diag_idx,
                                                unsigned epoch, int first) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll 1
  for (int base = 0; base < nb; base += 32) {
    if (warp == 0) {
      double d[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) d[i] = S[(base + i) * TILE_LD + base + lane];
      int bad = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const double piv = __shfl_sync(0xffffffffu, d[j], j);
        bad = (bad == 0 && !(piv > 0.0)) ? j + 1 : bad;
        const double r = rsqrt_nobranch(piv);
        const double u = d[j] * r;
        d[j] = u;
        if (lane == j) rs[j] = r;
#pragma unroll
        for (int i = j + 1; i < 32; ++i) {
          const double ui = __shfl_sync(0xffffffffu, u, i);
          d[i] = fma(-ui, u, d[i]);
        }
      }
      if (bad && lane == 0) set_info(pr, info, k0 + base + bad);
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (lane >= i) S[(base + i) * TILE_LD + base + lane] = d[i];
    }
    __syncthreads();
    const int W = NB - base - 32;  // 96, 64, 32, 0
    if (tid < W) {
      const int c = base + 32 + tid;
      double v[32];
#pragma unroll
      for (int l = 0; l < 32; ++l) v[l] = S[(base + l) * TILE_LD + c];
#pragma unroll
      for (int l = 0; l < 32; ++l) {
        const double x = v[l] * rs[l];
        v[l] = x;
#pragma unroll
        for (int r = l + 1; r < 32; ++r) v[r] = fma(-S[(base + l) * TILE_LD + base + r], x, v[r]);
      }
#pragma unroll
      for (int l = 0; l < 32; ++l) S[(base + l) * TILE_LD + c] = v[l];
    }
    __syncthreads();
    store_rows_urgent<true>(pr, me, first, S, G, ld, base, nb, nb);
    if (pr.R == 1) __threadfence(); else __threadfence_system();
    __syncthreads();
    if (tid == 0) publish_urgent(pr, me, first, prog, diag_idx, epoch, (base >> 5) + 1);
    if (W > 0) {
      potf2_trailing_update(S, base, W, warp, lane);
      __syncthreads();
    }
    push_rest<true>(pr, me, first, S, G, ld, base, nb, nb, diag_idx, epoch, (base >> 5) + 1);
  }
}

// Off-diagonal task (ti, tj): X = U(ti, ti)^{-T} B for the B tile in Ps (all 128 columns), one 32-row step behind the
// factorisation of U(ti, ti); every finished step is stored and published.
__device__ __forceinline__ void solve_pipelined(const Peers& pr, int me, double* __restrict__ Ps, double* __restrict__ Us,
                                                double* __restrict__ rinv, const double* __restrict__ Uii,
                                                long long ld, double* __restrict__ G, long long ldg, int nrows,
                                                int ncols, unsigned long long* prog, int ti, int tj, unsigned epoch,
                                                int* __restrict__ info, unsigned int* fault) {
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5, l4 = lane & 3, g8 = lane >> 2;
  // rows of U(ti, tj) are the A operand of block row tj: its diagonal task is ours, (tj, tj + 1) is rank (tj + 1) % R's
  const int first = pr.R > 1 ? (tj + 1) % pr.R : 0;
  const bool vecU = !(ld & 1) && !(((uintptr_t)Uii) & 15);
  int dready = 0;
#pragma unroll 1
  for (int b0 = 0; b0 < NB; b0 += 32) {
    const int g = b0 >> 5;
    if (tid == 0 && dready < g + 1)  // cached by thread 0
      dready = wait_progress(prog, MAX_COLS + ti, epoch, g + 1, info, fault, pr.R > 1);
    __syncthreads();  // U rows published (thread 0 acquired); previous step's readers of the slab are done
    // slab: rows b0 .. b0+31 of U(ti, ti), columns >= row
    for (int idx = tid; idx < SLAB_ROWS * (NB / 2); idx += THREADS) {
      const int r = idx >> 6, cc = (idx & 63) * 2, row = b0 + r;
      double2 v = make_double2(0.0, 0.0);
      if (row >= nrows) {  // ragged last row block (right-hand-side columns only): identity padding, nothing to read
        v = make_double2(cc == row ? 1.0 : 0.0, cc + 1 == row ? 1.0 : 0.0);
      } else if (cc + 1 >= row) {
        if (vecU) {
          v = __ldcg(reinterpret_cast<const double2*>(Uii + (long long)row * ld + cc));
        } else {
          v.x = __ldcg(Uii + (long long)row * ld + cc);
          v.y = __ldcg(Uii + (long long)row * ld + cc + 1);
        }
      }
      Us[r * TILE_LD + cc] = cc >= row ? v.x : 0.0;
      Us[r * TILE_LD + cc + 1] = v.y;
    }
    __syncthreads();
    if (tid < SLAB_ROWS) rinv[tid] = 1.0 / Us[tid * TILE_LD + b0 + tid];
    __syncthreads();
    if (tid < NB) {
      const int c = tid;
      double v[32];
#pragma unroll
      for (int l = 0; l < 32; ++l) v[l] = Ps[(b0 + l) * TILE_LD + c];
#pragma unroll
      for (int l = 0; l < 32; ++l) {
        const double x = v[l] * rinv[l];
        v[l] = x;
        const double* urow = Us + l * TILE_LD + b0;
#pragma unroll
        for (int r = l + 1; r < 32; ++r) v[r] = fma(-urow[r], x, v[r]);
      }
#pragma unroll
      for (int l = 0; l < 32; ++l) Ps[(b0 + l) * TILE_LD + c] = v[l];
    }
    __syncthreads();
    store_rows_urgent<false>(pr, me, first, Ps, G, ldg, b0, nrows, ncols);  // rows b0 .. b0+31 of X are final
    // rows below:  Ps[r0.., :] -= U[b0..b0+32, r0..]^T X[b0..b0+32, :]; warp wp owns columns 16 wp .. 16 wp + 15
    {
      double bf[2][8];
#pragma unroll
      for (int cb = 0; cb < 2; ++cb)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) bf[cb][kk] = Ps[(b0 + 4 * kk + l4) * TILE_LD + 16 * wp + 8 * cb + g8];
      for (int r0 = b0 + 32; r0 < NB; r0 += 32) {
        double2 cacc[2][4];
        double af[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int cb = 0; cb < 2; ++cb)
            cacc[cb][i] =
                *reinterpret_cast<const double2*>(Ps + (r0 + 8 * i + g8) * TILE_LD + 16 * wp + 8 * cb + 2 * l4);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) af[i][kk] = -Us[(4 * kk + l4) * TILE_LD + r0 + 8 * i + g8];
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int cb = 0; cb < 2; ++cb)
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(cacc[cb][i].x), "+d"(cacc[cb][i].y)
                           : "d"(af[i][kk]), "d"(bf[cb][kk]));
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int cb = 0; cb < 2; ++cb)
            *reinterpret_cast<double2*>(Ps + (r0 + 8 * i + g8) * TILE_LD + 16 * wp + 8 * cb + 2 * l4) = cacc[cb][i];
      }
    }
    if (pr.R == 1) __threadfence(); else __threadfence_system();
    asm volatile("fence.proxy.async;" ::: "memory");  // the rows will be read by TMA
    __syncthreads();
    if (tid == 0) publish_urgent(pr, me, first, prog, tj, epoch, 4 * ti + g + 1);
    push_rest<false>(pr, me, first, Ps, G, ldg, b0, nrows, ncols, tj, epoch, 4 * ti + g + 1);
  }
}

__global__ void __launch_bounds__(THREADS, 1)
potrf_dag2_kernel(const __grid_constant__ CUtensorMap tm_, const __grid_constant__ CUtensorMap tmB,
                  double* __restrict__ H, long long ld, int n, int T, double* __restrict__ Bm, long long ldb, int p,
                  int TB, int* __restrict__ info, unsigned long long* __restrict__ prog, unsigned epoch,
                  unsigned int* __restrict__ fault, const __grid_constant__ Peers pr,
                  const __grid_constant__ EmuMaps emu, int emu_ctas) {
  // emu_ctas > 0 (single-GPU test of the distributed code, ipm_internal_potrf_peer_emulated_f64): ONE cooperative grid
  // holds all R ranks -- CTAs [r * emu_ctas, (r + 1) * emu_ctas) act as rank r on rank r's copy of everything.  (Ranks
  // as separate launches on one device may never be co-resident.)
  const CUtensorMap* tmp = &tm_;
  int cta = blockIdx.x, ncta = gridDim.x, me = pr.me;
  if (emu_ctas > 0) {
    me = blockIdx.x / emu_ctas;
    cta = blockIdx.x - me * emu_ctas, ncta = emu_ctas;
    H = pr.H[me], prog = pr.prog[me], info = pr.info[me];
    tmp = &emu.m[me];
  }
  const CUtensorMap& tm = *tmp;
  extern __shared__ uint8_t smem_raw[];
  __shared__ double rinv[SLAB_ROWS];
  __shared__ double rs[32];
  using S = Shape128x128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  double* tile = reinterpret_cast<double*>(smem);   // NB x TILE_LD; aliases the TMA ring
  double* slab = tile + NB * TILE_LD;               // SLAB_ROWS x TILE_LD
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SCRATCH2_BYTES);
  const Ring ring{smem_u32(bars), smem_u32(bars + STAGES), smem_u32(smem)};
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(ring.full0 + 8 * s, CONSUMER_WARPS);
      mbar_init(ring.empty0 + 8 * s, CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();
  const LaneMap lm = make_lane_map<S>(warp, lane);
  Producer2 prod{&tm, &tmB, T, ring, prog, info, fault, epoch, pr.R > 1, 0, 0, 0, 0, 0, 0, 0u};
  uint32_t it = 0;
  // Task list of this rank, row-major: R == 1: all (ti, tj), tj in [ti, T + TB); R > 1: the columns tj = me (mod R).
  // The CTA runs tasks blockIdx.x, blockIdx.x + gridDim.x, ... of that list; `row` / `row_base` walk the rows.
  const int R = pr.R;
  const int ntasks = T * (T + 1) / 2 + T * TB;  // R == 1
  int row = 0, row_base = 0;
  for (int lin = cta;; lin += ncta) {
    int ti, tj;
    if (R == 1) {
      if (lin >= ntasks) break;
      decode_tile(lin, T, T + TB, true, ti, tj);  // row ti: columns ti .. T + TB - 1 (block columns >= T: right-hand sides)
    } else {
      int first = 0, cnt = 0;
      while (row < T) {
        first = row + ((me - row) % R + R) % R;  // first column >= row owned by this rank
        cnt = first < T ? (T - 1 - first) / R + 1 : 0;
        if (lin < row_base + cnt) break;
        row_base += cnt;
        ++row;
      }
      if (row >= T) break;
      ti = row;
      tj = first + (lin - row_base) * R;
    }
    const bool diag = ti == tj;
    double acc[S::MI][S::NI][2];
    zero_acc(acc);
    if (ti > 0) {
      prod.begin(ti, tj);
      for (int p = 0; p < PREFETCH; ++p) prod.issue(warp, lane);
      consume_ktiles<false, S, ISSUE_AT_ONE_TILE>(acc, ring, lm, nullptr, ti * NB, 0, ti * KT_PER_BLOCK, it, warp, lane,
                                                  prod);
      __syncthreads();  // every warp is done with the ring before it becomes the tile
    }
    const int k0 = ti * NB;
    const bool rhs = tj >= T;
    double* M = rhs ? Bm : H;                       // matrix the tile lives in
    const long long ldm = rhs ? ldb : ld;
    const int c0 = rhs ? (tj - T) * NB : tj * NB, mcols = rhs ? p : n;
    stage_tile(tile, M, ldm, n, mcols, k0, c0, diag, acc, lm);
    __syncthreads();
    if (diag) {
      // rows of U(ti, ti) are consumed first by the solve of (ti, ti + 1): rank (ti + 1) % R
      potf2_pipelined(pr, me, tile, rs, H + (long long)k0 * ld + k0, ld, min(NB, n - k0), k0, info, prog, MAX_COLS + ti,
                      epoch, pr.R > 1 ? (ti + 1) % pr.R : 0);
      if (tid == 0) publish_all(pr, prog, MAX_COLS + ti, epoch, 4);  // a ragged last tile has fewer than four steps
    } else {
      solve_pipelined(pr, me, tile, slab, rinv, H + (long long)k0 * ld + k0, ld, M + (long long)k0 * ldm + c0, ldm,
                      min(NB, n - k0), min(NB, mcols - c0), prog, ti, tj, epoch, info, fault);
    }
    // this task's generic accesses to the tile / slab precede the next task's TMA writes into the same bytes
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
  }
  // Distributed: the kernel -- and with it everything the stream runs next on this rank's copy of U -- ends only when
  // the other ranks' columns have arrived in full (off-diagonal tiles of column c: 4 c groups, diagonal tile: 4 steps).
  if (R > 1 && cta == 0) {
    for (int c = tid; c < T; c += THREADS) {
      wait_progress(prog, c, epoch, 4 * c, info, fault, true);
      wait_progress(prog, MAX_COLS + c, epoch, 4, info, fault, true);
    }
  }
}

struct Slot {
  cudaStream_t stream;
  unsigned long long* done;  // progress counters: MAX_COLS per block column + MAX_T per diagonal tile (pipelined kernel)
  bool used;
};
constexpr int kMaxDev = kMaxDevices, kSlotsPerDev = 4;
Slot g_slots[kMaxDev][kSlotsPerDev];
int g_sms[kMaxDev];   // CTAs of potrf_dag_kernel that can be resident at once on the device (occupancy x SM count)
std::mutex g_mutex;
std::atomic<unsigned> g_epoch{0};

// progress counters of (device, stream); nullptr when all slots belong to other streams (caller takes the
// stream-ordered path)
unsigned long long* get_counters(int dev, cudaStream_t st, int* rc) {
  std::lock_guard<std::mutex> lock(g_mutex);
  *rc = IPM_OK;
  for (int i = 0; i < kSlotsPerDev; ++i)
    if (g_slots[dev][i].used && g_slots[dev][i].stream == st) return g_slots[dev][i].done;
  for (int i = 0; i < kSlotsPerDev; ++i) {
    Slot* s = &g_slots[dev][i];
    if (s->used) continue;
    if (cudaMalloc(&s->done, (MAX_COLS + MAX_T) * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(s->done, 0, (MAX_COLS + MAX_T) * sizeof(unsigned long long)) != cudaSuccess) {
      *rc = ipm_set_cuda_error(cudaGetLastError());
      return nullptr;
    }
    s->stream = st;
    s->used = true;
    return s->done;
  }
  return nullptr;
}

// Which sizes ipm_potrf_upper_f64 sends through the single-launch tile-DAG kernel (the pipelined one).  Default:
// n >= 2048, where it wins on B200 (profiles/potrf_ab_r02a.txt, stream-ordered / tile-DAG / pipelined tile-DAG: 1.24 /
// 1.67 / 1.07 ms at n = 2048, 2.78 / 3.37 / 2.12 ms at 4096, 8.91 / 7.71 / 6.71 ms at 8192, 51.1 / 44.8 / 44.9 ms at
// 16384).  IPM_POTRF_DAG=1 forces it for every admissible size (the test suite runs once that way), =0 switches it off.
constexpr int kAutoMinN = 2048;
bool enabled(int n) {
  static const int mode = [] {
    const char* e = getenv("IPM_POTRF_DAG");
    return !e ? -1 : (e[0] == '1' ? 1 : (e[0] == '0' ? 0 : -1));
  }();
  return mode == 1 || (mode == -1 && n >= kAutoMinN);
}

// returns 1 when the problem is not handled here (caller falls through to the stream-ordered factorisation)
// Bm / ldb / p: optional right-hand sides solved in the same launch (pipelined kernel only), Bm <- U^{-T} Bm.
// peers / peer_epoch / max_ctas: distributed factorisation (see struct Peers): counters, info and the matrix live in
// peer-mapped memory owned by the caller, the epoch is the caller's (identical on every rank).
int potrf(double* H, int ld, int n, int* info_dev, cudaStream_t st, bool pipelined = true, double* Bm = nullptr,
          int ldb = 0, int p = 0, const Peers* peers = nullptr, unsigned peer_epoch = 0, int max_ctas = 0,
          bool emulate = false) {
  const int T = ceil_div(n, NB), TB = Bm ? ceil_div(p, NB) : 0;
  if (T < 3 || T > MAX_T || T + TB > MAX_COLS || (Bm && !pipelined) || (peers && (Bm || !pipelined))) return 1;
  int dev = 0, rc = IPM_OK;
  IPM_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) return 1;
  static bool attr_set[kMaxDevices];
  IPM_CUDA_CHECK(ensure_dynamic_smem(potrf_dag_kernel, SMEM, attr_set));
  if (!g_sms[dev]) {
    int sms = 0, per_sm = 0;
    IPM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    IPM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, potrf_dag_kernel, THREADS, SMEM));
    if (per_sm < 1) return 1;
    g_sms[dev] = sms * per_sm;
  }
  unsigned long long* done = peers ? peers->prog[peers->me] : get_counters(dev, st, &rc);
  if (rc) return rc;
  if (!done) return 1;
  CUtensorMap tm, tmB;
  if (make_operand_map(&tm, H, ld, n, n)) return 1;  // unaligned base: the stream-ordered path reports it
  tmB = tm;
  if (Bm && make_operand_map(&tmB, Bm, ldb, n, p)) return 1;
  unsigned epoch = peer_epoch;
  if (!peers) {
    epoch = ++g_epoch;
    if (epoch == 0) epoch = ++g_epoch;  // never 0 (the counters' initial value)
  }
  Peers pr;
  if (peers) {
    pr = *peers;
  } else {
    pr.R = 1, pr.me = 0;
    for (int r = 0; r < kMaxRanks; ++r) pr.H[r] = nullptr, pr.prog[r] = nullptr, pr.info[r] = nullptr;
    pr.H[0] = H, pr.prog[0] = done, pr.info[0] = info_dev;
  }
  int ntasks = T * (T + 1) / 2 + T * TB;
  if (peers) {  // tasks of this rank's columns
    ntasks = 0;
    for (int j = pr.me; j < T; j += pr.R) ntasks += j + 1;
    if (ntasks == 0) ntasks = 1;  // a rank without columns still runs the final wait
  }
  int grid = ntasks < g_sms[dev] ? ntasks : g_sms[dev];
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  static EmuMaps emu;  // only filled in emulation mode (a kernel parameter either way)
  int emu_ctas = 0;
  if (emulate) {
    if (!peers) return IPM_ERR_ARG;
    emu_ctas = g_sms[dev] / pr.R;
    if (emu_ctas < 1) return IPM_ERR_ARG;
    grid = emu_ctas * pr.R;
    for (int r = 0; r < pr.R; ++r)
      if (make_operand_map(&emu.m[r], pr.H[r], ld, n, n)) return IPM_ERR_ARG;
  }
  // cooperative: the CTAs wait for each other's tiles, so all of them must be resident (a second persistent kernel on
  // another stream, MPS SM limits or green contexts would otherwise leave round >= 1 tasks waiting for CTAs that are
  // never scheduled); every wait is additionally bounded by the watchdog
  if (pipelined) {
    static bool attr2_set[kMaxDevices];
    IPM_CUDA_CHECK(ensure_dynamic_smem(potrf_dag2_kernel, SMEM2, attr2_set));
    IPM_CUDA_CHECK(launch_cooperative(potrf_dag2_kernel, dim3(grid), dim3(THREADS), SMEM2, st, tm, tmB, H, (long long)ld,
                                      n, T, Bm, (long long)ldb, p, TB, info_dev, done, epoch,
                                      ipm_internal_fault_word(), pr, emu, emu_ctas));
  } else {
    IPM_CUDA_CHECK(launch_cooperative(potrf_dag_kernel, dim3(grid), dim3(THREADS), SMEM, st, tm, H, (long long)ld, n, T,
                                      info_dev, done, epoch, ipm_internal_fault_word()));
  }
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}
}  // namespace dag

// Library-owned side stream + events for the look-ahead (one set per device, created on first use).
struct SideStream {
  cudaStream_t stream;
  cudaEvent_t fork, join, chain, u2;
  bool ready;
};
static SideStream g_side[16];
static std::mutex g_side_mutex;  // creation only; the look-ahead itself is per (device, caller) stream-ordered

static int get_side_stream(SideStream** out) {
  int dev = 0;
  IPM_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return IPM_ERR_ARG;
  SideStream* s = &g_side[dev];
  std::lock_guard<std::mutex> lock(g_side_mutex);
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
// for each outer block of NBO = 384 rows:   [ potf2(128) ; trsm of its 128-row panel ; K=128 update of the block's
// remaining rows ] x 3 panels, then ONE K=384 DMMA update of the trailing matrix.
// ------------------------------------------------------------------------------------------------
static int potrf_stream_ordered(double* H, int ld, int n, int* info_dev, void* stream);

extern "C" int ipm_potrf_upper_f64(double* H, int ld, int n, int* info_dev, void* stream) {
  if (!H || !info_dev || n < 0 || ld < n || (ld & 1)) return IPM_ERR_ARG;
  IPM_CUDA_CHECK(cudaMemsetAsync(info_dev, 0, sizeof(int), (cudaStream_t)stream));
  if (dag::enabled(n)) {
    const int rc = dag::potrf(H, ld, n, info_dev, (cudaStream_t)stream);
    if (rc <= 0) return rc;  // done or failed; 1 = not handled there
  }
  return potrf_stream_ordered(H, ld, n, info_dev, stream);
}

// Same contract as ipm_potrf_upper_f64, always through the single-launch tile-DAG kernel when the size allows it
// (256 <= n <= 32768 and a free flag slot for the stream), else the stream-ordered code.
extern "C" int ipm_potrf_upper_dag_f64(double* H, int ld, int n, int* info_dev, void* stream) {
  if (!H || !info_dev || n < 0 || ld < n || (ld & 1)) return IPM_ERR_ARG;
  IPM_CUDA_CHECK(cudaMemsetAsync(info_dev, 0, sizeof(int), (cudaStream_t)stream));
  const int rc = dag::potrf(H, ld, n, info_dev, (cudaStream_t)stream);
  if (rc <= 0) return rc;
  return potrf_stream_ordered(H, ld, n, info_dev, stream);
}

extern "C" int ipm_trsm_upper_t_f64(const double* U, int ldu, int n, double* B, int ldb, int p, void* stream);

// H = U^T U in place AND B <- U^{-T} B (B: n x p row-major) -- the first two steps of the block elimination
// (NewtonSolverInfeasibleStart.py:398-426: cho_factor(H), cho_solve halves on A^T) -- in ONE persistent launch when the
// tile-DAG kernel covers the size; otherwise ipm_potrf_upper_f64 followed by ipm_trsm_upper_t_f64.
extern "C" int ipm_potrf_trsm_upper_f64(double* H, int ld, int n, double* B, int ldb, int p, int* info_dev,
                                        void* stream) {
  if (!H || !B || !info_dev || n < 0 || p < 0 || ld < n || ldb < p || (ld & 1) || (ldb & 1)) return IPM_ERR_ARG;
  IPM_CUDA_CHECK(cudaMemsetAsync(info_dev, 0, sizeof(int), (cudaStream_t)stream));
  if (p > 0 && dag::enabled(n)) {
    const int rc = dag::potrf(H, ld, n, info_dev, (cudaStream_t)stream, true, B, ldb, p);
    if (rc <= 0) return rc;
  }
  int rc = ipm_potrf_upper_f64(H, ld, n, info_dev, stream);
  if (rc) return rc;
  return ipm_trsm_upper_t_f64(H, ld, n, B, ldb, p, stream);
}

// Distributed factorisation over R <= 8 GPUs of one node (north_star: "replicated or 2D-block-cyclic factorisation" --
// here block columns dealt cyclically, push model over peer memory; see struct dag::Peers).  Every rank calls it at the
// same point of its stream with its own `me`:
//   peer_H[r]    : rank r's copy of the n x n matrix (leading dimension ld on every rank), upper triangle holds H
//   peer_info[r] : rank r's info word (all of them receive the index of a non-positive pivot, or -1 from the watchdog)
//   peer_prog[r] : rank r's progress counters, ipm_potrf_peer_prog_words() 64-bit words, zeroed once at allocation
//   epoch        : != 0, different from the previous call's, identical on all ranks
//   max_ctas     : 0 = one CTA per SM; smaller values let several "ranks" share one device (single-GPU tests)
// On return (stream order) the rank's copy holds the complete factor U.  Replaces the same call sites as
// ipm_potrf_upper_f64 in the row-sharded engine (NewtonSolver.py:286,303).
extern "C" int ipm_potrf_peer_prog_words(void) { return dag::MAX_COLS + dag::MAX_T; }
extern "C" int ipm_potrf_upper_peer_f64(void* const* peer_H, int ld, int n, void* const* peer_info,
                                        void* const* peer_prog, int me, int R, unsigned int epoch, int max_ctas,
                                        void* stream) {
  if (!peer_H || !peer_info || !peer_prog || n < 0 || ld < n || (ld & 1) || R < 1 || R > dag::kMaxRanks || me < 0 ||
      me >= R || epoch == 0)
    return IPM_ERR_ARG;
  dag::Peers pr;
  pr.R = R, pr.me = me;
  for (int r = 0; r < dag::kMaxRanks; ++r) {
    pr.H[r] = r < R ? (double*)peer_H[r] : nullptr;
    pr.prog[r] = r < R ? (unsigned long long*)peer_prog[r] : nullptr;
    pr.info[r] = r < R ? (int*)peer_info[r] : nullptr;
    if (r < R && (!pr.H[r] || !pr.prog[r] || !pr.info[r])) return IPM_ERR_ARG;
  }
  const int rc = dag::potrf(pr.H[me], ld, n, pr.info[me], (cudaStream_t)stream, true, nullptr, 0, 0, &pr, epoch, max_ctas);
  return rc == 1 ? IPM_ERR_ARG : rc;  // sizes the tile-DAG kernel does not take (n <= 256): the caller factors replicated
}

// Library-internal (kernel tests on ONE GPU): the distributed factorisation with all R ranks inside one cooperative
// grid, each rank's CTAs working on that rank's copy of the matrix / counters / info.
extern "C" int ipm_internal_potrf_peer_emulated_f64(void* const* peer_H, int ld, int n, void* const* peer_info,
                                                    void* const* peer_prog, int R, unsigned int epoch, void* stream) {
  if (!peer_H || !peer_info || !peer_prog || n < 0 || ld < n || (ld & 1) || R < 2 || R > dag::kMaxRanks || epoch == 0)
    return IPM_ERR_ARG;
  dag::Peers pr;
  pr.R = R, pr.me = 0;
  for (int r = 0; r < dag::kMaxRanks; ++r) {
    pr.H[r] = r < R ? (double*)peer_H[r] : nullptr;
    pr.prog[r] = r < R ? (unsigned long long*)peer_prog[r] : nullptr;
    pr.info[r] = r < R ? (int*)peer_info[r] : nullptr;
  }
  const int rc = dag::potrf(pr.H[0], ld, n, pr.info[0], (cudaStream_t)stream, true, nullptr, 0, 0, &pr, epoch, 0, true);
  return rc == 1 ? IPM_ERR_ARG : rc;
}

// Library-internal (kernel tests): the fused launch for every admissible size, whatever the size policy says.
extern "C" int ipm_internal_potrf_trsm_dag_f64(double* H, int ld, int n, double* B, int ldb, int p, int* info_dev,
                                               void* stream) {
  if (!H || !B || !info_dev || n < 0 || p <= 0 || ld < n || ldb < p || (ld & 1) || (ldb & 1)) return IPM_ERR_ARG;
  IPM_CUDA_CHECK(cudaMemsetAsync(info_dev, 0, sizeof(int), (cudaStream_t)stream));
  const int rc = dag::potrf(H, ld, n, info_dev, (cudaStream_t)stream, true, B, ldb, p);
  return rc == 1 ? IPM_ERR_ARG : rc;
}

// Library-internal (A/B timing in tools/ and bench.py --factorisation): always the stream-ordered code.
extern "C" int ipm_internal_potrf_stream_f64(double* H, int ld, int n, int* info_dev, void* stream) {
  if (!H || !info_dev || n < 0 || ld < n || (ld & 1)) return IPM_ERR_ARG;
  IPM_CUDA_CHECK(cudaMemsetAsync(info_dev, 0, sizeof(int), (cudaStream_t)stream));
  return potrf_stream_ordered(H, ld, n, info_dev, stream);
}

// Library-internal (A/B timing): the first tile-DAG kernel, which tracks dependencies per 128-row tile.
extern "C" int ipm_internal_potrf_dag1_f64(double* H, int ld, int n, int* info_dev, void* stream) {
  if (!H || !info_dev || n < 0 || ld < n || (ld & 1)) return IPM_ERR_ARG;
  IPM_CUDA_CHECK(cudaMemsetAsync(info_dev, 0, sizeof(int), (cudaStream_t)stream));
  const int rc = dag::potrf(H, ld, n, info_dev, (cudaStream_t)stream, false);
  if (rc <= 0) return rc;
  return potrf_stream_ordered(H, ld, n, info_dev, stream);
}

static int potrf_stream_ordered(double* H, int ld, int n, int* info_dev, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  // Look-ahead: the serial panel chain of outer block o+1 (two potf2 + two panel solves, single-CTA latency
  // bound) runs on a side stream concurrently with the bulk trailing update of block o on the caller's stream.
  //   side : chain(0), U1(0), chain(1), [wait U2(0)] U1(1), chain(2), ...
  //   main :            [wait chain(0)] U2(0), [wait chain(1)] U2(1), ...
  // U1(o) = update of the NEXT block's NBO rows (all the next chain needs); U2(o) = rows below them.
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
      int rcp = launch_potf2(Akk, ld, nb, k0, info_dev, cs);
      if (rcp) return rcp;
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
    if (rest > 0) {  // all rows below the outer block, K = NBO
      int rc = ipm_gemm_tn_f64(U + (long long)o0 * ldu + oend, ldu, B + (long long)o0 * ldb, ldb, nullptr, -1.0, 1.0,
                               B + (long long)oend * ldb, ldb, rest, p, oend - o0, 0, stream);
      if (rc) return rc;
    }
  }
  return IPM_OK;
}
