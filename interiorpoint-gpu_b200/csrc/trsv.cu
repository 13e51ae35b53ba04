// Single right-hand-side triangular solves with the Cholesky factor (FP64, upper / row-major, H = U^T U):
//     trans = 1:  U^T y = b   (forward, top -> bottom)        trans = 0:  U x = b   (backward, bottom -> top)
// Replaces the two dtrsv-like halves of scipy.linalg.cho_solve on the Newton path (NewtonSolver.py:303-313,
// NewtonSolverInfeasibleStart.py:455-490).
//
// HBM-bound by construction (U is read exactly once, 4 n^2 bytes), latency-bound in practice: the n/128
// diagonal blocks form a dependent chain.  ONE persistent launch per solve: CTA c owns the 128-entry solution
// blocks c, c + G, ... (G = grid <= #SMs, so every CTA is resident).  A block is solved LEFT-looking: its owner
// streams the 128x128 tiles of U that couple it to the already-solved blocks (register-prefetched one tile
// ahead of the flag it waits on), accumulates in registers, then does the 128x128 substitution in shared memory
// and publishes the block through a release store on a per-block flag.  The chain per block is therefore
// [flag -> one tile FMA -> 128-step substitution -> flag]; everything else streams behind it.
#include <atomic>

#include "common.cuh"

using namespace ipm;

namespace {

constexpr int NB = 128;
constexpr int TS_LD = NB + 1;       // padded: column access (backward solve) is conflict-free
constexpr int TV_THREADS = 256;
constexpr int MAX_BLOCKS = 4096;    // n <= 524288

__device__ unsigned int g_trsv_flag[MAX_BLOCKS];  // block j is published when its flag equals the call's epoch

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// 128x128 (nb x nb) substitution with the diagonal block held in shared memory (S, row-major, ld = TS_LD); xs is
// the right-hand side in, the solution out.  32-wide sub-blocks: warp 0 runs the 32-step dependent chain on
// registers + shuffles (the lane's coefficients are pre-scaled by the reciprocal pivot so that the chain is
// SHFL -> DFMA only), then all threads apply the solved sub-block to the entries still to be solved.
__device__ __forceinline__ void diag_block_solve(const double* __restrict__ S, double* __restrict__ xs, int nb,
                                                 int trans) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nsb = (nb + 31) >> 5;
  for (int s = 0; s < nsb; ++s) {
    const int sb = trans ? s : nsb - 1 - s;
    const int base = sb * 32;
    const int bl = min(32, nb - base);
    if (warp == 0) {
      const bool act = lane < bl;
      const double rdg = act ? 1.0 / S[(base + lane) * TS_LD + base + lane] : 1.0;
      double v = act ? xs[base + lane] * rdg : 0.0;
      double u[32];
      if (trans) {
#pragma unroll
        for (int l = 0; l < 32; ++l)
          u[l] = (act && l < lane) ? S[(base + l) * TS_LD + base + lane] * rdg : 0.0;
#pragma unroll
        for (int l = 0; l < 32; ++l) {
          const double yl = __shfl_sync(0xffffffffu, v, l);
          v = fma(-u[l], yl, v);  // u[l] == 0 for lanes <= l
        }
      } else {
#pragma unroll
        for (int l = 0; l < 32; ++l)
          u[l] = (act && l > lane && l < bl) ? S[(base + lane) * TS_LD + base + l] * rdg : 0.0;
#pragma unroll
        for (int l = 31; l >= 0; --l) {
          const double xl = __shfl_sync(0xffffffffu, v, l);
          v = fma(-u[l], xl, v);
        }
      }
      if (act) xs[base + lane] = v;
    }
    __syncthreads();
    if (trans) {
      // entries below the sub-block: two threads per entry (even / odd l)
      const int cidx = base + bl + (tid >> 1), h = tid & 1;
      double a = 0.0;
      if (cidx < nb) {
#pragma unroll 8
        for (int l = h; l < bl; l += 2) a = fma(S[(base + l) * TS_LD + cidx], xs[base + l], a);
      }
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      if (cidx < nb && h == 0) xs[cidx] -= a;  // disjoint from the entries read above
    } else {
      const int ridx = tid >> 1, h = tid & 1;
      double a = 0.0;
      if (ridx < base) {
#pragma unroll 8
        for (int l = h; l < bl; l += 2) a = fma(S[ridx * TS_LD + base + l], xs[base + l], a);
      }
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      if (ridx < base && h == 0) xs[ridx] -= a;
    }
    __syncthreads();
  }
}

// Two adjacent doubles of a tile row (zero beyond `ncol`).  VEC: 16-byte aligned rows.
template <bool VEC>
__device__ __forceinline__ double2 load_pair(const double* __restrict__ p, int c, int ncol) {
  if (VEC) {
    if (c + 1 < ncol) return __ldg(reinterpret_cast<const double2*>(p));
    return make_double2(c < ncol ? __ldg(p) : 0.0, 0.0);
  }
  return make_double2(c < ncol ? __ldg(p) : 0.0, c + 1 < ncol ? __ldg(p + 1) : 0.0);
}

template <bool VEC>
__device__ __forceinline__ void load_diag(const double* __restrict__ Ukk, long long ld, int nb, double* __restrict__ S) {
  const int tid = threadIdx.x;
#pragma unroll 8
  for (int q = 0; q < 32; ++q) {
    const int idx = tid + TV_THREADS * q;  // pair index: row = idx / 64, col = 2 * (idx % 64)
    const int r = idx >> 6, c = (idx & 63) * 2;
    if (r < nb && c + 1 >= r) {
      const double2 v = load_pair<VEC>(Ukk + (long long)r * ld + c, c, nb);
      S[r * TS_LD + c] = v.x;
      S[r * TS_LD + c + 1] = v.y;
    }
  }
}

__device__ __forceinline__ void wait_block(const unsigned* flag, unsigned epoch, unsigned int* fault) {
  if (threadIdx.x == 0) spin_wait([&] { return ld_acquire_u32(flag) == epoch; }, fault, IPM_FAULT_TRSV);
  __syncthreads();
}

__device__ __forceinline__ void publish_block(double* __restrict__ b, int k0, int nb, const double* __restrict__ xs,
                                              unsigned* flag, unsigned epoch) {
  if ((int)threadIdx.x < nb) __stcg(b + k0 + threadIdx.x, xs[threadIdx.x]);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    st_release_u32(flag, epoch);
  }
}

// ---------------------------------------------------------------------------------------------------------
// forward:  y_j = U_jj^{-T} ( b_j - sum_{i<j} U_ij^T y_i ).   Tile U[i-block rows][j-block cols]; thread
// (rg = tid / 64, cp = tid % 64) owns columns 2cp, 2cp+1 and the 32 tile rows rg*32 .. rg*32+31.
// ---------------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(TV_THREADS, 1)
trsv_forward_kernel(const double* __restrict__ U, long long ld, int n, double* __restrict__ b, unsigned epoch,
                    unsigned int* __restrict__ fault) {
  extern __shared__ double S[];  // NB x TS_LD
  __shared__ double xs[NB];
  __shared__ double ys[2][NB];
  __shared__ double part[4][NB];
  const int tid = threadIdx.x;
  const int cp = tid & 63, rg = tid >> 6, c = 2 * cp;
  const int nblk = (n + NB - 1) / NB;
  for (int j = blockIdx.x; j < nblk; j += gridDim.x) {
    const int k0 = j * NB, nb = min(NB, n - k0);
    load_diag<VEC>(U + (long long)k0 * ld + k0, ld, nb, S);
    double a0 = 0.0, a1 = 0.0;
    double2 m[32];
    const double* col = U + k0 + c;
    if (j > 0) {
#pragma unroll
      for (int q = 0; q < 32; ++q) m[q] = load_pair<VEC>(col + (long long)(rg * 32 + q) * ld, c, nb);
    }
    for (int i = 0; i < j; ++i) {
      wait_block(g_trsv_flag + i, epoch, fault);
      double* yb = ys[i & 1];
      if (tid < NB) yb[tid] = __ldcg(b + i * NB + tid);
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const double y = yb[rg * 32 + q];
        a0 = fma(m[q].x, y, a0);
        a1 = fma(m[q].y, y, a1);
      }
      if (i + 1 < j) {
        const double* nxt = col + (long long)((i + 1) * NB + rg * 32) * ld;
#pragma unroll
        for (int q = 0; q < 32; ++q) m[q] = load_pair<VEC>(nxt + (long long)q * ld, c, nb);
      }
    }
    part[rg][c] = a0;
    part[rg][c + 1] = a1;
    __syncthreads();  // also orders load_diag's stores before the substitution
    if (tid < nb) xs[tid] = b[k0 + tid] - ((part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]));
    __syncthreads();
    diag_block_solve(S, xs, nb, 1);
    publish_block(b, k0, nb, xs, g_trsv_flag + j, epoch);
    __syncthreads();  // S / xs are reused by this CTA's next block
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward:  x_k = U_kk^{-1} ( y_k - sum_{j>k} U_kj x_j ).   Tile U[k-block rows][j-block cols]; warp w owns the
// 16 tile rows w*16 .. w*16+15, lane owns columns {2 lane, 2 lane + 1, 64 + 2 lane, 64 + 2 lane + 1}.
// ---------------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(TV_THREADS, 1)
trsv_backward_kernel(const double* __restrict__ U, long long ld, int n, double* __restrict__ b, unsigned epoch,
                     unsigned int* __restrict__ fault) {
  extern __shared__ double S[];
  __shared__ double xs[NB];
  __shared__ double xj[2][NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (n + NB - 1) / NB;
  for (int kk = blockIdx.x; kk < nblk; kk += gridDim.x) {
    const int kb = nblk - 1 - kk;
    const int k0 = kb * NB, nb = min(NB, n - k0);
    load_diag<VEC>(U + (long long)k0 * ld + k0, ld, nb, S);
    double acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.0;
    double2 m0[16], m1[16];
    // rows of a non-final row block are always complete (nb == 128 whenever kb < nblk - 1)
    const double* rowp = U + (long long)(k0 + warp * 16) * ld + 2 * lane;
    if (kb < nblk - 1) {
      const int j = nblk - 1, ncol = min(NB, n - j * NB);
      const double* t = rowp + (long long)j * NB;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        m0[q] = load_pair<VEC>(t + (long long)q * ld, 2 * lane, ncol);
        m1[q] = load_pair<VEC>(t + (long long)q * ld + 64, 64 + 2 * lane, ncol);
      }
    }
    for (int j = nblk - 1; j > kb; --j) {
      wait_block(g_trsv_flag + j, epoch, fault);
      double* xb = xj[j & 1];
      if (tid < NB) xb[tid] = (j * NB + tid < n) ? __ldcg(b + j * NB + tid) : 0.0;
      __syncthreads();
      const double x0 = xb[2 * lane], x1 = xb[2 * lane + 1], x2 = xb[64 + 2 * lane], x3 = xb[65 + 2 * lane];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        double a = acc[q];
        a = fma(m0[q].x, x0, a);
        a = fma(m0[q].y, x1, a);
        a = fma(m1[q].x, x2, a);
        a = fma(m1[q].y, x3, a);
        acc[q] = a;
      }
      if (j - 1 > kb) {
        const double* t = rowp + (long long)(j - 1) * NB;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          m0[q] = load_pair<VEC>(t + (long long)q * ld, 2 * lane, NB);
          m1[q] = load_pair<VEC>(t + (long long)q * ld + 64, 64 + 2 * lane, NB);
        }
      }
    }
    // butterfly transpose-reduce: 16 row sums over 32 lanes in 16 shuffles; lane L ends with row
    // ((L&1)<<3)|((L&2)<<1)|((L&4)>>1)|((L&8)>>3)
#pragma unroll
    for (int bit = 0; bit < 4; ++bit) {
      const int half = 8 >> bit;
      const bool up = (lane >> bit) & 1;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q < half) {
          const double keep = up ? acc[q + half] : acc[q];
          const double send = up ? acc[q] : acc[q + half];
          acc[q] = keep + __shfl_xor_sync(0xffffffffu, send, 1 << bit);
        }
      }
    }
    acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 16);
    __syncthreads();  // previous block's xs readers are done; load_diag stores ordered before the substitution
    if (lane < 16) {
      const int q = ((lane & 1) << 3) | ((lane & 2) << 1) | ((lane & 4) >> 1) | ((lane & 8) >> 3);
      const int r = warp * 16 + q;
      if (r < nb) xs[r] = b[k0 + r] - acc[0];
    }
    __syncthreads();
    diag_block_solve(S, xs, nb, 0);
    publish_block(b, k0, nb, xs, g_trsv_flag + kb, epoch);
    __syncthreads();
  }
}

std::atomic<unsigned> g_epoch{0};

}  // namespace

// b is overwritten with the solution.  `ws` is unused (kept for ABI stability; may be null).
// Not re-entrant per device: two concurrent solves on different streams of one device would share the flags.
extern "C" int ipm_trsv_upper_f64(const double* U, int ld, int n, double* b, int trans, double* ws, void* stream) {
  (void)ws;
  if (!U || !b || n < 0 || ld < n) return IPM_ERR_ARG;
  if (n == 0) return IPM_OK;
  const int nblk = ceil_div(n, NB);
  if (nblk > MAX_BLOCKS) return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  static int num_sms[16] = {0};
  int dev = 0;
  IPM_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return IPM_ERR_ARG;
  const int smem = NB * TS_LD * 8;
  if (!num_sms[dev]) {
    int v = 0;
    IPM_CUDA_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    IPM_CUDA_CHECK(cudaFuncSetAttribute(trsv_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    IPM_CUDA_CHECK(cudaFuncSetAttribute(trsv_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    IPM_CUDA_CHECK(cudaFuncSetAttribute(trsv_backward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    IPM_CUDA_CHECK(cudaFuncSetAttribute(trsv_backward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    num_sms[dev] = v;
  }
  const int grid = nblk < num_sms[dev] ? nblk : num_sms[dev];
  unsigned epoch = ++g_epoch;
  if (epoch == 0) epoch = ++g_epoch;  // never 0 (the flags' initial value)
  const bool vec = !(ld & 1) && !(((uintptr_t)U) & 15);
  unsigned int* fault = ipm_internal_fault_word();
  if (trans) {
    if (vec) trsv_forward_kernel<true><<<grid, TV_THREADS, smem, st>>>(U, ld, n, b, epoch, fault);
    else trsv_forward_kernel<false><<<grid, TV_THREADS, smem, st>>>(U, ld, n, b, epoch, fault);
  } else {
    if (vec) trsv_backward_kernel<true><<<grid, TV_THREADS, smem, st>>>(U, ld, n, b, epoch, fault);
    else trsv_backward_kernel<false><<<grid, TV_THREADS, smem, st>>>(U, ld, n, b, epoch, fault);
  }
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}
