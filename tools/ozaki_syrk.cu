// EXPERIMENT RECORD (tools/, not on the product path; the kernel that grew out of it is interiorpoint-gpu_b200/csrc/hess_i8.cu,
// which has since replaced the issue loop below -- see its header): the Hessian contraction H = C' diag(w) C in FP64 accuracy on the INT8
// tensor pipe of sm_100a (tcgen05.mma.kind::i8, accumulators in TMEM), by error-free slicing along the contraction index
// (Ozaki scheme).  tools/ozaki_probe.py measured the idea with a library INT8 GEMM; this is the hand-written kernel.
//
//   X = diag(sqrt w) C  (K x n, K = m rows).  Column i is scaled by the power of two sigma_i = 2^(E_i + 2), where
//   2^E_i <= max_k |X_ki| < 2^(E_i + 1), and cut into s signed digits q_t in [-64, 64] (round to nearest, base 128):
//       X_ki = sigma_i sum_{t < s} q_t[k, i] 128^-(t+1)  +  O(sigma_i 128^-s)
//       H_ij = sigma_i sigma_j sum_{d < s} 128^-(d+2) sum_{t + u = d} (Q_t' Q_u)_ij
//   Every Q_t' Q_u is an exact INT8 x INT8 -> INT32 product (|q q'| <= 2^12, K <= 2^17 terms, up to 8 pairs per d:
//   < 2^31).  The pairs of one diagonal d = t + u share an accumulator, so an output tile of 128 x 64 holds
//   s <= 8 accumulators x 64 columns = all 512 TMEM columns of the SM.
//
// Kernels:
//   colmax_kernel      amax_i = max_k sqrt(w_k) |C_ki|                       (HBM: reads C once)
//   slice_kernel       Q[t][i][k] (int8, K-major = transposed), sigma_i     (HBM: reads C once, writes s bytes / entry)
//   ozaki_syrk_kernel  persistent, 1 CTA / SM, warp roles: TMA producer / MMA issuer / 4 epilogue warps.
//                      v1 (32-byte k-blocks, one 128x64x32 MMA per slice pair) was exact but ran at 1.5 POP/s: every MMA
//                      re-read its 4 KB A operand from shared memory (6 KB per 32 clk > the 128 B/clk of the SM) and the
//                      TMA wrote 32-byte rows at ~1 row / clk.  v2:
//                        * 128-byte rows (SWIZZLE_128B), k-chunks of 128 bytes = 4 MMA k-steps (descriptor start + 32 B);
//                        * the B side of a chunk -- all s slices of the 64 rows, [s][64][128 B] = one contiguous K-major
//                          operand of 64 s rows -- sits in one of 2 stages; the A slices stream through a ring of 16 KB
//                          slots;
//                        * slice t of A meets slices 0..s-1-t of B, whose accumulators (diagonals t..s-1) are adjacent
//                          TMEM columns: ONE wide MMA (N = 64 (s - t), split at 256) instead of s - t narrow ones, so A
//                          is read once per wide MMA: 12 instead of 36 A reads per k-step at s = 8;
//                      epilogue: tcgen05.ld, Horner in FP64, scale by sigma_i sigma_j, store.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

namespace {
constexpr int TM = 128, TN = 64, KC = 128, KSTEP = 32, SMAX = 8;    // k-chunk of 128 bytes = 4 MMA k-steps
constexpr int NA = 6, NB = 2;                                         // A ring slots, B stages
constexpr int A_SLOT = TM * KC, B_SLICE = TN * KC, B_STAGE = SMAX * B_SLICE;  // 16 KB, 8 KB, 64 KB
constexpr int SMEM_BYTES = NA * A_SLOT + NB * B_STAGE + 1024 + 256;   // + alignment slack + barriers
constexpr int THREADS = 224;  // warps: 0 A producer, 1 MMA issuer, 2-5 epilogue, 6 B producer

// ---------------------------------------------------------------- slicing
__global__ void __launch_bounds__(128) colmax_kernel(const double* __restrict__ C, long long ldc, int m, int n,
                                                     const double* __restrict__ w, unsigned long long* __restrict__ amax) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  const int k0 = blockIdx.y * 128, k1 = min(m, k0 + 128);
  if (i >= n) return;
  double mx = 0.0;
#pragma unroll 8
  for (int k = k0; k < k1; ++k) mx = fmax(mx, sqrt(w[k]) * fabs(C[(long long)k * ldc + i]));
  if (mx > 0.0) atomicMax(amax + i, (unsigned long long)__double_as_longlong(mx));  // non-negative doubles order as integers
}

constexpr int SL_COLS = 32, SL_K = 128, SL_PITCH = SL_K + 4;
__global__ void __launch_bounds__(256) slice_kernel(const double* __restrict__ C, long long ldc, int m, int n,
                                                    const double* __restrict__ w,
                                                    const unsigned long long* __restrict__ amax, int s,
                                                    int8_t* __restrict__ Q, long long n_pad, long long k_pad,
                                                    double* __restrict__ sigma) {
  __shared__ __align__(16) int8_t tile[SMAX * SL_COLS * SL_PITCH];
  const int ci = threadIdx.x & 31, kr = threadIdx.x >> 5;
  const int i = blockIdx.x * SL_COLS + ci, k0 = blockIdx.y * SL_K;
  double inv = 0.0;
  if (i < n) {
    const unsigned long long ef = amax[i] >> 52;  // exponent field of amax (sign bit is 0)
    if (ef != 0) {
      inv = __longlong_as_double((long long)(2044ull - ef) << 52);          // 2^-(E+2)
      if (blockIdx.y == 0 && kr == 0) sigma[i] = __longlong_as_double((long long)(ef + 2ull) << 52);
    } else if (blockIdx.y == 0 && kr == 0) {
      sigma[i] = 0.0;
    }
  }
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: (r + magic) - magic = rint(r), low mantissa bits = the integer
  for (int kk = kr; kk < SL_K; kk += 8) {
    const int k = k0 + kk;
    double r = 0.0;
    if (i < n && k < m) r = sqrt(w[k]) * C[(long long)k * ldc + i] * inv;  // |r| <= 1/2
    for (int t = 0; t < s; ++t) {
      r *= 128.0;
      const double tmp = r + magic;
      const double q = tmp - magic;
      r -= q;                                                                // exact; stays in [-1/2, 1/2]
      tile[(t * SL_COLS + ci) * SL_PITCH + kk] = (int8_t)(int)__double2loint(tmp);
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = warp; row < s * SL_COLS; row += 8) {
    const int t = row / SL_COLS, c = row % SL_COLS;
    const long long gi = (long long)blockIdx.x * SL_COLS + c;
    if (gi >= n) continue;
    const uint32_t v = *reinterpret_cast<const uint32_t*>(&tile[row * SL_PITCH + lane * 4]);
    *reinterpret_cast<uint32_t*>(Q + ((long long)t * n_pad + gi) * k_pad + k0 + lane * 4) = v;
  }
}

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a wrong descriptor must end in a reported failure, not in a hung box
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* fail, int code) {
  for (long long spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
    if ((spins & 1023) == 1023) {
      if (*(volatile int*)fail != 0) return false;
      if (spins > (1ll << 22)) {
        atomicCAS(fail, 0, code);
        return false;
      }
    }
  }
}
__device__ __forceinline__ bool mbar_wait_timed(uint32_t bar, uint32_t parity, int* fail, int code, long long& acc) {
  const long long t0 = clock64();
  const bool ok = mbar_wait(bar, parity, fail, code);
  acc += clock64() - t0;
  return ok;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct Args {
  const int2* tiles;  // (row block of 128, column block of 64)
  int ntiles, nkc, s, n;
  const double* sigma;
  double* H;
  long long ldh;
  uint64_t desc_template;  // UMMA shared-memory descriptor without the start address
  uint32_t idesc;          // instruction descriptor without N
  int* fail;        // 0 = ok; otherwise the code of the wait that timed out
  int* dbg;         // raw accumulators [s][128][64] of tile `dbg_tile`, or NULL
  int dbg_tile;
  long long* prof;  // per CTA: cycles the MMA thread waited for [A slot, B stage, TMEM], its total, producer waits [A, B]
};

__device__ __forceinline__ uint64_t smem_desc(uint64_t tmpl, uint32_t addr) {
  return tmpl | (uint64_t)((addr & 0x3FFFFu) >> 4);
}

__global__ void __launch_bounds__(THREADS, 1)
ozaki_syrk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smB = base, smA = base + NB * B_STAGE;
  const uint32_t bars = smA + NA * A_SLOT;
  const uint32_t a_full = bars, a_empty = bars + 8 * NA, b_full = bars + 16 * NA, b_empty = b_full + 8 * NB,
                 tfull = b_empty + 8 * NB, tempty = tfull + 8, tptr = tfull + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = a.s;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) {
      mbar_init(a_full + 8 * i, 1);
      mbar_init(a_empty + 8 * i, 1);
    }
    for (int i = 0; i < NB; ++i) {
      mbar_init(b_full + 8 * i, 1);
      mbar_init(b_empty + 8 * i, 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tptr));

  if (warp == 0) {  // A slices: one 16 KB slot per (chunk, slice), in the order the MMA thread consumes them
    if (lane == 0) {
      long long ia = 0, wa = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < a.ntiles && ok; tile += gridDim.x) {
        const int2 tl = a.tiles[tile];
        for (int kc = 0; kc < a.nkc && ok; ++kc) {
          for (int t = 0; t < s && ok; ++t, ++ia) {
            const int sl = (int)(ia % NA);
            if (ia >= NA) ok = mbar_wait_timed(a_empty + 8 * sl, (uint32_t)((ia / NA) - 1) & 1u, a.fail, 5, wa);
            if (!ok) break;
            mbar_expect_tx(a_full + 8 * sl, A_SLOT);
            tma_load_3d(smA + sl * A_SLOT, &tmA, kc * KC, tl.x * TM, t, a_full + 8 * sl);
          }
        }
      }
      if (a.prof) a.prof[blockIdx.x * 8 + 4] = wa;
    }
  } else if (warp == 6) {  // B side of a chunk: its own thread, so that a stage is refilled the moment it is released
    if (lane == 0) {
      long long ib = 0, wb = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < a.ntiles && ok; tile += gridDim.x) {
        const int2 tl = a.tiles[tile];
        for (int kc = 0; kc < a.nkc && ok; ++kc, ++ib) {
          const int st = (int)(ib % NB);
          if (ib >= NB) ok = mbar_wait_timed(b_empty + 8 * st, (uint32_t)((ib / NB) - 1) & 1u, a.fail, 1, wb);
          if (!ok) break;
          mbar_expect_tx(b_full + 8 * st, (uint32_t)s * B_SLICE);
          tma_load_3d(smB + st * B_STAGE, &tmB, kc * KC, tl.y * TN, 0, b_full + 8 * st);
        }
      }
      if (a.prof) a.prof[blockIdx.x * 8 + 5] = wb;
    }
  } else if (warp == 1) {
    if (lane == 0) {
      long long ia = 0, ib = 0;
      long long wa = 0, wb = 0, wt = 0;
      const long long tstart = clock64();
      int tcount = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < a.ntiles && ok; tile += gridDim.x, ++tcount) {
        if (tcount > 0) {
          ok = mbar_wait_timed(tempty, (uint32_t)(tcount - 1) & 1u, a.fail, 2, wt);
          if (!ok) break;
          tc_fence_after();
        }
        for (int kc = 0; kc < a.nkc && ok; ++kc, ++ib) {
          const int st = (int)(ib % NB);
          ok = mbar_wait_timed(b_full + 8 * st, (uint32_t)(ib / NB) & 1u, a.fail, 3, wb);
          if (!ok) break;
          const uint32_t sb = smB + st * B_STAGE;
          for (int t = 0; t < s && ok; ++t, ++ia) {
            const int sl = (int)(ia % NA);
            ok = mbar_wait_timed(a_full + 8 * sl, (uint32_t)(ia / NA) & 1u, a.fail, 6, wa);
            if (!ok) break;
            tc_fence_after();
            const uint32_t sa = smA + sl * A_SLOT;
            const int ncols = TN * (s - t);  // B rows = slices 0 .. s-1-t; accumulators of the diagonals t .. s-1
#pragma unroll
            for (int j = 0; j < KC / KSTEP; ++j) {
              const uint64_t da = smem_desc(a.desc_template, sa + j * KSTEP);
              for (int n0 = 0; n0 < ncols; n0 += 256) {
                const int nn = min(256, ncols - n0);
                const uint64_t db = smem_desc(a.desc_template, sb + n0 * KC + j * KSTEP);
                umma_i8(tmem + t * TN + n0, da, db, a.idesc | ((uint32_t)(nn >> 3) << 17),
                        (kc > 0 || t > 0 || j > 0) ? 1u : 0u);
              }
            }
            umma_commit(a_empty + 8 * sl);  // frees the A slot once these MMAs have read it
          }
          if (ok) umma_commit(b_empty + 8 * st);
        }
        if (ok) umma_commit(tfull);
      }
      if (a.prof) {
        a.prof[blockIdx.x * 8 + 0] = wa;
        a.prof[blockIdx.x * 8 + 1] = wb;
        a.prof[blockIdx.x * 8 + 2] = wt;
        a.prof[blockIdx.x * 8 + 3] = clock64() - tstart;
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    int tcount = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < a.ntiles && ok; tile += gridDim.x, ++tcount) {
      const int2 tl = a.tiles[tile];
      ok = mbar_wait(tfull, (uint32_t)tcount & 1u, a.fail, 4);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      tc_fence_after();
      const long long i = (long long)tl.x * TM + row;
      const double si = i < a.n ? a.sigma[i] * 0x1p-14 : 0.0;
      for (int c0 = 0; c0 < TN; c0 += 8) {
        int v[SMAX][8];
#pragma unroll
        for (int d = 0; d < SMAX; ++d)
          if (d < s) tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + d * TN + c0, v[d]);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < SMAX; ++d)  // the loaded registers are valid only after the wait: pin their uses behind it
          asm volatile("" : "+r"(v[d][0]), "+r"(v[d][1]), "+r"(v[d][2]), "+r"(v[d][3]), "+r"(v[d][4]), "+r"(v[d][5]),
                            "+r"(v[d][6]), "+r"(v[d][7]));
        if (a.dbg && tile == a.dbg_tile) {
#pragma unroll
          for (int d = 0; d < SMAX; ++d)
            if (d < s)
              for (int c = 0; c < 8; ++c) a.dbg[(d * TM + row) * TN + c0 + c] = v[d][c];
        }
        const long long j0 = (long long)tl.y * TN + c0;
        if (i < a.n) {
          double out[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            double r = 0.0;
#pragma unroll
            for (int d = SMAX - 1; d >= 0; --d)
              if (d < s) r = fma(r, 0x1p-7, (double)v[d][c]);
            out[c] = r;
          }
          double* dst = a.H + i * a.ldh + j0;
          if (j0 + 8 <= a.n) {
            const double2 sj0 = *reinterpret_cast<const double2*>(a.sigma + j0);
            const double2 sj1 = *reinterpret_cast<const double2*>(a.sigma + j0 + 2);
            const double2 sj2 = *reinterpret_cast<const double2*>(a.sigma + j0 + 4);
            const double2 sj3 = *reinterpret_cast<const double2*>(a.sigma + j0 + 6);
            reinterpret_cast<double2*>(dst)[0] = make_double2(out[0] * si * sj0.x, out[1] * si * sj0.y);
            reinterpret_cast<double2*>(dst)[1] = make_double2(out[2] * si * sj1.x, out[3] * si * sj1.y);
            reinterpret_cast<double2*>(dst)[2] = make_double2(out[4] * si * sj2.x, out[5] * si * sj2.y);
            reinterpret_cast<double2*>(dst)[3] = make_double2(out[6] * si * sj3.x, out[7] * si * sj3.y);
          } else {
            for (int c = 0; c < 8; ++c)
              if (j0 + c < a.n) dst[c] = out[c] * si * a.sigma[j0 + c];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
int slice_map(CUtensorMap* tm, const int8_t* Q, long long n_pad, long long k_pad, int s, int box_rows, int box_slices) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1;
  cuuint64_t gdim[3] = {(cuuint64_t)k_pad, (cuuint64_t)n_pad, (cuuint64_t)s};
  cuuint64_t gstride[2] = {(cuuint64_t)k_pad, (cuuint64_t)(n_pad * k_pad)};
  cuuint32_t box[3] = {KC, (cuuint32_t)box_rows, (cuuint32_t)box_slices};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)Q, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}
}  // namespace

static long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

extern "C" long long ozaki_n_pad(int n) { return round_up(n, 128); }
extern "C" long long ozaki_k_pad(int m) { return round_up(m, 128); }
extern "C" int ozaki_tile_count(int n) {
  const int rb = (int)(ozaki_n_pad(n) / TM);
  int cnt = 0;
  for (int bi = 0; bi < rb; ++bi) cnt += 2 * rb - 2 * bi;
  return cnt;
}
// tiles of the upper triangle, panels of `panel` column blocks so that one wave of CTAs shares few A and B row blocks
extern "C" int ozaki_tile_list(int n, int panel, int* out /* 2 ints per tile */) {
  const int rb = (int)(ozaki_n_pad(n) / TM), cb = 2 * rb;
  int cnt = 0;
  for (int p0 = 0; p0 < cb; p0 += panel)
    for (int bi = 0; bi < rb; ++bi)
      for (int bj = p0; bj < p0 + panel && bj < cb; ++bj)
        if (bj >= 2 * bi) {
          out[2 * cnt] = bi;
          out[2 * cnt + 1] = bj;
          ++cnt;
        }
  return cnt;
}

// amax: n_pad u64 (zeroed here); Q: s * n_pad * k_pad bytes whose padding is zero (zero it once); sigma: n_pad doubles
extern "C" int ozaki_slice_f64(const double* C, long long ldc, int m, int n, const double* w, int s,
                               unsigned long long* amax, int8_t* Q, double* sigma, void* stream) {
  if (s < 1 || s > SMAX) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n_pad = ozaki_n_pad(n), k_pad = ozaki_k_pad(m);
  cudaMemsetAsync(amax, 0, n_pad * 8, st);
  colmax_kernel<<<dim3((n + 127) / 128, (m + 127) / 128), 128, 0, st>>>(C, ldc, m, n, w, amax);
  slice_kernel<<<dim3((unsigned)(n_pad / SL_COLS), (unsigned)(k_pad / SL_K)), 256, 0, st>>>(C, ldc, m, n, w, amax, s, Q, n_pad,
                                                                                          k_pad, sigma);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int ozaki_syrk_i8(const int8_t* Q, int m, int n, int s, const double* sigma, const int* tiles, int ntiles,
                             double* H, long long ldh, unsigned long long desc_template, int* fail, int* dbg, int dbg_tile,
                             int max_ctas, long long* prof, void* stream) {
  if (s < 1 || s > SMAX) return -1;
  const long long n_pad = ozaki_n_pad(n), k_pad = ozaki_k_pad(m);
  CUtensorMap tmA, tmB;
  int rc = slice_map(&tmA, Q, n_pad, k_pad, s, TM, 1);
  if (rc) return rc;
  rc = slice_map(&tmB, Q, n_pad, k_pad, s, TN, s);
  if (rc) return rc;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(ozaki_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
      return -4;
    attr = true;
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = ntiles < sms ? ntiles : sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  Args a;
  a.tiles = reinterpret_cast<const int2*>(tiles);
  a.ntiles = ntiles;
  a.nkc = (int)(k_pad / KC);
  a.s = s;
  a.n = n;
  a.sigma = sigma;
  a.H = H;
  a.ldh = ldh;
  a.desc_template = desc_template;
  a.idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TM >> 4) << 24);  // S32 accumulate, signed int8 A and B, K-major
  a.fail = fail;
  a.dbg = dbg;
  a.dbg_tile = dbg_tile;
  a.prof = prof;
  ozaki_syrk_kernel<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tmA, tmB, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -5;
}
