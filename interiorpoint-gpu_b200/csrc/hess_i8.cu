// Barrier Hessian H = beta H + C' diag(w) C (w >= 0) in FP64 accuracy on the INT8 tensor pipe of sm_100a
// (tcgen05.mma.kind::i8, accumulators in TMEM) -- the same contraction as ipm_gemm_tn_f64(upper = 1)
// (FunctionManager.py:301-312, 564-576, 801-813), 2.5x faster at the cfg-2 shape (profiles/hess_i8_sizes_r02.jsonl).
//
// Error-free slicing along the contraction index (Ozaki scheme).  X = diag(sqrt w) C (K x n, K = m rows).  Column i is
// scaled by the power of two sigma_i = 2^(E_i + 2), 2^E_i <= max_k |X_ki| < 2^(E_i + 1), and cut into s signed digits
// q_t in [-64, 64] (round to nearest, base 128):
//     X_ki = sigma_i sum_{t < s} q_t[k, i] 128^-(t+1)  +  O(sigma_i 128^-s)
//     H_ij = sigma_i sigma_j sum_{d < s} 128^-(d+2) sum_{t + u = d} (Q_t' Q_u)_ij
// Every Q_t' Q_u is an EXACT INT8 x INT8 -> INT32 product (|q q'| <= 2^12, K < 2^16 terms, <= 8 pairs per diagonal
// d = t + u: below 2^31 in the worst case), so the only roundings are the slicing itself and the FP64 recombination of
// s integers per entry.  s = 8 (56 bits) reproduces the DMMA kernel to 2e-15 of sum_k |x_ki x_kj| on weights spanning 20
// decades.
//
// Kernels
//   colmax_kernel   amax_i = max_k sqrt(w_k) |C_ki|                        HBM: reads C once
//   slice_kernel    Q[t][i][k] (int8, K-major = transposed) and sigma_i    HBM: reads C once, writes s bytes per entry
//   syrk_kernel     persistent, 1 CTA per SM, warp roles: A producer / MMA issuer / 4 epilogue warps / B producer.
//     * output tile 128 x 64; one accumulator (64 TMEM columns) per diagonal d: s <= 8 accumulators = all 512 columns;
//     * k-chunks of 128 bytes (rows of 128 B, SWIZZLE_128B in the tensor map and in the UMMA descriptors; the 4 MMA
//       k-steps of a chunk advance the descriptor start address by 32 B);
//     * the B side of a chunk -- all s slices of the 64 rows, [s][64][128 B] = ONE contiguous K-major operand of 64 s
//       rows -- sits in one of 2 stages; the A slices stream through a ring of 6 slots of 16 KB;
//     * slice t of A meets slices 0..s-1-t of B, whose accumulators (diagonals t..s-1) are adjacent TMEM columns: one
//       wide MMA (N = 64 (s - t), split at 256, the second piece reusing the A operand from the collector buffer)
//       instead of s - t narrow ones -- 12 MMAs and 8 A-operand reads per k-step at s = 8 instead of 36;
//     * the MMA warp runs its loop with warp-uniform control flow and compile-time descriptor offsets (issue_slice):
//       one MMA costs a couple of uniform adds.  The first version issued from `if (lane == 0)`, where every operand
//       is a per-thread value that has to be moved to the uniform datapath: ~200 clocks per MMA whatever its shape
//       (tools/umma_probe.cu), tensor pipe 49 % active, 17.9 ms.  Now: pipe 86 % active, 11.6 ms at the cfg-2 shape
//       (profiles/hess_i8_syrk_ncu_r02.csv) = 3.4 POP/s, 0.92 of the library INT8 GEMM rate on the same box;
//     * epilogue: tcgen05.ld of the s accumulators, Horner in FP64, scale by sigma_i sigma_j, [+ beta H], store -- or,
//       row-sharded (SCATTER), the half tile goes into the inbox of the rank that owns its 128 x 128 tile.
// Every wait is bounded by the library's watchdog word (common.cuh).
#include <cuda.h>
#include <stdint.h>

#include <type_traits>

#include "common.cuh"
#include "tensormap.cuh"

using namespace ipm;

namespace {
constexpr int TM = 128, TN = 64, KC = 128, KSTEP = 32, SMAX = 8;     // k-chunk of 128 bytes = 4 MMA k-steps
constexpr int NA = 6, NB = 2;                                          // A ring slots, B stages
constexpr int A_SLOT = TM * KC, B_SLICE = TN * KC, B_STAGE = SMAX * B_SLICE;  // 16 KB, 8 KB, 64 KB
constexpr int SMEM_BYTES = NA * A_SLOT + NB * B_STAGE + 1024 + 256;    // + alignment slack + barriers
constexpr int kMaxRows = 65408;  // 8 pairs x K x 2^12 < 2^31 for the padded K
constexpr int THREADS = 224;  // warps: 0 A producer, 1 MMA issuer, 2-5 epilogue, 6 B producer
// UMMA shared-memory descriptor without the start address: K-major, 128-byte swizzle, 8-row groups 1024 B apart
// (LBO>>4 at bit 16 -- ignored for swizzled K-major --, SBO>>4 at bit 32, version 1 at bit 46, SWIZZLE_128B = 2 at 61)
constexpr uint64_t DESC = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
// instruction descriptor without N: S32 accumulator, signed 8-bit A and B, both K-major, M = 128
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TM >> 4) << 24);

// ---------------------------------------------------------------- slicing
__global__ void __launch_bounds__(128) colmax_kernel(const double* __restrict__ C, long long ldc, int m, int n,
                                                     const double* __restrict__ w, unsigned long long* __restrict__ amax) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  const int k0 = blockIdx.y * 128, k1 = min(m, k0 + 128);
  if (i >= n) return;
  double mx = 0.0;
  bool bad = false;  // NaN (negative or NaN weight, NaN entry) or inf: must reach H as it does in the FP64 kernel
#pragma unroll 8
  for (int k = k0; k < k1; ++k) {
    const double v = sqrt(w[k]) * fabs(C[(long long)k * ldc + i]);
    bad |= !(v <= 1.7976931348623157e308);
    mx = fmax(mx, v);
  }
  if (bad)  // non-negative doubles order as integers, the NaN pattern above all of them
    atomicMax(amax + i, 0x7FF8000000000000ull);
  else if (mx > 0.0)
    atomicMax(amax + i, (unsigned long long)__double_as_longlong(mx));
}

constexpr int SL_COLS = 32, SL_K = 128, SL_PITCHW = SL_K / 4 + 1;  // tile rows of 32 words (128 digits) + 1 word of padding
__global__ void __launch_bounds__(256) slice_kernel(const double* __restrict__ C, long long ldc, int m, int n,
                                                    const double* __restrict__ w,
                                                    const unsigned long long* __restrict__ amax, int s,
                                                    int8_t* __restrict__ Q, long long n_pad, long long k_pad,
                                                    double* __restrict__ sigma) {
  __shared__ uint32_t tile[SMAX * SL_COLS * SL_PITCHW];  // [slice][column][4 consecutive k per word]
  const int ci = threadIdx.x & 31, kq0 = threadIdx.x >> 5;
  const int i = blockIdx.x * SL_COLS + ci, k0 = blockIdx.y * SL_K;
  double inv = 0.0;
  if (i < n) {
    const unsigned long long ef = amax[i] >> 52;  // exponent field of amax (its sign bit is 0); 0: column of zeros
    const bool finite = ef < 0x7FDull;            // inf / NaN / about to overflow: the column of H becomes NaN
    if (ef != 0 && finite) inv = __longlong_as_double((long long)(2044ull - ef) << 52);  // 2^-(E+2)
    if (blockIdx.y == 0 && kq0 == 0)
      sigma[i] = !finite ? __longlong_as_double(0x7FF8000000000000ll)
                         : (ef != 0 ? __longlong_as_double((long long)(ef + 2ull) << 52) : 0.0);
  }
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: r + magic has rint(r) in its low mantissa bits
#pragma unroll
  for (int g = 0; g < SL_K / 32; ++g) {     // 4 consecutive k per thread and pass: one packed word per slice
    const int kq = kq0 + 8 * g;
    double r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + 4 * kq + e;
      r[e] = (i < n && k < m) ? sqrt(w[k]) * C[(long long)k * ldc + i] * inv : 0.0;  // |r| <= 1/2
    }
    for (int t = 0; t < s; ++t) {
      uint32_t word = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const double tmp = fma(r[e], 128.0, magic);  // r * 128 is exact: one rounding, to the nearest integer
        r[e] = fma(r[e], 128.0, magic - tmp);        // exact; the remainder stays in [-1/2, 1/2]
        word |= ((uint32_t)__double2loint(tmp) & 0xFFu) << (8 * e);
      }
      tile[(t * SL_COLS + ci) * SL_PITCHW + kq] = word;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = warp; row < s * SL_COLS; row += 8) {  // 128 contiguous bytes of k per (slice, column)
    const int t = row / SL_COLS, c = row % SL_COLS;
    const long long gi = (long long)blockIdx.x * SL_COLS + c;
    if (gi >= n) continue;
    *reinterpret_cast<uint32_t*>(Q + ((long long)t * n_pad + gi) * k_pad + k0 + lane * 4) = tile[row * SL_PITCHW + lane];
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
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, unsigned int* fault) {
  return spin_wait([&] { return mbar_try(bar, parity); }, fault, IPM_FAULT_HESS_I8);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// COLLECT: 0 plain, 1 keep the A operand in the collector buffer for the next MMA, 2 reuse it (and release it)
template <int COLLECT>
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  if constexpr (COLLECT == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
  else if constexpr (COLLECT == 2)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
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
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) { return DESC | (uint64_t)((addr & 0x3FFFFu) >> 4); }

struct Args {
  const int2* tiles;  // (row block of 128, column block of 64), upper triangle
  int ntiles, nkc, s, n;
  const double* sigma;
  double* H;
  long long ldh;
  double beta;
  unsigned int* fault;
  // row-sharded mode (SCATTER): the 128 x 64 partial tile is filed in this rank's own exchange buffer under the rank
  // that owns the 128 x 128 tile it is one half of, and that owner's flag is raised (flags as PeerScatterEpilogue in
  // gemm_tn.cu; the owner then PULLS the tile: remote 16-byte stores from this epilogue cost 60 us per tile)
  double* inbox[8];
  unsigned int* flags[8];
  unsigned int* halves;  // per 128 x 128 tile: half tiles stored so far (the second one raises the owner's flag)
  int me, R, slots;
  unsigned int epoch;
};

// all MMAs of one (chunk, slice t): 4 k-steps x the <= 2 pieces of the N = 64 (S - t) wide operand.  Everything but the
// two operand base addresses is a compile-time constant, so one MMA costs one add per descriptor: the issuing warp runs
// this with uniform control flow and only the instruction itself sits behind elect.sync (an `if (lane == 0)` around
// the loop made every operand a per-thread value: ~200 clocks of uniform-datapath conversions per MMA, which starved
// the tensor pipe -- tools/umma_probe.cu, profiles/umma_probe_r02.jsonl).
template <int S, int T>
__device__ __forceinline__ void issue_slice(uint32_t tmem, uint64_t da0, uint64_t db0, bool first_chunk, bool leader) {
  constexpr int NCOLS = TN * (S - T);
#pragma unroll
  for (int j = 0; j < KC / KSTEP; ++j) {
    const uint64_t da = da0 + (uint64_t)((j * KSTEP) >> 4);
    const uint32_t acc = (T > 0 || j > 0) ? 1u : (first_chunk ? 0u : 1u);
    if constexpr (NCOLS > 256) {
      const uint64_t db = db0 + (uint64_t)((j * KSTEP) >> 4), db2 = db + (uint64_t)((256 * KC) >> 4);
      if (leader) {
        umma_i8<1>(tmem + T * TN, da, db, IDESC | ((uint32_t)(256 >> 3) << 17), acc);
        umma_i8<2>(tmem + T * TN + 256, da, db2, IDESC | ((uint32_t)((NCOLS - 256) >> 3) << 17), acc);
      }
    } else {
      if (leader) umma_i8<0>(tmem + T * TN, da, db0 + (uint64_t)((j * KSTEP) >> 4), IDESC | ((uint32_t)(NCOLS >> 3) << 17), acc);
    }
  }
}

template <int S, bool SCATTER>
__global__ void __launch_bounds__(THREADS, 1)
syrk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smB = base, smA = base + NB * B_STAGE;
  const uint32_t bars = smA + NA * A_SLOT;
  const uint32_t a_full = bars, a_empty = bars + 8 * NA, b_full = bars + 16 * NA, b_empty = b_full + 8 * NB,
                 tfull = b_empty + 8 * NB, tempty = tfull + 8, tptr = tfull + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int s = S;

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
  if (warp == 1) {  // all 512 TMEM columns: the launch bounds and the shared-memory size keep this the only CTA of the SM
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
      long long ia = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int2 tl = a.tiles[tile];
        for (int kc = 0; kc < a.nkc; ++kc) {
          for (int t = 0; t < s; ++t, ++ia) {
            const int sl = (int)(ia % NA);
            if (ia >= NA) mbar_wait(a_empty + 8 * sl, (uint32_t)((ia / NA) - 1) & 1u, a.fault);
            mbar_expect_tx(a_full + 8 * sl, A_SLOT);
            tma_load_3d(smA + sl * A_SLOT, &tmA, kc * KC, tl.x * TM, t, a_full + 8 * sl);
          }
        }
      }
    }
  } else if (warp == 6) {  // B side of a chunk: its own thread, so that a stage is refilled the moment it is released
    if (lane == 0) {
      long long ib = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int2 tl = a.tiles[tile];
        for (int kc = 0; kc < a.nkc; ++kc, ++ib) {
          const int st = (int)(ib % NB);
          if (ib >= NB) mbar_wait(b_empty + 8 * st, (uint32_t)((ib / NB) - 1) & 1u, a.fault);
          mbar_expect_tx(b_full + 8 * st, (uint32_t)s * B_SLICE);
          tma_load_3d(smB + st * B_STAGE, &tmB, kc * KC, tl.y * TN, 0, b_full + 8 * st);
        }
      }
    }
  } else if (warp == 1) {  // MMA issue: the whole warp walks the loop (uniform values), one elected lane issues
    const bool leader = elect_one();
    long long ia = 0, ib = 0;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tcount) {
      if (tcount > 0) {  // the epilogue of the previous tile has drained the accumulators
        mbar_wait(tempty, (uint32_t)(tcount - 1) & 1u, a.fault);
        __syncwarp();
        tc_fence_after();
      }
      for (int kc = 0; kc < a.nkc; ++kc, ++ib) {
        const int st = (int)(ib % NB);
        mbar_wait(b_full + 8 * st, (uint32_t)(ib / NB) & 1u, a.fault);
        const uint64_t db0 = smem_desc(smB + st * B_STAGE);
        auto slice = [&](auto tc) {
          constexpr int T = decltype(tc)::value;
          if constexpr (T < S) {
            const int sl = (int)(ia % NA);
            mbar_wait(a_full + 8 * sl, (uint32_t)(ia / NA) & 1u, a.fault);
            __syncwarp();
            tc_fence_after();
            issue_slice<S, T>(tmem, smem_desc(smA + sl * A_SLOT), db0, kc == 0, leader);
            if (leader) umma_commit(a_empty + 8 * sl);  // frees the A slot once these MMAs have read it
            ++ia;
          }
        };
        slice(std::integral_constant<int, 0>{});
        slice(std::integral_constant<int, 1>{});
        slice(std::integral_constant<int, 2>{});
        slice(std::integral_constant<int, 3>{});
        slice(std::integral_constant<int, 4>{});
        slice(std::integral_constant<int, 5>{});
        slice(std::integral_constant<int, 6>{});
        slice(std::integral_constant<int, 7>{});
        if (leader) umma_commit(b_empty + 8 * st);
      }
      if (leader) umma_commit(tfull);
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tcount) {
      const int2 tl = a.tiles[tile];
      mbar_wait(tfull, (uint32_t)tcount & 1u, a.fault);
      __syncwarp();
      tc_fence_after();
      const long long i = (long long)tl.x * TM + row;
      const double si = i < a.n ? a.sigma[i] * 0x1p-14 : 0.0;
      int scatter_t = 0;
      double* scatter_dst = nullptr;
      if constexpr (SCATTER) {
        const int T = (a.n + 127) / 128, ti = tl.x, tj = tl.y >> 1;
        scatter_t = ti * T - ti * (ti - 1) / 2 + (tj - ti);  // upper_tile_index of gemm_tn.cu
        // the tile stays in THIS rank's buffer, filed under its owner: the owner pulls it (ipm_hess_reduce_bcast_pull_f64)
        scatter_dst = a.inbox[a.me] + ((size_t)(scatter_t % a.R) * a.slots + scatter_t / a.R) * (128 * 128) + 64 * (tl.y & 1);
      }
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
        const long long j0 = (long long)tl.y * TN + c0;
        if (SCATTER || (i < a.n && j0 < a.n)) {
          double out[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            double r = 0.0;
#pragma unroll
            for (int d = SMAX - 1; d >= 0; --d)
              if (d < s) r = fma(r, 0x1p-7, (double)v[d][c]);
            out[c] = r * si;
          }
          if constexpr (SCATTER) {  // sigma is zero beyond n: the padding of the tile is stored as zeros
            double* dst = scatter_dst + row * 128 + c0;
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
              const double2 sj = *reinterpret_cast<const double2*>(a.sigma + j0 + c);
              *reinterpret_cast<double2*>(dst + c) = make_double2(out[c] * sj.x, out[c + 1] * sj.y);
            }
          } else {
            double* dst = a.H + i * a.ldh + j0;
            if (j0 + 8 <= a.n) {
#pragma unroll
              for (int c = 0; c < 8; c += 2) {
                const double2 sj = *reinterpret_cast<const double2*>(a.sigma + j0 + c);
                double2 h = make_double2(out[c] * sj.x, out[c + 1] * sj.y);
                if (a.beta != 0.0) {  // only the upper triangle of H is defined on entry
                  const double2 old = *reinterpret_cast<const double2*>(dst + c);
                  if (j0 + c >= i) h.x = fma(a.beta, old.x, h.x);
                  if (j0 + c + 1 >= i) h.y = fma(a.beta, old.y, h.y);
                }
                *reinterpret_cast<double2*>(dst + c) = h;
              }
            } else {
              for (int c = 0; c < 8; ++c)
                if (j0 + c < a.n) {
                  const double h = out[c] * a.sigma[j0 + c];
                  dst[c] = (a.beta != 0.0 && j0 + c >= i) ? fma(a.beta, dst[c], h) : h;
                }
            }
          }
        }
      }
      if constexpr (SCATTER) {  // publish: all 128 rows stored and visible system-wide, then count the half tile
        __threadfence_system();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (row == 0) {
          if (atomicAdd(a.halves + scatter_t, 1u) & 1u) {
            __threadfence_system();
            unsigned int* f = a.flags[scatter_t % a.R] + (size_t)(scatter_t / a.R) * a.R + a.me;
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(a.epoch) : "memory");
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

int slice_map(CUtensorMap* tm, const int8_t* Q, long long n_pad, long long k_pad, int s, int box_rows, int box_slices) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return IPM_ERR_NO_DEVICE;
  cuuint64_t gdim[3] = {(cuuint64_t)k_pad, (cuuint64_t)n_pad, (cuuint64_t)s};
  cuuint64_t gstride[2] = {(cuuint64_t)k_pad, (cuuint64_t)(n_pad * k_pad)};
  cuuint32_t box[3] = {KC, (cuuint32_t)box_rows, (cuuint32_t)box_slices};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)Q, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? IPM_OK : IPM_ERR_ARG;
}

template <int S, bool SCATTER>
int launch_syrk_s(int grid, cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, const Args& a) {
  static bool attr_done[kMaxDevices] = {};
  IPM_CUDA_CHECK(ensure_dynamic_smem(syrk_kernel<S, SCATTER>, SMEM_BYTES, attr_done));
  syrk_kernel<S, SCATTER><<<grid, THREADS, SMEM_BYTES, st>>>(tmA, tmB, a);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}
template <bool SCATTER>
int launch_syrk(int slices, int grid, cudaStream_t st, const CUtensorMap& tmA, const CUtensorMap& tmB, const Args& a) {
  switch (slices) {
    case 1: return launch_syrk_s<1, SCATTER>(grid, st, tmA, tmB, a);
    case 2: return launch_syrk_s<2, SCATTER>(grid, st, tmA, tmB, a);
    case 3: return launch_syrk_s<3, SCATTER>(grid, st, tmA, tmB, a);
    case 4: return launch_syrk_s<4, SCATTER>(grid, st, tmA, tmB, a);
    case 5: return launch_syrk_s<5, SCATTER>(grid, st, tmA, tmB, a);
    case 6: return launch_syrk_s<6, SCATTER>(grid, st, tmA, tmB, a);
    case 7: return launch_syrk_s<7, SCATTER>(grid, st, tmA, tmB, a);
    default: return launch_syrk_s<8, SCATTER>(grid, st, tmA, tmB, a);
  }
}

long long round_up_ll(long long a, long long b) { return (a + b - 1) / b * b; }

// workspace: [amax n_pad u64][sigma n_pad f64][tiles ntiles int2][half-tile counters][pad to 1024][Q s x n_pad x k_pad int8]
struct Layout {
  long long n_pad, k_pad, off_sigma, off_tiles, off_halves, off_q, bytes;
  int ntiles;
};
Layout layout_for(int m, int n, int s) {
  Layout L;
  L.n_pad = round_up_ll(n, TM);
  L.k_pad = round_up_ll(m, KC);
  const int rb = (int)(L.n_pad / TM);
  L.ntiles = rb * (rb + 1);  // sum over row blocks bi of the 2 rb - 2 bi column blocks of 64 with bj >= 2 bi
  L.off_sigma = L.n_pad * 8;
  L.off_tiles = 2 * L.n_pad * 8;
  L.off_halves = L.off_tiles + (long long)L.ntiles * 8;
  L.off_q = round_up_ll(L.off_halves + (long long)rb * (rb + 1) / 2 * 4, 1024);
  L.bytes = L.off_q + (long long)s * L.n_pad * L.k_pad;
  return L;
}
}  // namespace

// Bytes of the workspace of ipm_hess_i8_f64 for an m x n operand cut into `slices` digits (0 on invalid arguments).
extern "C" long long ipm_hess_i8_ws_bytes(int m, int n, int slices) {
  if (m <= 0 || n <= 0 || slices < 1 || slices > SMAX || m > kMaxRows) return 0;
  return layout_for(m, n, slices).bytes;
}

// One-time set-up of a workspace (256-byte aligned device memory of ipm_hess_i8_ws_bytes bytes): zeroes the slice
// buffer -- its padding rows / columns must stay zero, the slicing kernel never writes them -- and uploads the list of
// upper-triangle output tiles (synchronous copy: call it outside the hot loop).
extern "C" int ipm_hess_i8_prepare(void* ws, int m, int n, int slices, void* stream) {
  if (!ws || ((uintptr_t)ws & 255) || ipm_hess_i8_ws_bytes(m, n, slices) == 0) return IPM_ERR_ARG;
  const Layout L = layout_for(m, n, slices);
  cudaStream_t st = (cudaStream_t)stream;
  IPM_CUDA_CHECK(cudaMemsetAsync(ws, 0, (size_t)L.bytes, st));
  const int rb = (int)(L.n_pad / TM), cb = 2 * rb;
  int* host = new int[2 * (size_t)L.ntiles];
  int cnt = 0;
  for (int bi = 0; bi < rb; ++bi)  // concurrent CTAs share the A rows of one or two row blocks
    for (int bj = 2 * bi; bj < cb; ++bj) {
      host[2 * cnt] = bi;
      host[2 * cnt + 1] = bj;
      ++cnt;
    }
  IPM_CUDA_CHECK(cudaStreamSynchronize(st));
  cudaError_t e = cudaMemcpy((char*)ws + L.off_tiles, host, sizeof(int) * 2 * (size_t)L.ntiles, cudaMemcpyHostToDevice);
  delete[] host;
  IPM_CUDA_CHECK(e);
  return cnt == L.ntiles ? IPM_OK : IPM_ERR_ARG;
}

namespace {
// slicing + SYRK shared by the two entry points; H == nullptr: scatter mode
int run_hess_i8(const double* C, int ldc, int m, int n, const double* w, double beta, double* H, int ldh, int slices,
                void* ws, void* const* peer_inbox, void* const* peer_flags, int me, int R, int slots, unsigned int epoch,
                void* stream) {
  const Layout L = layout_for(m, n, slices);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* amax = (unsigned long long*)ws;
  double* sigma = (double*)((char*)ws + L.off_sigma);
  const int2* tiles = (const int2*)((char*)ws + L.off_tiles);
  int8_t* Q = (int8_t*)((char*)ws + L.off_q);
  CUtensorMap tmA, tmB;
  int rc = slice_map(&tmA, Q, L.n_pad, L.k_pad, slices, TM, 1);
  if (rc) return rc;
  rc = slice_map(&tmB, Q, L.n_pad, L.k_pad, slices, TN, slices);
  if (rc) return rc;
  int dev = 0, sms = 0;
  IPM_CUDA_CHECK(cudaGetDevice(&dev));
  IPM_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

  IPM_CUDA_CHECK(cudaMemsetAsync(amax, 0, (size_t)L.n_pad * 8, st));
  colmax_kernel<<<dim3(ceil_div(n, 128), ceil_div(m, 128)), 128, 0, st>>>(C, ldc, m, n, w, amax);
  IPM_LAUNCH_CHECK();
  slice_kernel<<<dim3((unsigned)(L.n_pad / SL_COLS), (unsigned)(L.k_pad / SL_K)), 256, 0, st>>>(C, ldc, m, n, w, amax, slices,
                                                                                              Q, L.n_pad, L.k_pad, sigma);
  IPM_LAUNCH_CHECK();
  Args a = {};
  a.tiles = tiles;
  a.ntiles = L.ntiles;
  a.nkc = (int)(L.k_pad / KC);
  a.s = slices;
  a.n = n;
  a.sigma = sigma;
  a.H = H;
  a.ldh = ldh;
  a.beta = beta;
  a.fault = ipm_internal_fault_word();
  const int grid = L.ntiles < sms ? L.ntiles : sms;
  if (H) return launch_syrk<false>(slices, grid, st, tmA, tmB, a);
  for (int r = 0; r < R; ++r) {
    a.inbox[r] = (double*)peer_inbox[r];
    a.flags[r] = (unsigned int*)peer_flags[r];
  }
  a.halves = (unsigned int*)((char*)ws + L.off_halves);
  a.me = me, a.R = R, a.slots = slots, a.epoch = epoch;
  return launch_syrk<true>(slices, grid, st, tmA, tmB, a);
}
}  // namespace

// H (upper tiles; n x n, ldh) = beta * H + C' diag(w) C  for C: m x n (ldc), w: m weights >= 0, through `slices` INT8
// digits per entry (8: FP64-level accuracy; fewer digits are faster and less accurate, 7 bits each).  ws: prepared by
// ipm_hess_i8_prepare for the same (m, n, slices).  Elements below the diagonal inside diagonal tiles are written too
// (as by ipm_gemm_tn_f64 with upper = 1).  FunctionManager.py:301-312, 564-576, 801-813.
extern "C" int ipm_hess_i8_f64(const double* C, int ldc, int m, int n, const double* w, double beta, double* H, int ldh,
                               int slices, void* ws, void* stream) {
  if (!C || !w || !H || !ws || ((uintptr_t)ws & 255) || ldc < n || ldh < n || (ldh & 1) || ((uintptr_t)H & 15) ||
      ipm_hess_i8_ws_bytes(m, n, slices) == 0)
    return IPM_ERR_ARG;
  return run_hess_i8(C, ldc, m, n, w, beta, H, ldh, slices, ws, nullptr, nullptr, 0, 1, 0, 0, stream);
}

// Row-sharded variant: the partial Hessian of this rank's m local rows, filed tile by tile under the owners in this
// rank's exchange buffer (peer_inbox[me] + (owner * slots + slot) tiles), the owners' flags raised as by
// ipm_syrk_scatter_f64 (same buffers, slots and epoch; no local addend); to be followed by
// ipm_hess_reduce_bcast_pull_f64.  A 128 x 128 tile is finished as two 128 x 64 halves by two CTAs; the one that stores
// second raises the owner's flag.
extern "C" int ipm_hess_i8_scatter_f64(const double* C, int ldc, int m, int n, const double* w, int slices, void* ws,
                                       void* const* peer_inbox, void* const* peer_flags, int me, int R, int slots,
                                       unsigned int epoch, void* stream) {
  if (!C || !w || !ws || ((uintptr_t)ws & 255) || ldc < n || !peer_inbox || !peer_flags || R < 1 || R > 8 || me < 0 ||
      me >= R || slots < 1 || epoch == 0 || ipm_hess_i8_ws_bytes(m, n, slices) == 0)
    return IPM_ERR_ARG;
  return run_hess_i8(C, ldc, m, n, w, 0.0, nullptr, 0, slices, ws, peer_inbox, peer_flags, me, R, slots, epoch, stream);
}
