// FP64 "TN" contraction core for sm_100a:   acc[m][n] = sum_k  w[k] * A[k][m] * B[k][n]
//
// A is K x M and B is K x N, both row-major (the contracted index k is the ROW index, so a tile
// [16 k-rows] x [128 contiguous columns] is what both operands look like in HBM).  That one flavour
// serves every dense contraction of the interior-point path:
//   * Hessian         H = C^T diag(w) C              (A = B = C, upper tiles only)
//   * Cholesky update A22 -= U12^T U12               (A = B = U12, alpha = -1, beta = 1)
//   * Schur           S = Y^T Y,  TRSM updates, Lasso Q~ (u - alpha)   (A != B)
//
// Blackwell mapping: tcgen05.mma has no FP64 kind, so the math is DMMA (mma.sync.m8n8k4.f64, the only
// FP64 tensor shape sm_100a issues natively -- m16n8k{4,8,16} lower to it); operand tiles are staged by
// TMA (cp.async.bulk.tensor.2d, 128B swizzle) into a 4-stage mbarrier ring by one producer warp and
// consumed by 8 MMA warps (warp tile 64x32, 64 FP64 accumulators per thread).
//
// Shared-memory tile layout (per operand, per stage): 8 column chunks of [16 k-rows][16 doubles = 128 B],
// each written by one TMA box with CU_TENSOR_MAP_SWIZZLE_128B: 16-byte unit c of row r lands at unit
// c ^ (r & 7).  An m8n8k4 fragment needs (k = lane&3, m = lane>>2); mapping the four k of one MMA to rows
// {0,2,4,6} / {1,3,5,7} / {8,..} / {9,..} makes the 16 lanes of each half-warp hit 16 distinct 8-byte bank
// pairs, i.e. conflict-free LDS.64 (the k order inside a tile is arbitrary as long as A and B agree).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ipm {
namespace gemm {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int CHUNK_BYTES = BK * 128;               // one TMA box: 16 rows x 128 B
constexpr int OPERAND_BYTES = (BM / 16) * CHUNK_BYTES;  // 16 KiB
constexpr int STAGE_BYTES = 2 * OPERAND_BYTES;      // A + B
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 2 * STAGES * 8;

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Decode the linear CTA index into a tile pair.  upper: enumerate tiles (ti <= tj) row by row.
__device__ __forceinline__ void decode_tile(int lin, int tiles_m, int tiles_n, bool upper, int& ti, int& tj) {
  if (!upper) {
    ti = lin / tiles_n;
    tj = lin - ti * tiles_n;
    return;
  }
  // row ti starts at offset ti*T - ti*(ti-1)/2 ; invert with a float guess + fix-up
  const int T = tiles_n;
  float tf = (2.0f * T + 1.0f - sqrtf((2.0f * T + 1.0f) * (2.0f * T + 1.0f) - 8.0f * (float)lin)) * 0.5f;
  int t = (int)tf;
  if (t < 0) t = 0;
  if (t > T - 1) t = T - 1;
  while (t > 0 && (long long)t * T - (long long)t * (t - 1) / 2 > lin) --t;
  while ((long long)(t + 1) * T - (long long)(t + 1) * t / 2 <= lin) ++t;
  ti = t;
  tj = t + (lin - (t * T - t * (t - 1) / 2));
}

// Epilogue concept:  void tile(const double (&acc)[8][4][2], int m_base, int n_base, int g8, int l4) const
//   called once per consumer warp with its 64x32 accumulator tile; the functor does its own bounds checks.
template <bool HAS_W, class Epilogue>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
               const double* __restrict__ w, int upper, Epilogue epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
  const uint32_t tiles0 = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  {
    // CTAs beyond the tile count belong to the epilogue functor (e.g. the Lasso tail rows that would otherwise
    // cost a whole extra tile row); functors without such work never get launched with extra CTAs.
    const int ntiles = upper ? tiles_n * (tiles_n + 1) / 2 : tiles_m * tiles_n;
    if ((int)blockIdx.x >= ntiles) {
      epi.extra((int)blockIdx.x - ntiles);
      return;
    }
  }
  int ti, tj;
  decode_tile(blockIdx.x, tiles_m, tiles_n, upper != 0, ti, tj);
  const int m0 = ti * BM, n0 = tj * BN;
  const int ktiles = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == CONSUMER_WARPS) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      for (int kt = 0; kt < ktiles; ++kt) {
        const int s = kt % STAGES;
        if (kt >= STAGES) mbar_wait(empty0 + 8 * s, ((kt / STAGES) - 1) & 1);
        const uint32_t full = full0 + 8 * s;
        mbar_expect_tx(full, STAGE_BYTES);
        const uint32_t dstA = tiles0 + s * STAGE_BYTES, dstB = dstA + OPERAND_BYTES;
#pragma unroll
        for (int c = 0; c < BM / 16; ++c) tma_load_2d(dstA + c * CHUNK_BYTES, &tmA, m0 + c * 16, kt * BK, full);
#pragma unroll
        for (int c = 0; c < BN / 16; ++c) tma_load_2d(dstB + c * CHUNK_BYTES, &tmB, n0 + c * 16, kt * BK, full);
      }
    }
    return;
  }

  // ---------------- DMMA consumers: warp grid 2 (m) x 4 (n), warp tile 64 x 32 ----------------
  const int wm = warp >> 2, wn = warp & 3;
  const int l4 = lane & 3, g8 = lane >> 2;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // byte offset of this lane's element inside a [16 x 128B] chunk, for even / odd column-block index and for
  // the two row parities (jb = j & 1):   unit16 = ((blk & 1) * 4 + (g8 >> 1)) ^ (2 * l4 + jb)
  uint32_t off[2][2];
#pragma unroll
  for (int e = 0; e < 2; ++e)
#pragma unroll
    for (int jb = 0; jb < 2; ++jb)
      off[e][jb] = (uint32_t)(((((e * 4) + (g8 >> 1)) ^ (2 * l4 + jb)) << 4) + (g8 & 1) * 8);

  const uint32_t a_warp = (uint32_t)(wm * 4) * CHUNK_BYTES;                   // 64 cols = 4 chunks
  const uint32_t b_warp = OPERAND_BYTES + (uint32_t)(wn * 2) * CHUNK_BYTES;   // 32 cols = 2 chunks

  for (int kt = 0; kt < ktiles; ++kt) {
    const int s = kt % STAGES;
    double wk[4];
    if (HAS_W) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = kt * BK + (j >> 1) * 8 + 2 * l4 + (j & 1);
        wk[j] = kk < K ? __ldg(w + kk) : 0.0;
      }
    }
    mbar_wait(full0 + 8 * s, (kt / STAGES) & 1);
    const uint32_t st = tiles0 + s * STAGE_BYTES;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t rowoff = (uint32_t)((j >> 1) * 8 + 2 * l4 + (j & 1)) * 128u;
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t addr = st + a_warp + (uint32_t)(i >> 1) * CHUNK_BYTES + rowoff + off[i & 1][j & 1];
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[i]) : "r"(addr));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t addr = st + b_warp + (uint32_t)(i >> 1) * CHUNK_BYTES + rowoff + off[i & 1][j & 1];
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b[i]) : "r"(addr));
        if (HAS_W) b[i] *= wk[j];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) dmma884(acc[i][jn][0], acc[i][jn][1], a[i], b[jn]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * s);
  }

  // ---------------- epilogue ----------------
  // The functor sees the warp's whole 64x32 accumulator tile so it can batch its global loads:
  //   acc[i][jn][e] -> row = m_base + i*8 + g8,  col = n_base + jn*8 + 2*l4 + e
  epi.tile(acc, m0 + wm * 64, n0 + wn * 32, g8, l4);
}

}  // namespace gemm
}  // namespace ipm
