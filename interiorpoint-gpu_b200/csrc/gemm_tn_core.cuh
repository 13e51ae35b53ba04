// FP64 "TN" contraction core for sm_100a:   acc[m][n] = sum_k  w[k] * A[k][m] * B[k][n]
//
// A is K x M and B is K x N, both row-major (the contracted index k is the ROW index, so a tile
// [32 k-rows] x [128 contiguous columns] is what both operands look like in HBM).  That one flavour
// serves every dense contraction of the interior-point path:
//   * Hessian         H = C^T diag(w) C              (A = B = C, upper tiles only)
//   * Cholesky update A22 -= U12^T U12               (A = B = U12, alpha = -1, beta = 1)
//   * Schur           S = Y^T Y,  TRSM updates, Lasso Q~ (u - alpha)   (A != B)
//
// Blackwell mapping: tcgen05.mma has no FP64 kind, so the math is DMMA (mma.sync.m8n8k4.f64, the only
// FP64 tensor shape sm_100a issues natively -- m16n8k{4,8,16} lower to it); operand tiles are staged by
// TMA (cp.async.bulk.tensor.2d, 128B swizzle) into a 3-stage mbarrier ring of 32-row k-tiles and consumed by 8 MMA warps (warp
// tile 64x32, 64 FP64 accumulators per thread).  There is no dedicated producer warp: a ninth warp would put
// three warps on one SM sub-partition and cap every thread at 168 registers (16K registers per sub-partition),
// which spills the double-buffered fragments.  Instead every warp keeps the ring PREFETCH k-tiles ahead of its
// own consumption: lane 0 of warp c issues the two TMA boxes of column chunk c (one of A, one of B) of each
// k-tile, so the producer work is symmetric (a single producing warp becomes the pace-setter of the CTA: the
// other seven run into the prefetch horizon and idle their DMMA pipes).
//
// Shared-memory tile layout (per operand, per stage): 8 column chunks of [32 k-rows][16 doubles = 128 B],
// each written by one TMA box with CU_TENSOR_MAP_SWIZZLE_128B: 16-byte unit c of row r lands at unit
// c ^ (r & 7).  An m8n8k4 fragment needs (k = lane&3, m = lane>>2); mapping the four k of one MMA to rows
// {0,2,4,6} / {1,3,5,7} / {8,..} / {9,..} / ... makes the 16 lanes of each half-warp hit 16 distinct 8-byte bank
// pairs, i.e. conflict-free LDS.64 (the k order inside a tile is arbitrary as long as A and B agree).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace ipm {
namespace gemm {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int KGROUPS = BK / 4;        // m8n8k4 k-groups per k-tile
constexpr int PREFETCH = STAGES - 1;   // k-tiles issued ahead of a warp's own consumption
// k-group at which a warp tops the ring up (refills the slot everybody released during the previous k-tile).  The
// later, the less the empty-wait spins: measured on the persistent Hessian kernel 2 / 4 / 6 / 7 -> 33.0 / 33.0 / 32.7 /
// 32.4 ms (issuing at the start of the NEXT k-tile with one k-tile of prefetch: 32.5).  The one-CTA-per-tile kernel
// has short k loops (Lasso: 16 k-tiles) and prefers the data earlier: 4 -> 95.3 us, 7 -> 96.5 us per ADMM iteration.
constexpr int ISSUE_AT_PERSISTENT = 7, ISSUE_AT_ONE_TILE = 4;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = CONSUMER_WARPS * 32;
constexpr int CHUNK_BYTES = BK * 128;               // one TMA box: BK rows x 128 B
constexpr int OPERAND_BYTES = (BM / 16) * CHUNK_BYTES;  // 32 KiB
constexpr int STAGE_BYTES = 2 * OPERAND_BYTES;      // A + B
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 2 * STAGES * 8;

// CTA tile shape: 8 warps as a WM x (8 / WM) grid, each warp MI x NI blocks of 8 x 8 (MI * NI * 2 accumulators per
// thread).  The pipeline pieces below are written against the shape so that other warp grids can be tried; the two
// that were (128 x 112 as 4 x 2 warps for a one-wave Lasso grid, and 16 warps of 32 x 32) measured slower and are not
// kept (negative results).
struct Shape128x128 {
  static constexpr int WM = 2, MI = 8, NI = 4;
};
template <class S>
struct ShapeTraits {
  static constexpr int WN = CONSUMER_WARPS / S::WM;
  static constexpr int TILE_M = S::WM * S::MI * 8, TILE_N = WN * S::NI * 8;
  static constexpr int A_CHUNKS = TILE_M / 16, B_CHUNKS = TILE_N / 16;
  static_assert(TILE_M == BM && TILE_N <= BN && TILE_N % 16 == 0, "unsupported CTA tile");
};

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

// L2-friendly enumeration for the persistent kernel: tile rows are grouped into bands of BAND rows and a band is
// walked column by column, so the ~148 tiles in flight at any time form a ~12 x 12 block that shares 12 A panels and
// ~13 B panels (instead of 1 + 148 with the row-major order): every k-slice of a panel is fetched from HBM once and
// hit in L2 by the other CTAs of the block.  upper: only tiles with tj >= ti.
constexpr int BAND = 12;
__device__ __forceinline__ void decode_tile_banded(int lin, int tiles_m, int tiles_n, bool upper, int& ti, int& tj) {
  int b0 = 0, h = 0;
  while (true) {
    h = min(BAND, tiles_m - b0);
    const int cnt = upper ? h * (tiles_n - b0) - h * (h - 1) / 2 : h * tiles_n;
    if (lin < cnt || b0 + h >= tiles_m) break;
    lin -= cnt;
    b0 += h;
  }
  if (!upper) {
    tj = lin / h;
    ti = b0 + lin - tj * h;
    return;
  }
  const int tri = h * (h + 1) / 2;  // the band's first h columns hold 1, 2, ..., h tiles
  if (lin < tri) {
    int c = 0;
    while ((c + 1) * (c + 2) / 2 <= lin) ++c;
    ti = b0 + (lin - c * (c + 1) / 2);
    tj = b0 + c;
  } else {
    lin -= tri;
    const int c = lin / h;
    tj = b0 + h + c;
    ti = b0 + lin - c * h;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Pipeline pieces.  `it` counts the k-tiles that went through the CTA's ring (stage = it % STAGES, phase parity
// = (it / STAGES) & 1); the producer cursor and the consumers advance it identically, so the ring keeps
// streaming across tile boundaries in the persistent kernel.
// ---------------------------------------------------------------------------------------------------------
struct Ring {
  uint32_t full0, empty0, tiles0;
};

__device__ __forceinline__ Ring setup_ring(uint8_t* smem_raw, int arrivals = CONSUMER_WARPS) {
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  Ring ring{smem_u32(bars), smem_u32(bars + STAGES), smem_u32(smem)};
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(ring.full0 + 8 * s, arrivals);
      mbar_init(ring.empty0 + 8 * s, arrivals);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  return ring;
}

// Work of one CTA = a short list of segments (output tile, k-tile range).  Schedule concept:
//   int count() const;   void segment(int sg, int& m0, int& n0, int& k0, int& k1) const;   (k1 > k0)
//
// Producer cursor (every warp keeps its own copy; lane 0 issues): walks the segment list one k-tile per call.
template <class Schedule, int B_CHUNKS = BN / 16>
struct Producer {
  const Schedule& sch;
  const CUtensorMap *tmA, *tmB;
  Ring ring;
  int sg, kt, k1, m0, n0;
  uint32_t it;

  __device__ __forceinline__ Producer(const Schedule& s, const CUtensorMap* a, const CUtensorMap* b, const Ring& r)
      : sch(s), tmA(a), tmB(b), ring(r), sg(-1), kt(0), k1(0), m0(0), n0(0), it(0) {
    next_segment();
  }
  __device__ __forceinline__ void next_segment() {
    ++sg;
    if (sg < sch.count()) sch.segment(sg, m0, n0, kt, k1);
  }
  // Every warp calls this once per consumed k-tile: lane 0 of warp `wp` waits until the ring slot is free, posts
  // its share of the transaction bytes and issues the two TMA boxes (column chunk `wp` of A and of B).
  __device__ __forceinline__ void issue(int wp, int lane) {
    if (sg >= sch.count()) return;
    if (lane == 0) {
      const uint32_t s = it % STAGES;
      if (it >= STAGES) mbar_wait(ring.empty0 + 8 * s, ((it / STAGES) - 1) & 1);
      const uint32_t full = ring.full0 + 8 * s;
      mbar_expect_tx(full, (wp < B_CHUNKS ? 2 : 1) * CHUNK_BYTES);
      const uint32_t dstA = ring.tiles0 + s * STAGE_BYTES + wp * CHUNK_BYTES, dstB = dstA + OPERAND_BYTES;
      tma_load_2d(dstA, tmA, m0 + wp * 16, kt * BK, full);
      if (wp < B_CHUNKS) tma_load_2d(dstB, tmB, n0 + wp * 16, kt * BK, full);
    }
    __syncwarp();
    ++it;
    if (++kt >= k1) next_segment();
  }
};

// Per-lane constants of a consumer warp.  Column block b (8 columns) of an operand lives in chunk b >> 1 at
// 16-byte unit ((b & 1) * 4 + (g8 >> 1)) ^ (2 * l4 + jb) of row (.., jb = row parity); the warp's first block may be
// odd for shapes with odd NI, so its parity is folded into the two offset tables.
struct LaneMap {
  uint32_t a_off[2][2], b_off[2][2];  // [block index parity relative to the warp's first block][row parity]
  int a_blk0, b_blk0;
  int l4, g8, wm, wn;
};

template <class S>
__device__ __forceinline__ LaneMap make_lane_map(int warp, int lane) {
  LaneMap lm;
  lm.l4 = lane & 3;
  lm.g8 = lane >> 2;
  constexpr int WN = ShapeTraits<S>::WN;
  lm.wm = warp / WN, lm.wn = warp % WN;
  lm.a_blk0 = lm.wm * S::MI, lm.b_blk0 = lm.wn * S::NI;
#pragma unroll
  for (int e = 0; e < 2; ++e)
#pragma unroll
    for (int jb = 0; jb < 2; ++jb) {
      const int pa = (lm.a_blk0 + e) & 1, pb = (lm.b_blk0 + e) & 1;
      lm.a_off[e][jb] = (uint32_t)((((pa * 4) + (lm.g8 >> 1)) ^ (2 * lm.l4 + jb)) << 4) + (lm.g8 & 1) * 8;
      lm.b_off[e][jb] = (uint32_t)((((pb * 4) + (lm.g8 >> 1)) ^ (2 * lm.l4 + jb)) << 4) + (lm.g8 & 1) * 8;
    }
  return lm;
}

// Fragments of k-group j (4 k-rows) of the stage at shared address st.
template <class S>
__device__ __forceinline__ void load_frags(double (&a)[S::MI], double (&b)[S::NI], uint32_t st, const LaneMap& lm,
                                           int j) {
  const uint32_t rowoff = (uint32_t)((j >> 1) * 8 + 2 * lm.l4 + (j & 1)) * 128u;
#pragma unroll
  for (int i = 0; i < S::MI; ++i) {
    const uint32_t addr = st + (uint32_t)((lm.a_blk0 + i) >> 1) * CHUNK_BYTES + rowoff + lm.a_off[i & 1][j & 1];
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[i]) : "r"(addr));
  }
#pragma unroll
  for (int i = 0; i < S::NI; ++i) {
    const uint32_t addr =
        st + OPERAND_BYTES + (uint32_t)((lm.b_blk0 + i) >> 1) * CHUNK_BYTES + rowoff + lm.b_off[i & 1][j & 1];
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b[i]) : "r"(addr));
  }
}

// Row weights of the four k-groups 4h .. 4h+3 (h = half-tile index, counted over the whole contraction: 16 rows).
__device__ __forceinline__ void load_weights(double (&wk)[4], const double* __restrict__ w, int half, int K, int l4) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int kk = half * 16 + (j >> 1) * 8 + 2 * l4 + (j & 1);
    wk[j] = kk < K ? __ldg(w + kk) : 0.0;
  }
}

// acc += sum over k-tiles [kt_begin, kt_end).  Software pipelined: the fragments of k-group j+1 (and the row
// weights of the next k-tile) are in flight while the 32 DMMAs of group j issue; every warp tops the TMA ring up by
// its own two boxes per consumed k-tile (prod.issue at k-group ISSUE_AT).
template <bool HAS_W, class S, int ISSUE_AT, class Prod>
__device__ __forceinline__ void consume_ktiles(double (&acc)[S::MI][S::NI][2], const Ring& ring, const LaneMap& lm,
                                               const double* __restrict__ w, int K, int kt_begin, int kt_end,
                                               uint32_t& it, int warp, int lane, Prod& prod) {
  static_assert(KGROUPS == 8, "the weight double-buffering below assumes two 16-row halves per k-tile");
  double wk[4], wn[4];
  if (HAS_W) load_weights(wk, w, 2 * kt_begin, K, lm.l4);
  uint32_t s = it % STAGES;
  mbar_wait(ring.full0 + 8 * s, (it / STAGES) & 1);
  uint32_t st = ring.tiles0 + s * STAGE_BYTES;
  double a[2][S::MI], b[2][S::NI];
  load_frags<S>(a[0], b[0], st, lm, 0);
  for (int kt = kt_begin; kt < kt_end; ++kt) {
    const bool has_next = kt + 1 < kt_end;
    uint32_t s_next = s, st_next = st;
#pragma unroll
    for (int j = 0; j < KGROUPS; ++j) {
      const int cur = j & 1, nxt = cur ^ 1;
      if (HAS_W && j == 0) load_weights(wn, w, 2 * kt + 1, K, lm.l4);                  // second half of this tile
      if (HAS_W && j == 4 && has_next) load_weights(wn, w, 2 * kt + 2, K, lm.l4);      // first half of the next
      if (j == ISSUE_AT) prod.issue(warp, lane);
      if (j < KGROUPS - 1) {
        load_frags<S>(a[nxt], b[nxt], st, lm, j + 1);
      } else if (has_next) {
        s_next = (it + 1) % STAGES;
        mbar_wait(ring.full0 + 8 * s_next, ((it + 1) / STAGES) & 1);
        st_next = ring.tiles0 + s_next * STAGE_BYTES;
        load_frags<S>(a[nxt], b[nxt], st_next, lm, 0);
      }
      if (HAS_W) {
#pragma unroll
        for (int i = 0; i < S::NI; ++i) b[cur][i] *= wk[j & 3];
      }
#pragma unroll
      for (int i = 0; i < S::MI; ++i)
#pragma unroll
        for (int jn = 0; jn < S::NI; ++jn) dmma884(acc[i][jn][0], acc[i][jn][1], a[cur][i], b[cur][jn]);
      if (HAS_W && (j & 3) == 3) {
#pragma unroll
        for (int q = 0; q < 4; ++q) wk[q] = wn[q];
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(ring.empty0 + 8 * s);
    ++it;
    s = s_next;
    st = st_next;
  }
}

template <int MI, int NI>
__device__ __forceinline__ void zero_acc(double (&acc)[MI][NI][2]) {
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// Epilogue concept:  template <int MI, int NI> void tile([const] double (&acc)[MI][NI][2], int m_base, int n_base,
//                                                        int g8, int l4) const
//   called once per consumer warp with its (8 MI) x (8 NI) accumulator tile; the functor does its own bounds checks.
//   acc[i][jn][e] -> row = m_base + i*8 + g8,  col = n_base + jn*8 + 2*l4 + e
//                    void prefetch(int m0, int n0, int tile_m, int tile_n) const
//   called by every thread when the CTA starts (one-CTA-per-tile kernel only): L2 prefetch of epilogue operands.
//                    void after_tile(int ti, int tiles_m, int n0, int tile_n) const
//   called by every thread of the CTA after its tile (one-CTA-per-tile kernel only): per-tile extra work such as
//   the Lasso tail rows.
//
// ---------------------------------------------------------------------------------------------------------
// One CTA per output tile (grid = #tiles [+ extra CTAs owned by the epilogue functor]).  Used for short-K
// contractions (Cholesky / TRSM updates, Schur, Lasso) where CTAs must retire quickly so that concurrent streams
// (the Cholesky look-ahead chain) get SMs.
// ---------------------------------------------------------------------------------------------------------
struct OneTile {
  int m0, n0, ktiles;
  __device__ __forceinline__ int count() const { return ktiles > 0 ? 1 : 0; }
  __device__ __forceinline__ void segment(int, int& m, int& n, int& k0, int& k1) const {
    m = m0, n = n0, k0 = 0, k1 = ktiles;
  }
};

template <bool HAS_W, class Epilogue, class S = Shape128x128>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
               const double* __restrict__ w, int upper, Epilogue epi) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int TN = ShapeTraits<S>::TILE_N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + TN - 1) / TN;
  {
    // CTAs beyond the tile count belong to the epilogue functor (e.g. the Lasso tail rows that would otherwise
    // cost a whole extra tile row); functors without such work never get launched with extra CTAs.
    const int ntiles = upper ? tiles_n * (tiles_n + 1) / 2 : tiles_m * tiles_n;
    if ((int)blockIdx.x >= ntiles) {
      pdl_wait();
      epi.extra((int)blockIdx.x - ntiles);
      return;
    }
  }
  const Ring ring = setup_ring(smem_raw);
  int ti, tj;
  decode_tile(blockIdx.x, tiles_m, tiles_n, upper != 0, ti, tj);
  pdl_wait();  // everything above (barrier init, tile decode) overlaps the previous kernel's drain
  const OneTile sch{ti * BM, tj * TN, (K + BK - 1) / BK};
  Producer<OneTile, ShapeTraits<S>::B_CHUNKS> prod(sch, &tmA, &tmB, ring);
  for (int p = 0; p < PREFETCH; ++p) prod.issue(warp, lane);
  epi.prefetch(sch.m0, sch.n0, BM, TN);
  const LaneMap lm = make_lane_map<S>(warp, lane);
  double acc[S::MI][S::NI][2];
  zero_acc(acc);
  uint32_t it = 0;
  if (sch.ktiles > 0)
    consume_ktiles<HAS_W, S, ISSUE_AT_ONE_TILE>(acc, ring, lm, w, K, 0, sch.ktiles, it, warp, lane, prod);
  epi.tile(acc, sch.m0 + lm.wm * (S::MI * 8), sch.n0 + lm.wn * (S::NI * 8), lm.g8, lm.l4);
  epi.after_tile(ti, tiles_m, sch.n0, TN);
}

// ---------------------------------------------------------------------------------------------------------
// Persistent variant for long-K contractions (the Hessian C' diag(w) C): grid = G <= #SMs CTAs, every CTA
// resident.  The first (ntiles / G) * G tiles are processed data-parallel (tile r*G + c); the remaining
// rem < G tiles would cost a whole extra wave (2080 tiles on 148 SMs = 14.05 waves), so their rem * ktiles
// k-tile units are split evenly over the first `ctas` CTAs instead ("stream-K" remainder).  A CTA whose slice
// does not start at a tile's first k-tile writes its 128x128 partial to its workspace slot and raises its flag;
// the CTA that owns the tile's first k-tile adds the partials in ascending CTA order (deterministic) and runs
// the epilogue.  The TMA ring streams across tile boundaries (no pipeline drain between tiles).
// ---------------------------------------------------------------------------------------------------------
struct StreamK {
  double* partials;       // gridDim.x slots of BM*BN doubles
  unsigned int* flags;    // gridDim.x flags; a slot is valid when its flag equals `epoch`
  unsigned int epoch;
  int ctas;               // CTAs 0 .. ctas-1 share the remainder tiles (1 <= ctas <= min(gridDim.x, rem * ktiles))
  int gemm_ctas;          // 0: every CTA of the grid works on tiles; else CTAs >= gemm_ctas belong to epi.extra()
  unsigned int* fault;    // watchdog word of the owner's wait (common.cuh: spin_wait); may be null
};

struct PersistentSchedule {
  int tiles_m, tiles_n, upper, ktiles, G, c, waves, nseg;
  int seg_tile[2], seg_k0[2], seg_k1[2];
  __device__ __forceinline__ int count() const { return nseg + waves; }
  __device__ __forceinline__ int tile_of(int sg) const { return sg < nseg ? seg_tile[sg] : (sg - nseg) * G + c; }
  __device__ __forceinline__ void segment(int sg, int& m, int& n, int& k0, int& k1) const {
    int ti, tj;
    decode_tile_banded(tile_of(sg), tiles_m, tiles_n, upper != 0, ti, tj);
    m = ti * BM, n = tj * BN;
    k0 = sg < nseg ? seg_k0[sg] : 0;
    k1 = sg < nseg ? seg_k1[sg] : ktiles;
  }
};

__device__ __forceinline__ void cta_barrier_1() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <bool HAS_W, class Epilogue>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tn_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                          int K, const double* __restrict__ w, int upper, Epilogue epi, StreamK sk) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  PersistentSchedule sch;
  sch.tiles_m = (M + BM - 1) / BM, sch.tiles_n = (N + BN - 1) / BN, sch.upper = upper;
  const int ntiles = upper ? sch.tiles_n * (sch.tiles_n + 1) / 2 : sch.tiles_m * sch.tiles_n;
  const int ktiles = (K + BK - 1) / BK;
  const int G = sk.gemm_ctas ? sk.gemm_ctas : gridDim.x, c = blockIdx.x;
  if (c >= G) {  // CTAs owned by the epilogue functor (Lasso tail rows), co-resident with the tile CTAs
    pdl_wait();
    epi.extra(c - G);
    return;
  }
  sch.ktiles = ktiles, sch.G = G, sch.c = c, sch.waves = ntiles / G, sch.nseg = 0;
  const int dp_tiles = sch.waves * G, rem = ntiles - dp_tiles;
  const int P = sk.ctas;
  const long long U = (long long)rem * ktiles;
  const long long u0 = c < P ? U * c / P : 0, u1 = c < P ? U * (c + 1) / P : 0;
  sch.seg_tile[0] = sch.seg_tile[1] = sch.seg_k0[0] = sch.seg_k0[1] = sch.seg_k1[0] = sch.seg_k1[1] = 0;
  if (u1 > u0) {
    // the CTA's (at most two) stream-K segments
    const int ta = (int)(u0 / ktiles), ka = (int)(u0 - (long long)ta * ktiles);
    const int len = (int)(u1 - u0);
    const int first = min(len, ktiles - ka);
    sch.seg_tile[0] = dp_tiles + ta, sch.seg_k0[0] = ka, sch.seg_k1[0] = ka + first, sch.nseg = 1;
    if (len > first) sch.seg_tile[1] = dp_tiles + ta + 1, sch.seg_k0[1] = 0, sch.seg_k1[1] = len - first, sch.nseg = 2;
  }
  const Ring ring = setup_ring(smem_raw);
  pdl_wait();
  Producer<PersistentSchedule> prod(sch, &tmA, &tmB, ring);
  for (int p = 0; p < PREFETCH; ++p) prod.issue(warp, lane);
  using S = Shape128x128;
  const LaneMap lm = make_lane_map<S>(warp, lane);
  uint32_t it = 0;
  double acc[S::MI][S::NI][2];
  for (int sg = 0; sg < sch.count(); ++sg) {
    int m0, n0, k0, k1;
    sch.segment(sg, m0, n0, k0, k1);
    const bool is_sk = sg < sch.nseg;
    zero_acc(acc);
    consume_ktiles<HAS_W, S, ISSUE_AT_PERSISTENT>(acc, ring, lm, w, K, k0, k1, it, warp, lane, prod);
    if (is_sk && k0 != 0) {
      // partial of a tile owned by a lower CTA
      double* slot = sk.partials + (size_t)c * (BM * BN) + tid;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __stcg(slot + ((i * 4 + j) * 2 + 0) * THREADS, acc[i][j][0]);
          __stcg(slot + ((i * 4 + j) * 2 + 1) * THREADS, acc[i][j][1]);
        }
      __threadfence();
      cta_barrier_1();
      if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(sk.flags + c), "r"(sk.epoch) : "memory");
      continue;
    }
    if (is_sk && k1 < ktiles) {
      // owner of the tile's first k-tiles: add the partials of the CTAs that cover the rest of the tile
      const long long tile_end = (long long)(sch.seg_tile[sg] - dp_tiles + 1) * ktiles;
      for (int c2 = c + 1; c2 < P && U * c2 / P < tile_end; ++c2) {
        if (tid == 0) {
          const unsigned int* flag = sk.flags + c2;
          spin_wait(
              [&] {
                unsigned v;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                return v == sk.epoch;
              },
              sk.fault, IPM_FAULT_STREAMK);
        }
        cta_barrier_1();
        const double* slot = sk.partials + (size_t)c2 * (BM * BN) + tid;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[i][j][0] += __ldcg(slot + ((i * 4 + j) * 2 + 0) * THREADS);
            acc[i][j][1] += __ldcg(slot + ((i * 4 + j) * 2 + 1) * THREADS);
          }
      }
    }
    epi.tile(acc, m0 + lm.wm * 64, n0 + lm.wn * 32, lm.g8, lm.l4);
  }
}

}  // namespace gemm
}  // namespace ipm
