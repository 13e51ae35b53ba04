// Batched ADMM Lasso, MANY iterations per launch (LassoSolver.py:240-337: the whole `for it in range(max_iters)` loop body,
// `check_stop` iterations at a time, including the batch-coupled stop test :273-298).
//
//   x      = bA + Q~ z                      z = u - alpha kept as its own array (the MMA operand)
//   alpha+ = prox(x + u, eta_c)             soft threshold per column c, bias row exempt (:517-543)
//   u+     = u + x - alpha+ ,  z+ = u+ - alpha+
//
// The K problems (columns) are independent, so nothing in an iteration needs a grid-wide barrier: the output tile
// (row tile r, column group g) of iteration i+1 only needs z+ of column group g from iteration i.  The kernel is a
// persistent grid (every CTA resident, 2 per SM) that pulls work units
//       unit = (iteration i, column group g, row tile r)          ordered i-major, then g, then r
// from one atomic counter.  done[g] counts the finished units of column group g over all iterations (release add by the
// finishing CTA); unit (i, g, r) waits until done[g] >= i * units_per_group (acquire) and then runs
//       TMA ring (Q~ k-tile, z k-tile) -> FP64 DMMA -> rank-1 terms of the contraction rows beyond the last full k-tile
//       -> prox / dual update epilogue on the accumulators -> z+ into the OTHER z buffer.
// Every dependency of a unit has a smaller unit number and has therefore been claimed by a resident CTA that waits on
// nothing larger: no deadlock.  Consequences on B200 (n = 513, K = 4096):
//   * no kernel boundary and no launch gap between iterations, no global barrier: a CTA that finishes early moves on
//     to the next iteration of a column group that is complete;
//   * units flow continuously over all 148 SMs instead of 128 tiles per launch on 148 SMs (0.86 fill);
//   * (measured and dropped: a 128 x 64 tile with two CTAs per SM so that one CTA's L2-bound epilogue overlaps the other's
//     DMMAs -- only two 48 KB ring stages fit twice, the TMA latency was exposed at every k-tile and the kernel ran at
//     107 us per iteration against 91 us for one launch per iteration; profiles/lasso_multi_r02a.jsonl);
//   * alpha is neither read nor written except in the last two iterations of a launch (the stop test needs alpha and
//     alpha+; in between alpha = u - z is implicit): 4 instead of 6 state arrays per iteration through L2.
// Small batches (the per-GPU shards of a 2/4/8-way split, K <= 2048) use a 64 x 32 tile so that there are still >= 128
// units per iteration; with fewer units than SMs the grid is one CTA per SM, so two units never share an SM while
// others idle.
//
// Stop test on the device: the last iteration of a launch leaves four partial sums of squares per warp in a slot that
// depends only on (g, r, warp); lasso_batch_end_kernel adds them in slot order (deterministic whatever CTA ran the
// unit), evaluates  r < tol_primal && d < tol_dual  (LassoSolver.py:284-298) and raises state[0].  A launch that finds
// state[0] != 0 returns without touching anything, so the host can enqueue launches ahead and read state later.
#include "common.cuh"
#include "gemm_tn_core.cuh"
#include "tensormap.cuh"

namespace ipm {
namespace lasso {
using namespace gemm;  // mbarrier / TMA / DMMA primitives, BK, CHUNK_BYTES

constexpr int LM_THREADS = 256;
constexpr int LM_TAIL_MAX = 8;  // rows beyond the last full row tile handled as dot products (n = 513: the one extra row)

// MINB = CTAs per SM the kernel is compiled for (register budget) and launched with at most.
struct ShapeBig {  // 128 x 128 tile, warps 2 x 4, warp tile 64 x 32 (the shape of gemm_tn_core.cuh), 3 x 64 KB ring
  static constexpr int WM = 2, MI = 8, NI = 4, NSTAGES = 3, MINB = 1;
};
struct ShapeSmall {  // 64 x 32 tile, warps 4 x 2, warp tile 16 x 16, 4 x 24 KB ring
  static constexpr int WM = 4, MI = 2, NI = 2, NSTAGES = 4, MINB = 2;
};
template <class S>
struct Geo {
  static constexpr int WN = 8 / S::WM;
  static constexpr int TM = S::WM * S::MI * 8, TN = WN * S::NI * 8;
  static constexpr int A_CHUNKS = TM / 16, B_CHUNKS = TN / 16;
  static constexpr int A_BYTES = A_CHUNKS * CHUNK_BYTES, B_BYTES = B_CHUNKS * CHUNK_BYTES;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int ISSUERS = A_CHUNKS > B_CHUNKS ? A_CHUNKS : B_CHUNKS;  // warps whose lane 0 issues TMA boxes
  static constexpr int SMEM = 1024 + S::NSTAGES * STAGE + 2 * S::NSTAGES * 8;
  static constexpr int KPARTS = LM_THREADS / TN;  // k-ranges of the tail-row dot products
  static_assert(A_CHUNKS <= 8 && B_CHUNKS <= 8 && STAGE % 1024 == 0, "unsupported tile");
};

// state words at the head of the workspace
enum { ST_STOP = 0, ST_ITERS = 1, ST_SCHED = 2, ST_WORDS = 8 };

struct Params {
  const double* bA;
  const double* eta;
  double* alpha;
  double* u;
  double* z[2];        // z[p] is the input of an iteration whose global index has parity p
  const double* Qt;
  long long ld, ldq;
  int n, K, n_main, k_main;
  double rho;
  int add_bias, positive, n_iters, want_norms;
  int RT, CG, U;       // row tiles, column groups, units per column group (RT + 1 if there are tail rows)
  int* state;          // ST_WORDS ints
  unsigned int* done;  // CG counters
  double* partials;    // [CG * U][8 warps][4]
  unsigned int* fault;
};

struct Sums {
  double r, d, a, u;
};

// One element of the prox / dual update.  `ao` is only meaningful when want_norms.
__device__ __forceinline__ void admm_update(const Params& p, int row, double xacc, double b, double uo, double ao,
                                            double et, double& an, double& un, bool want_norms, Sums& s) {
  const double x = b + xacc;
  const double v = x + uo;
  an = fmax(v - et, 0.0);
  if (!p.positive) an -= fmax(-v - et, 0.0);
  if (p.add_bias && row == 0) an = v;
  un = uo + x - an;
  if (want_norms) {
    const double r = x - an, dd = p.rho * (an - ao);
    s.r = fma(r, r, s.r);
    s.d = fma(dd, dd, s.d);
    s.a = fma(an, an, s.a);
    s.u = fma(un, un, s.u);
  }
}

__device__ __forceinline__ void flush_sums(const Params& p, int unit_in_iter, Sums s) {
  s.r = warp_sum(s.r);
  s.d = warp_sum(s.d);
  s.a = warp_sum(s.a);
  s.u = warp_sum(s.u);
  if ((threadIdx.x & 31) == 0) {
    double* q = p.partials + ((long long)unit_in_iter * 8 + (threadIdx.x >> 5)) * 4;
    q[0] = s.r; q[1] = s.d; q[2] = s.a; q[3] = s.u;
  }
}

// Per-lane fragment addressing (same layout as gemm_tn_core.cuh: 128B-swizzled chunks of [32 k-rows][16 doubles]).
template <class S>
struct Lanes {
  uint32_t a_off[2][2], b_off[2][2];
  int a_blk0, b_blk0, l4, g8, wm, wn;
};
template <class S>
__device__ __forceinline__ Lanes<S> make_lanes(int warp, int lane) {
  Lanes<S> lm;
  lm.l4 = lane & 3;
  lm.g8 = lane >> 2;
  lm.wm = warp / Geo<S>::WN, lm.wn = warp % Geo<S>::WN;
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
template <class S>
__device__ __forceinline__ void frags(double (&a)[S::MI], double (&b)[S::NI], uint32_t st, const Lanes<S>& lm, int j) {
  const uint32_t rowoff = (uint32_t)((j >> 1) * 8 + 2 * lm.l4 + (j & 1)) * 128u;
#pragma unroll
  for (int i = 0; i < S::MI; ++i) {
    const uint32_t addr = st + (uint32_t)((lm.a_blk0 + i) >> 1) * CHUNK_BYTES + rowoff + lm.a_off[i & 1][j & 1];
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a[i]) : "r"(addr));
  }
#pragma unroll
  for (int i = 0; i < S::NI; ++i) {
    const uint32_t addr = st + Geo<S>::A_BYTES + (uint32_t)((lm.b_blk0 + i) >> 1) * CHUNK_BYTES + rowoff +
                          lm.b_off[i & 1][j & 1];
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b[i]) : "r"(addr));
  }
}

// Variant builds (-DIPM_LASSO_TIMING, tools/lasso_timing.py): per-CTA cycle counters of the phases of a unit --
// [0] whole kernel, [1] claim, [2] dependency wait, [3] TMA + DMMA loop, [4] rank-1 terms, [5] epilogue / tail-row unit,
// [6] publish, [7] tile units, [8] tail units.  Thread 0's view (it passes every CTA barrier).
#ifdef IPM_LASSO_TIMING
__device__ long long g_lm_t[1024 * 16];
#define LM_T0() long long lm_t = clock64(); const long long lm_t_kernel = lm_t
#define LM_MARK(slot)                                                        \
  do {                                                                       \
    if (threadIdx.x == 0) {                                                  \
      const long long now = clock64();                                       \
      g_lm_t[(blockIdx.x & 1023) * 16 + (slot)] += now - lm_t;               \
      lm_t = now;                                                            \
    }                                                                        \
  } while (0)
#define LM_COUNT(slot)                                                       \
  do {                                                                       \
    if (threadIdx.x == 0) g_lm_t[(blockIdx.x & 1023) * 16 + (slot)] += 1;    \
  } while (0)
#define LM_TOTAL()                                                                              \
  do {                                                                                          \
    if (threadIdx.x == 0) g_lm_t[(blockIdx.x & 1023) * 16] += clock64() - lm_t_kernel;          \
  } while (0)
#else
#define LM_T0() \
  do {          \
  } while (0)
#define LM_MARK(slot) \
  do {                \
  } while (0)
#define LM_COUNT(slot) \
  do {                 \
  } while (0)
#define LM_TOTAL() \
  do {             \
  } while (0)
#endif

template <class S>
__global__ void __launch_bounds__(LM_THREADS, S::MINB)
lasso_admm_multi_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmZ0,
                        const __grid_constant__ CUtensorMap tmZ1, const Params p) {
  using G = Geo<S>;
  constexpr int NST = S::NSTAGES;
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_unit;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (*(volatile const int*)(p.state + ST_STOP) != 0) return;  // the stop test of an earlier launch held
  const int iters0 = *(volatile const int*)(p.state + ST_ITERS);
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t tiles0 = smem_u32(smem);
  // tail-row units have no ring traffic: their [KPARTS][LM_TAIL_MAX][TN] partial sums alias the (idle) ring
  double* s_tail = reinterpret_cast<double*>(smem);
  static_assert(G::KPARTS * LM_TAIL_MAX * G::TN * 8 <= S::NSTAGES * G::STAGE, "tail scratch aliases the ring");
  const uint32_t full0 = tiles0 + NST * G::STAGE, empty0 = full0 + 8 * NST;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(full0 + 8 * s, G::ISSUERS);
      mbar_init(empty0 + 8 * s, CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const Lanes<S> lm = make_lanes<S>(warp, lane);
  const int per_iter = p.CG * p.U;
  const int total = p.n_iters * per_iter;
  const int ktiles = (p.k_main + BK - 1) / BK;
  const unsigned done_base = (unsigned)iters0 * (unsigned)p.U;
  unsigned int* sched = reinterpret_cast<unsigned int*>(p.state + ST_SCHED);
  uint32_t it = 0;   // k-tiles consumed through the ring (all warps)
  uint32_t pit = 0;  // k-tiles issued into the ring (all warps keep the same count)

  // (Claiming the next unit early, to hide the atomic's round trip, was measured and dropped: a reserved unit cannot be
  // taken by an idle CTA, and with the lock-step dependencies of a column group one late unit delays the whole group --
  // K = 1024: 27.9 -> 33.9 us per iteration.)
  LM_T0();
  while (true) {
    __syncthreads();  // everybody is done with s_unit and the ring / tail scratch of the previous unit
    if (tid == 0) s_unit = (int)atomicAdd(sched, 1u);
    __syncthreads();
    const int unit = s_unit;
    if (unit >= total) break;
    LM_MARK(1);
    const int li = unit / per_iter;            // iteration inside this launch
    const int rem = unit - li * per_iter;
    const int g = rem / p.U, r = rem - g * p.U;
    const int par = (iters0 + li) & 1;
    const double* z_in = p.z[par];
    double* z_out = p.z[par ^ 1];
    const bool last = li == p.n_iters - 1;
    const bool norms = last && p.want_norms;
    const bool write_alpha = li >= p.n_iters - 2;
    const int n0 = g * G::TN;
    // ---- dependency: column group g finished iteration li - 1 (units of earlier launches are complete)
    if (li > 0) {
      if (tid == 0) {
        const unsigned need = done_base + (unsigned)li * (unsigned)p.U;
        const unsigned int* flag = p.done + g;
        spin_wait(
            [&] {
              unsigned v;
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
              return (int)(v - need) >= 0;
            },
            p.fault, IPM_FAULT_LASSO);
      }
      __syncthreads();
    }
    LM_MARK(2);
    Sums sums{0.0, 0.0, 0.0, 0.0};
    if (r < p.RT) {
      // ================================ DMMA tile =========================================
      const int m0 = r * G::TM;
      const CUtensorMap* tmZ = par ? &tmZ1 : &tmZ0;
      int kt_issue = 0;
      auto issue = [&]() {
        if (kt_issue >= ktiles) return;
        if (warp < G::ISSUERS && lane == 0) {
          const uint32_t s = pit % NST;
          if (pit >= NST) mbar_wait(empty0 + 8 * s, ((pit / NST) - 1) & 1);
          const uint32_t full = full0 + 8 * s;
          mbar_expect_tx(full, (warp < G::A_CHUNKS ? CHUNK_BYTES : 0) + (warp < G::B_CHUNKS ? CHUNK_BYTES : 0));
          const uint32_t dst = tiles0 + s * G::STAGE;
          if (warp < G::A_CHUNKS) tma_load_2d(dst + warp * CHUNK_BYTES, &tmQ, m0 + warp * 16, kt_issue * BK, full);
          if (warp < G::B_CHUNKS)
            tma_load_2d(dst + G::A_BYTES + warp * CHUNK_BYTES, tmZ, n0 + warp * 16, kt_issue * BK, full);
        }
        __syncwarp();
        ++pit;
        ++kt_issue;
      };
      if (warp < G::ISSUERS && lane == 0) asm volatile("fence.proxy.async;" ::: "memory");  // z was written generically
      for (int q = 0; q < NST - 1; ++q) issue();
      double acc[S::MI][S::NI][2];
      zero_acc(acc);
      {
        uint32_t s = it % NST;
        mbar_wait(full0 + 8 * s, (it / NST) & 1);
        uint32_t st = tiles0 + s * G::STAGE;
        double a[2][S::MI], b[2][S::NI];
        frags<S>(a[0], b[0], st, lm, 0);
        for (int kt = 0; kt < ktiles; ++kt) {
          const bool has_next = kt + 1 < ktiles;
          uint32_t s_next = s, st_next = st;
#pragma unroll
          for (int j = 0; j < KGROUPS; ++j) {
            const int cur = j & 1, nxt = cur ^ 1;
            if (j == 4) issue();
            if (j < KGROUPS - 1) {
              frags<S>(a[nxt], b[nxt], st, lm, j + 1);
            } else if (has_next) {
              s_next = (it + 1) % NST;
              mbar_wait(full0 + 8 * s_next, ((it + 1) / NST) & 1);
              st_next = tiles0 + s_next * G::STAGE;
              frags<S>(a[nxt], b[nxt], st_next, lm, 0);
            }
#pragma unroll
            for (int i = 0; i < S::MI; ++i)
#pragma unroll
              for (int jn = 0; jn < S::NI; ++jn) dmma884(acc[i][jn][0], acc[i][jn][1], a[cur][i], b[cur][jn]);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty0 + 8 * s);
          ++it;
          s = s_next;
          st = st_next;
        }
      }
      const int m_base = m0 + lm.wm * (S::MI * 8), n_base = n0 + lm.wn * (S::NI * 8);
      LM_MARK(3);
      // contraction rows beyond the last full k-tile (n = 513: one row) as rank-1 terms
      for (int k = p.k_main; k < p.n; ++k) {
        double qk[S::MI], zk[S::NI][2];
#pragma unroll
        for (int i = 0; i < S::MI; ++i) {
          const int row = m_base + i * 8 + lm.g8;
          qk[i] = row < p.n_main ? __ldg(p.Qt + (long long)k * p.ldq + row) : 0.0;
        }
#pragma unroll
        for (int jn = 0; jn < S::NI; ++jn)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = n_base + jn * 8 + 2 * lm.l4 + e;
            zk[jn][e] = col < p.K ? __ldcg(z_in + (long long)k * p.ld + col) : 0.0;
          }
#pragma unroll
        for (int i = 0; i < S::MI; ++i)
#pragma unroll
          for (int jn = 0; jn < S::NI; ++jn) {
            acc[i][jn][0] = fma(qk[i], zk[jn][0], acc[i][jn][0]);
            acc[i][jn][1] = fma(qk[i], zk[jn][1], acc[i][jn][1]);
          }
      }
      LM_MARK(4);
      // ---- epilogue on the accumulators
      const bool interior = (m_base + S::MI * 8 <= p.n_main) && (n_base + S::NI * 8 <= p.K);
      if (interior) {
        double2 et[S::NI];
#pragma unroll
        for (int jn = 0; jn < S::NI; ++jn) {
          et[jn].x = __ldg(p.eta + n_base + jn * 8 + 2 * lm.l4);
          et[jn].y = __ldg(p.eta + n_base + jn * 8 + 2 * lm.l4 + 1);
        }
#pragma unroll
        for (int i = 0; i < S::MI; ++i) {
          const int row = m_base + i * 8 + lm.g8;
          double2 vb[S::NI], vu[S::NI], va[S::NI];
#pragma unroll
          for (int jn = 0; jn < S::NI; ++jn) {
            const long long idx = (long long)row * p.ld + n_base + jn * 8 + 2 * lm.l4;
            vb[jn] = __ldg(reinterpret_cast<const double2*>(p.bA + idx));
            vu[jn] = __ldcg(reinterpret_cast<const double2*>(p.u + idx));
            va[jn] = norms ? __ldcg(reinterpret_cast<const double2*>(p.alpha + idx)) : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int jn = 0; jn < S::NI; ++jn) {
            const long long idx = (long long)row * p.ld + n_base + jn * 8 + 2 * lm.l4;
            double2 an, un;
            admm_update(p, row, acc[i][jn][0], vb[jn].x, vu[jn].x, va[jn].x, et[jn].x, an.x, un.x, norms, sums);
            admm_update(p, row, acc[i][jn][1], vb[jn].y, vu[jn].y, va[jn].y, et[jn].y, an.y, un.y, norms, sums);
            if (write_alpha) __stcg(reinterpret_cast<double2*>(p.alpha + idx), an);
            __stcg(reinterpret_cast<double2*>(p.u + idx), un);
            __stcg(reinterpret_cast<double2*>(z_out + idx), make_double2(un.x - an.x, un.y - an.y));
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < S::MI; ++i) {
          const int row = m_base + i * 8 + lm.g8;
#pragma unroll
          for (int jn = 0; jn < S::NI; ++jn)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = n_base + jn * 8 + 2 * lm.l4 + e;
              if (row < p.n_main && col < p.K) {
                const long long idx = (long long)row * p.ld + col;
                double an, un;
                admm_update(p, row, acc[i][jn][e], __ldg(p.bA + idx), __ldcg(p.u + idx),
                            norms ? __ldcg(p.alpha + idx) : 0.0, __ldg(p.eta + col), an, un, norms, sums);
                if (write_alpha) __stcg(p.alpha + idx, an);
                __stcg(p.u + idx, un);
                __stcg(z_out + idx, un - an);
              }
            }
        }
      }
    } else {
      // ================================ tail rows [n_main, n) as dot products =====================
      // thread (part, c): partial sum over the k-range `part` for column n0 + c; the parts are added in a fixed order
      const int nt = p.n - p.n_main;
      const int c = tid % G::TN, part = tid / G::TN;
      const int col = n0 + c;
      const int klen = (p.n + G::KPARTS - 1) / G::KPARTS;
      const int kb = part * klen, ke = min(p.n, kb + klen);
      double accq[LM_TAIL_MAX];
#pragma unroll
      for (int q = 0; q < LM_TAIL_MAX; ++q) accq[q] = 0.0;
      if (col < p.K) {
        int k = kb;
        for (; k + 7 < ke; k += 8) {
          double zk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) zk[j] = __ldcg(z_in + (long long)(k + j) * p.ld + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const double* qrow = p.Qt + (long long)(k + j) * p.ldq + p.n_main;
#pragma unroll
            for (int q = 0; q < LM_TAIL_MAX; ++q)
              if (q < nt) accq[q] = fma(__ldg(qrow + q), zk[j], accq[q]);
          }
        }
        for (; k < ke; ++k) {
          const double zk = __ldcg(z_in + (long long)k * p.ld + col);
          const double* qrow = p.Qt + (long long)k * p.ldq + p.n_main;
#pragma unroll
          for (int q = 0; q < LM_TAIL_MAX; ++q)
            if (q < nt) accq[q] = fma(__ldg(qrow + q), zk, accq[q]);
        }
      }
#pragma unroll
      for (int q = 0; q < LM_TAIL_MAX; ++q) s_tail[(part * LM_TAIL_MAX + q) * G::TN + c] = accq[q];
      __syncthreads();
      if (part == 0 && col < p.K) {
        const double et = __ldg(p.eta + col);
        for (int q = 0; q < nt; ++q) {
          double xacc = 0.0;
          for (int pp = 0; pp < G::KPARTS; ++pp) xacc += s_tail[(pp * LM_TAIL_MAX + q) * G::TN + c];
          const int row = p.n_main + q;
          const long long idx = (long long)row * p.ld + col;
          double an, un;
          admm_update(p, row, xacc, __ldg(p.bA + idx), __ldcg(p.u + idx), norms ? __ldcg(p.alpha + idx) : 0.0, et, an,
                      un, norms, sums);
          if (write_alpha) __stcg(p.alpha + idx, an);
          __stcg(p.u + idx, un);
          __stcg(z_out + idx, un - an);
        }
      }
    }
    if (norms) flush_sums(p, rem, sums);
    LM_MARK(5);
    // ---- publish: this unit's z+ / u+ -> device scope (and the async proxy of the CTAs that will TMA-load z+)
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.done + g) : "memory");
    LM_MARK(6);
    LM_COUNT(r < p.RT ? 7 : 8);
  }
  LM_TOTAL();
}

// End of a launch: add the partial sums in slot order, evaluate the stop test (LassoSolver.py:284-298), advance the
// iteration count, rewind the work counter.  No-op once the stop flag is up.
__global__ void __launch_bounds__(256)
lasso_batch_end_kernel(int* __restrict__ state, const double* __restrict__ partials, int slots, int n_iters,
                       int want_norms, double stop_mult, double eps_rel, double rho, double* __restrict__ norms_out) {
  __shared__ double red[32];
  if (state[ST_STOP] != 0) return;
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  if (want_norms) {
    for (int i = threadIdx.x; i < slots; i += blockDim.x)
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q] += partials[(long long)i * 4 + q];
  }
  double tot[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) tot[q] = block_sum(a[q], red);
  if (threadIdx.x == 0) {
    state[ST_ITERS] += n_iters;
    state[ST_SCHED] = 0;
    if (want_norms) {
      const double r_norm = sqrt(tot[0]), d_norm = sqrt(tot[1]), a_norm = sqrt(tot[2]), u_norm = sqrt(tot[3]);
#pragma unroll
      for (int q = 0; q < 4; ++q) norms_out[q] = tot[q];
      const double tol_primal = stop_mult + eps_rel * a_norm;
      const double tol_dual = stop_mult + eps_rel * rho * u_norm;
      if (r_norm < tol_primal && d_norm < tol_dual) state[ST_STOP] = 1;
    }
  }
}

static inline int n_main_of(int n, int TM) {
  const int tail = n % TM;
  return (n >= TM && tail > 0 && tail <= LM_TAIL_MAX) ? n - tail : n;
}

struct Layout {
  bool small;
  int TM, TN, RT, CG, U, n_main;
  long long off_done, off_norms, off_partials, bytes;
};

static Layout layout_for(int n, int K, int sms) {
  Layout L;
  // 128 x 128 tiles while they still fill most of the machine; otherwise the 64 x 32 tile (16x as many units)
  const int rt_big = ceil_div(n_main_of(n, 128), 128), cg_big = ceil_div(K, 128);
  L.small = (long long)rt_big * cg_big * 20 < (long long)sms * 17;
  L.TM = L.small ? 64 : 128;
  L.TN = L.small ? 32 : 128;
  L.n_main = n_main_of(n, L.TM);
  L.RT = ceil_div(L.n_main, L.TM);
  L.CG = ceil_div(K, L.TN);
  L.U = L.RT + (L.n_main < n ? 1 : 0);
  L.off_done = ST_WORDS * 4;
  L.off_norms = (L.off_done + 4ll * L.CG + 7) / 8 * 8;
  L.off_partials = L.off_norms + 4 * 8;
  L.bytes = L.off_partials + (long long)L.CG * L.U * 8 * 4 * 8;
  return L;
}

}  // namespace lasso
}  // namespace ipm

using namespace ipm;

static int g_lm_sms[kMaxDevices];

static int lm_num_sms(int* out) {
  int dev = 0;
  IPM_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return IPM_ERR_ARG;
  if (!g_lm_sms[dev]) IPM_CUDA_CHECK(cudaDeviceGetAttribute(&g_lm_sms[dev], cudaDevAttrMultiProcessorCount, dev));
  *out = g_lm_sms[dev];
  return IPM_OK;
}

template <class S>
static int launch_multi(const CUtensorMap& tmQ, const CUtensorMap& tmZ0, const CUtensorMap& tmZ1, const lasso::Params& p,
                        int sms, long long per_iter, long long units, cudaStream_t st) {
  auto kern = lasso::lasso_admm_multi_kernel<S>;
  static bool attr_set[kMaxDevices];
  static int per_sm[kMaxDevices];
  IPM_CUDA_CHECK(ensure_dynamic_smem(kern, lasso::Geo<S>::SMEM, attr_set));
  int dev = 0;
  IPM_CUDA_CHECK(cudaGetDevice(&dev));
  if (!per_sm[dev])
    IPM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[dev], kern, lasso::LM_THREADS,
                                                                 lasso::Geo<S>::SMEM));
  if (per_sm[dev] < 1) return IPM_ERR_ARG;
  int ctas_per_sm = per_sm[dev] < S::MINB ? per_sm[dev] : S::MINB;
  if (per_iter <= sms) ctas_per_sm = 1;
  long long grid = (long long)sms * ctas_per_sm;
  if (grid > units) grid = units;
  IPM_CUDA_CHECK(launch_cooperative(kern, dim3((unsigned)grid), dim3(lasso::LM_THREADS), lasso::Geo<S>::SMEM, st, tmQ,
                                    tmZ0, tmZ1, p));
  return IPM_OK;
}

// Bytes of the workspace `ws` of ipm_lasso_admm_steps_f64 (zero it before the first launch of a solve):
// [stop, iterations done, work counter, ...] | done[column groups] | norms[4] | partial sums.
extern "C" long long ipm_lasso_steps_ws_bytes(int n, int K) {
  if (n <= 0 || K <= 0) return 0;
  // the layout depends on the SM count only through the tile choice; reserve for the smaller tile (more units)
  lasso::Layout a = lasso::layout_for(n, K, 1), b = lasso::layout_for(n, K, 1 << 20);
  return a.bytes > b.bytes ? a.bytes : b.bytes;
}

// `n_iters` ADMM iterations for all K problems in ONE launch (+ a one-CTA kernel that ends the batch).
//   Qt: n x n (ldq), symmetric;  bA, alpha, u, z0, z1: n x K (ld);  z0 holds z = u - alpha when the iteration count in
//   ws is even, z1 when it is odd (after `it` iterations in total the current z is in z[it & 1]).
//   want_norms: evaluate the reference's stop test after the last iteration of the launch:
//       ||x - alpha+|| < stop_mult + eps_rel ||alpha+||   and   ||rho (alpha+ - alpha)|| < stop_mult + eps_rel rho ||u+||
//   (Frobenius norms over the whole batch, stop_mult = eps_abs * sqrt(n K)).  ws (device, ipm_lasso_steps_ws_bytes,
//   zeroed by the caller at the start of a solve): ((int*)ws)[0] = 1 once the test held -- later launches are then
//   no-ops --, ((int*)ws)[1] = iterations performed so far; the four squared norms of the last test are at
//   ipm_lasso_steps_norms_offset() bytes.  alpha is valid after every launch.
extern "C" int ipm_lasso_admm_steps_f64(const double* Qt, int ldq, int n, int K, const double* bA, const double* eta,
                                        double rho, double* alpha, double* u, double* z0, double* z1, int ld,
                                        int add_bias, int positive, int n_iters, int want_norms, double stop_mult,
                                        double eps_rel, void* ws, void* stream) {
  if (!Qt || !bA || !eta || !alpha || !u || !z0 || !z1 || z0 == z1 || !ws || n <= 0 || K <= 0 || ldq < n || ld < K ||
      (ld & 1) || (ldq & 1) || n_iters < 1)
    return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 0;
  int rc = lm_num_sms(&sms);
  if (rc) return rc;
  const lasso::Layout L = lasso::layout_for(n, K, sms);
  if ((long long)n_iters * L.CG * L.U > 0x7fffffffll) return IPM_ERR_ARG;
  CUtensorMap tmQ, tmZ0, tmZ1;
  rc = make_operand_map(&tmQ, Qt, ldq, n, n);  // Q~ symmetric: Q~[k][i] is the A operand (k rows, i columns)
  if (rc) return rc;
  rc = make_operand_map(&tmZ0, z0, ld, n, K);
  if (rc) return rc;
  rc = make_operand_map(&tmZ1, z1, ld, n, K);
  if (rc) return rc;
  const int k_tail = n % gemm::BK;
  const int k_main = (n > gemm::BK && k_tail > 0 && k_tail <= 4) ? n - k_tail : n;
  uint8_t* wsb = (uint8_t*)ws;
  lasso::Params p;
  p.bA = bA, p.eta = eta, p.alpha = alpha, p.u = u, p.z[0] = z0, p.z[1] = z1, p.Qt = Qt;
  p.ld = ld, p.ldq = ldq, p.n = n, p.K = K, p.n_main = L.n_main, p.k_main = k_main, p.rho = rho;
  p.add_bias = add_bias, p.positive = positive, p.n_iters = n_iters, p.want_norms = want_norms;
  p.RT = L.RT, p.CG = L.CG, p.U = L.U;
  p.state = (int*)wsb;
  p.done = (unsigned int*)(wsb + L.off_done);
  p.partials = (double*)(wsb + L.off_partials);
  p.fault = ipm_internal_fault_word();
  const long long units = (long long)n_iters * L.CG * L.U;
  // grid: every CTA resident (cooperative launch).  One CTA per SM unless an iteration has more units than SMs and the
  // shape allows two: with fewer units than SMs a second CTA per SM would let two units share an SM while others idle.
  const long long per_iter = (long long)L.CG * L.U;
  if (L.small) {
    rc = launch_multi<lasso::ShapeSmall>(tmQ, tmZ0, tmZ1, p, sms, per_iter, units, st);
  } else {
    rc = launch_multi<lasso::ShapeBig>(tmQ, tmZ0, tmZ1, p, sms, per_iter, units, st);
  }
  if (rc) return rc;
  IPM_LAUNCH_CHECK();
  lasso::lasso_batch_end_kernel<<<1, 256, 0, st>>>(p.state, p.partials, L.CG * L.U * 8, n_iters, want_norms, stop_mult,
                                                   eps_rel, rho, (double*)(wsb + L.off_norms));
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

#ifdef IPM_LASSO_TIMING
// variant builds only: copies the per-CTA cycle counters to the host and clears them
extern "C" int ipm_internal_lasso_timing(long long* out, int count) {
  static long long zero[1024 * 16];
  if (count > 1024 * 16) count = 1024 * 16;
  IPM_CUDA_CHECK(cudaDeviceSynchronize());
  IPM_CUDA_CHECK(cudaMemcpyFromSymbol(out, lasso::g_lm_t, count * sizeof(long long)));
  IPM_CUDA_CHECK(cudaMemcpyToSymbol(lasso::g_lm_t, zero, sizeof(zero)));
  return IPM_OK;
}
#endif

// Byte offset inside `ws` of the four squared norms {|x - alpha+|^2, |rho (alpha+ - alpha)|^2, |alpha+|^2, |u+|^2} of the
// last stop test (for the device the caller is on).
extern "C" long long ipm_lasso_steps_norms_offset(int n, int K) {
  int sms = 0;
  if (n <= 0 || K <= 0 || lm_num_sms(&sms)) return -1;
  return lasso::layout_for(n, K, sms).off_norms;
}
