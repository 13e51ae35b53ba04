// ipm_gemm_tn_f64: D = beta*D + alpha * A^T diag(w) B   (FP64, TMA + DMMA, see gemm_tn_core.cuh)
#include <atomic>
#include <mutex>

#include "common.cuh"
#include "gemm_tn_core.cuh"
#include "tensormap.cuh"

namespace ipm {

struct PlainEpilogue {
  double* D;
  long long ldd;
  int M, N;
  double alpha, beta;
  int tri;  // != 0: store only col >= row (upper triangle of the output block)

  template <int MI, int NI>
  __device__ __forceinline__ void tile(const double (&acc)[MI][NI][2], int m_base, int n_base, int g8, int l4) const {
    static_assert(MI % 2 == 0, "rows are processed two 8-row groups at a time");
    const bool vec_ok = ((ldd & 1) == 0) && ((((uintptr_t)D) & 15) == 0);
    // fast path: the warp tile is fully inside the matrix, fully on/above the diagonal, 16-byte aligned
    const bool interior = vec_ok && (m_base + MI * 8 <= M) && (n_base + NI * 8 <= N) &&
                          (!tri || n_base >= m_base + MI * 8 - 1);
    if (interior) {
#pragma unroll
      for (int ib = 0; ib < MI; ib += 2) {
        double2 old[2][NI];
        if (beta != 0.0) {
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int jn = 0; jn < NI; ++jn)
              old[i][jn] = *reinterpret_cast<const double2*>(D + (long long)(m_base + (ib + i) * 8 + g8) * ldd +
                                                             n_base + jn * 8 + 2 * l4);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int jn = 0; jn < NI; ++jn) {
            double2 v;
            if (beta != 0.0) {
              v.x = fma(alpha, acc[ib + i][jn][0], beta * old[i][jn].x);
              v.y = fma(alpha, acc[ib + i][jn][1], beta * old[i][jn].y);
            } else {
              v.x = alpha * acc[ib + i][jn][0];
              v.y = alpha * acc[ib + i][jn][1];
            }
            *reinterpret_cast<double2*>(D + (long long)(m_base + (ib + i) * 8 + g8) * ldd + n_base + jn * 8 +
                                        2 * l4) = v;
          }
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int row = m_base + i * 8 + g8;
#pragma unroll
      for (int jn = 0; jn < NI; ++jn) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = n_base + jn * 8 + 2 * l4 + e;
          if (row < M && col < N && !(tri && col < row)) {
            double* p = D + (long long)row * ldd + col;
            *p = (beta == 0.0) ? alpha * acc[i][jn][e] : fma(alpha, acc[i][jn][e], beta * *p);
          }
        }
      }
    }
  }
  __device__ __forceinline__ void prefetch(int, int, int, int) const {}
  __device__ __forceinline__ void after_tile(int, int, int, int) const {}
  __device__ __forceinline__ void extra(int) const {}  // never launched with extra CTAs
};

}  // namespace ipm

using namespace ipm;

// Stream-K scratch of the persistent kernel: library-owned, one slot per (device, stream) that ever launched a
// long-K contraction (allocated once, never per call).  Calls on one stream serialise, so a slot is never shared
// by two kernels in flight.
namespace {
constexpr int kMaxDev = kMaxDevices, kSlotsPerDev = 4;
struct SkSlot {
  cudaStream_t stream;
  double* partials;
  unsigned int* flags;
  bool used;
};
SkSlot g_sk[kMaxDev][kSlotsPerDev];
int g_num_sms[kMaxDev];
std::atomic<unsigned> g_sk_epoch{0};
std::mutex g_sk_mutex;

// returns nullptr when every slot of the device belongs to another stream (caller falls back to one CTA per tile)
SkSlot* get_sk_slot(int dev, cudaStream_t st, int* rc) {
  std::lock_guard<std::mutex> lock(g_sk_mutex);
  *rc = IPM_OK;
  for (int i = 0; i < kSlotsPerDev; ++i)
    if (g_sk[dev][i].used && g_sk[dev][i].stream == st) return &g_sk[dev][i];
  for (int i = 0; i < kSlotsPerDev; ++i) {
    SkSlot* s = &g_sk[dev][i];
    if (s->used) continue;
    const size_t slots = (size_t)g_num_sms[dev];
    if (cudaMalloc(&s->partials, slots * gemm::BM * gemm::BN * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&s->flags, slots * sizeof(unsigned)) != cudaSuccess ||
        cudaMemset(s->flags, 0, slots * sizeof(unsigned)) != cudaSuccess) {
      *rc = ipm_set_cuda_error(cudaGetLastError());
      return nullptr;
    }
    s->stream = st;
    s->used = true;
    return s;
  }
  return nullptr;
}
}  // namespace

// Library-internal (not part of the public header): stream-K scratch of `stream` for other translation units
// (lasso.cu).  *partials / *flags are null when no slot is free.
extern "C" int ipm_internal_sk_slot(void* stream, double** partials, unsigned int** flags, unsigned int* epoch,
                                    int* num_sms) {
  int dev = 0, rc = IPM_OK;
  IPM_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) return IPM_ERR_ARG;
  if (!g_num_sms[dev]) IPM_CUDA_CHECK(cudaDeviceGetAttribute(&g_num_sms[dev], cudaDevAttrMultiProcessorCount, dev));
  SkSlot* slot = get_sk_slot(dev, (cudaStream_t)stream, &rc);
  if (rc) return rc;
  *partials = slot ? slot->partials : nullptr;
  *flags = slot ? slot->flags : nullptr;
  unsigned e = ++g_sk_epoch;
  if (e == 0) e = ++g_sk_epoch;
  *epoch = e;
  *num_sms = g_num_sms[dev];
  return IPM_OK;
}

extern "C" int ipm_gemm_tn_f64(const double* A, int lda, const double* B, int ldb, const double* w, double alpha,
                               double beta, double* D, int ldd, int M, int N, int K, int upper, void* stream) {
  if (!A || !B || !D || M <= 0 || N <= 0 || K < 0 || lda < M || ldb < N || ldd < N) return IPM_ERR_ARG;
  if (upper == 1 && M != N) return IPM_ERR_ARG;
  if (upper < 0 || upper > 2) return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmA, tmB;
  int rc = make_operand_map(&tmA, A, lda, K, M);
  if (rc) return rc;
  rc = make_operand_map(&tmB, B, ldb, K, N);
  if (rc) return rc;
  const int tm = ceil_div(M, gemm::BM), tn = ceil_div(N, gemm::BN);
  const int tri_tiles = upper == 1;  // upper == 2: all tiles, but still store only col >= row
  const int tiles = tri_tiles ? tn * (tn + 1) / 2 : tm * tn;
  const int ktiles = ceil_div(K, gemm::BK);
  PlainEpilogue epi{D, ldd, M, N, alpha, beta, upper};

  // Persistent CTAs + stream-K in the two cases where one CTA per tile leaves SMs idle:
  //   * long-K contraction with more tiles than SMs (the Hessian): no partial last wave;
  //   * fewer tiles than SMs (the Cholesky panel chain's band / next-block updates once the trailing matrix is
  //     small, the Schur complement): the k-tiles of every tile are spread over all SMs, so a 128 x rest update with
  //     K = 128 takes one k-tile per SM plus the fix-up instead of a whole tile on a quarter of the SMs.
  int dev = 0;
  IPM_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < kMaxDev && !g_num_sms[dev])
    IPM_CUDA_CHECK(cudaDeviceGetAttribute(&g_num_sms[dev], cudaDevAttrMultiProcessorCount, dev));
  const int G = (dev >= 0 && dev < kMaxDev) ? g_num_sms[dev] : 0;
  const bool many = tiles > G && ktiles >= 1024 / gemm::BK;
  const bool few = tiles < G && ktiles >= 2 && (long long)tiles * ktiles >= 8;
  if (G > 0 && (many || few)) {
    SkSlot* slot = get_sk_slot(dev, st, &rc);
    if (rc) return rc;
    if (slot) {
      const int rem = tiles % G;
      const long long U = (long long)rem * ktiles;
      long long P = many ? U / (128 / gemm::BK) : U;  // many: at least K = 128 per stream-K slice; few: one k-tile
      if (P < rem) P = rem;
      if (P > G) P = G;
      if (P < 1) P = 1;
      const int grid = many ? G : (int)P;
      unsigned epoch = ++g_sk_epoch;
      if (epoch == 0) epoch = ++g_sk_epoch;
      gemm::StreamK sk{slot->partials, slot->flags, epoch, (int)P, 0, ipm_internal_fault_word()};
      if (w) {
        auto kern = gemm::gemm_tn_persistent_kernel<true, PlainEpilogue>;
        IPM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES));
        IPM_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(gemm::THREADS), gemm::SMEM_BYTES, st, tmA, tmB, M, N, K, w,
                                  tri_tiles, epi, sk));
      } else {
        auto kern = gemm::gemm_tn_persistent_kernel<false, PlainEpilogue>;
        IPM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES));
        IPM_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(gemm::THREADS), gemm::SMEM_BYTES, st, tmA, tmB, M, N, K,
                                  (const double*)nullptr, tri_tiles, epi, sk));
      }
      IPM_LAUNCH_CHECK();
      return IPM_OK;
    }
  }
  if (w) {
    auto kern = gemm::gemm_tn_kernel<true, PlainEpilogue>;
    IPM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES));
    IPM_CUDA_CHECK(launch_pdl(kern, dim3(tiles), dim3(gemm::THREADS), gemm::SMEM_BYTES, st, tmA, tmB, M, N, K, w,
                              tri_tiles, epi));
  } else {
    auto kern = gemm::gemm_tn_kernel<false, PlainEpilogue>;
    IPM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES));
    IPM_CUDA_CHECK(launch_pdl(kern, dim3(tiles), dim3(gemm::THREADS), gemm::SMEM_BYTES, st, tmA, tmB, M, N, K,
                              (const double*)nullptr, tri_tiles, epi));
  }
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// =========================================================================================================
// Row-sharded Hessian over peer memory (BASELINE north_star: "partial A'DA per GPU + allreduce over
// NVLink before a replicated factorisation") WITHOUT a separate collective: the SYRK's epilogue is the
// reduce-scatter, a small second kernel is the reduction + all-gather.
//
//   tile t of the upper triangle is OWNED by rank t % R.  Every rank runs the persistent SYRK over its own rows; the
//   epilogue stores each finished 128x128 partial tile straight into the owner's inbox (peer memory over NVLink /
//   NVSwitch, plain 16-byte stores) and raises a system-scope flag there.  ipm_hess_reduce_bcast_f64 then, for every
//   owned tile, waits for the R flags, adds the partials in rank order (one owner, fixed order: all ranks end up with
//   bit-identical Hessians, which the replicated factorisation relies on), writes the final tile into EVERY rank's H
//   and bumps that rank's completion counter; ipm_hess_wait_f64 parks the stream until all tiles have arrived.
//   The partial pushes overlap the SYRK main loop tile by tile; nothing on this path calls NCCL.
// =========================================================================================================
namespace ipm {

constexpr int kMaxPeers = 8;
constexpr int kTileElems = gemm::BM * gemm::BN;

struct PeerPtrs {
  double* inbox[kMaxPeers];         // rank r's inbox: [src rank][slot][128 * 128]
  unsigned int* flags[kMaxPeers];   // rank r's arrival flags: [slot][src rank]
};

__device__ __forceinline__ int upper_tile_index(int ti, int tj, int T) { return ti * T - ti * (ti - 1) / 2 + (tj - ti); }

struct PeerScatterEpilogue {
  PeerPtrs peers;
  const double* base;  // optional local n x ldb matrix added to the partial (rank 0: t*P + bound diagonal), upper part
  long long ldb;
  int n, T, me, R, slots;
  double alpha;
  unsigned int epoch;

  template <int MI, int NI>
  __device__ __forceinline__ void tile(const double (&acc)[MI][NI][2], int m_base, int n_base, int g8, int l4) const {
    const int ti = m_base / gemm::BM, tj = n_base / gemm::BN;
    const int t = upper_tile_index(ti, tj, T);
    const int owner = t % R, slot = t / R;
    double* dst = peers.inbox[owner] + ((size_t)me * slots + slot) * kTileElems;
    const int r0 = m_base - ti * gemm::BM, c0 = n_base - tj * gemm::BN;  // warp offset inside the tile
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int rl = r0 + i * 8 + g8, row = ti * gemm::BM + rl;
#pragma unroll
      for (int jn = 0; jn < NI; ++jn) {
        const int cl = c0 + jn * 8 + 2 * l4, col = tj * gemm::BN + cl;
        double2 v = make_double2(alpha * acc[i][jn][0], alpha * acc[i][jn][1]);
        if (base) {
          if (row < n && col < n && col >= row) v.x += base[(long long)row * ldb + col];
          if (row < n && col + 1 < n && col + 1 >= row) v.y += base[(long long)row * ldb + col + 1];
        }
        *reinterpret_cast<double2*>(dst + rl * gemm::BN + cl) = v;
      }
    }
    // all eight warps of the CTA call tile() for the same output tile: publish once they have all stored
    __threadfence_system();
    asm volatile("bar.sync 2, 256;" ::: "memory");
    if (threadIdx.x == 0) {
      unsigned int* f = peers.flags[owner] + (size_t)slot * R + me;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
    }
  }
  __device__ __forceinline__ void extra(int) const {}
};

struct PeerHs {
  double* H[kMaxPeers];
  unsigned int* done[kMaxPeers];  // rank r's completion counter
};

// One CTA per owned tile (grid-stride): wait for the R partials, add them in rank order, write the final tile into
// every rank's H (upper part only), count it as delivered there.
// pull.inbox[0] != nullptr: the partials stay where they were computed (rank src keeps [owner][slot] tiles, written by
// ipm_hess_i8_scatter_f64) and the owner reads them over NVLink with coalesced loads; otherwise they were pushed into this
// rank's inbox ([src][slot], ipm_syrk_scatter_f64).
struct PullSrc {
  const double* inbox[kMaxPeers];
};
__global__ void __launch_bounds__(256) hess_reduce_bcast_kernel(const double* __restrict__ inbox, PullSrc pull,
                                                                const unsigned int* __restrict__ flags, PeerHs out,
                                                                long long ldh, int n, int T, int me, int R, int slots,
                                                                unsigned int epoch, const double* __restrict__ P,
                                                                long long ldp, double tP,
                                                                unsigned int* __restrict__ fault) {
  const int ntiles = T * (T + 1) / 2;
  for (int slot = blockIdx.x; slot < slots; slot += gridDim.x) {
    const int t = slot * R + me;
    if (t >= ntiles) break;
    // invert upper_tile_index
    int ti = 0, first = 0;
    while (t >= first + (T - ti)) first += T - ti, ++ti;
    const int tj = ti + (t - first);
    if (threadIdx.x < R) {
      const unsigned int* f = flags + (size_t)slot * R + threadIdx.x;
      spin_wait(
          [&] {
            unsigned v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            return v == epoch;
          },
          fault, IPM_FAULT_PEER_REDUCE);  // a late or failed peer: carry on with what is there, the host sees the fault
    }
    __syncthreads();
    // thread -> 32 double2 of the tile: idx2 = threadIdx.x + 256 q  (row = idx2 / 64, col = 2 (idx2 % 64))
    double2 sum[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) sum[q] = make_double2(0.0, 0.0);
    for (int src = 0; src < R; ++src) {
      const double2* p = reinterpret_cast<const double2*>(
          pull.inbox[0] ? pull.inbox[src] + ((size_t)me * slots + slot) * kTileElems
                        : inbox + ((size_t)src * slots + slot) * kTileElems);
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const double2 v = __ldcg(p + threadIdx.x + 256 * q);
        sum[q].x += v.x;
        sum[q].y += v.y;
      }
    }
    if (P) {  // objective curvature t * P (replicated on every rank, added once by the tile's owner)
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int idx2 = threadIdx.x + 256 * q;
        const int row = ti * gemm::BM + (idx2 >> 6), col = tj * gemm::BN + 2 * (idx2 & 63);
        if (row < n && col < n) sum[q].x = fma(tP, P[(long long)row * ldp + col], sum[q].x);
        if (row < n && col + 1 < n) sum[q].y = fma(tP, P[(long long)row * ldp + col + 1], sum[q].y);
      }
    }
    for (int dstr = 0; dstr < R; ++dstr) {
      double* Hd = out.H[(dstr + me) % R];  // stagger the destinations over the ranks
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int idx2 = threadIdx.x + 256 * q;
        const int row = ti * gemm::BM + (idx2 >> 6), col = tj * gemm::BN + 2 * (idx2 & 63);
        if (row < n) {
          double* p = Hd + (long long)row * ldh + col;
          if (col >= row && col + 1 < n) {
            *reinterpret_cast<double2*>(p) = sum[q];
          } else {
            if (col >= row && col < n) p[0] = sum[q].x;
            if (col + 1 >= row && col + 1 < n) p[1] = sum[q].y;
          }
        }
      }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < R) {
      unsigned int* d = out.done[threadIdx.x];
      asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(d) : "memory");
    }
    __syncthreads();
  }
}

__global__ void hess_wait_kernel(const unsigned int* done, unsigned int target, unsigned int* fault) {
  spin_wait(
      [&] {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(done) : "memory");
        return (int)(v - target) >= 0;  // wrap-safe "v >= target"
      },
      fault, IPM_FAULT_PEER_WAIT);
}

}  // namespace ipm

// Partial Hessian of this rank's rows, scattered tile by tile into the owners' inboxes.
//   C: K x n local rows (ldc), w[K];  base/ldb: optional local addend (upper part);  peers: R device pointers each
//   (host arrays) to every rank's inbox and flag array;  slots = ceil(#tiles / R);  epoch: same on all ranks, != 0.
extern "C" int ipm_syrk_scatter_f64(const double* Cm, int ldc, const double* w, int n, int K, double alpha,
                                    const double* base, int ldb, void* const* peer_inbox, void* const* peer_flags,
                                    int me, int R, int slots, unsigned int epoch, void* stream) {
  if (!Cm || !w || !peer_inbox || !peer_flags || n <= 0 || K <= 0 || ldc < n || R < 1 || R > kMaxPeers || me < 0 ||
      me >= R || epoch == 0)
    return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tm;
  int rc = make_operand_map(&tm, Cm, ldc, K, n);
  if (rc) return rc;
  int dev = 0;
  IPM_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) return IPM_ERR_ARG;
  if (!g_num_sms[dev]) IPM_CUDA_CHECK(cudaDeviceGetAttribute(&g_num_sms[dev], cudaDevAttrMultiProcessorCount, dev));
  const int G = g_num_sms[dev];
  SkSlot* slot = get_sk_slot(dev, st, &rc);
  if (rc) return rc;
  if (!slot) return IPM_ERR_ARG;
  const int T = ceil_div(n, gemm::BN), tiles = T * (T + 1) / 2, ktiles = ceil_div(K, gemm::BK);
  if (slots < ceil_div(tiles, R)) return IPM_ERR_ARG;
  const int rem = tiles % G;
  const long long U = (long long)rem * ktiles;
  long long P = U / (128 / gemm::BK);
  if (P < rem) P = rem;
  if (P > G) P = G;
  if (P < 1) P = 1;
  unsigned sk_epoch = ++g_sk_epoch;
  if (sk_epoch == 0) sk_epoch = ++g_sk_epoch;
  gemm::StreamK sk{slot->partials, slot->flags, sk_epoch, (int)P, 0, ipm_internal_fault_word()};
  PeerScatterEpilogue epi;
  for (int r = 0; r < kMaxPeers; ++r) {
    epi.peers.inbox[r] = r < R ? (double*)peer_inbox[r] : nullptr;
    epi.peers.flags[r] = r < R ? (unsigned int*)peer_flags[r] : nullptr;
  }
  epi.base = base, epi.ldb = ldb, epi.n = n, epi.T = T, epi.me = me, epi.R = R, epi.slots = slots, epi.alpha = alpha;
  epi.epoch = epoch;
  auto kern = gemm::gemm_tn_persistent_kernel<true, PeerScatterEpilogue>;
  IPM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES));
  kern<<<G, gemm::THREADS, gemm::SMEM_BYTES, st>>>(tm, tm, n, n, K, w, 1, epi, sk);
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

// Reduce the owned tiles and deliver them to every rank's H; then wait (on the stream) until this rank's H is complete.
//   done_target: value this rank's completion counter reaches once all tiles of this step have arrived
//   (callers add #tiles per step to a running target).  P (optional, n x ldp, replicated): t * P is added by the owner.
static int reduce_bcast(const double* inbox, void* const* peer_inbox, const unsigned int* flags, void* const* peer_H,
                        void* const* peer_done, int ldh, int n, int me, int R, int slots, unsigned int epoch,
                        unsigned int done_target, const double* P, int ldp, double tP, void* stream) {
  if ((!inbox && !peer_inbox) || !flags || !peer_H || !peer_done || n <= 0 || ldh < n || (ldh & 1) || R < 1 ||
      R > kMaxPeers || me < 0 || me >= R)
    return IPM_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  PullSrc pull = {};
  if (peer_inbox)
    for (int r = 0; r < R; ++r) pull.inbox[r] = (const double*)peer_inbox[r];
  PeerHs out;
  for (int r = 0; r < kMaxPeers; ++r) {
    out.H[r] = r < R ? (double*)peer_H[r] : nullptr;
    out.done[r] = r < R ? (unsigned int*)peer_done[r] : nullptr;
  }
  const int T = ceil_div(n, gemm::BN);
  int grid = slots < 296 ? slots : 296;
  if (grid < 1) grid = 1;
  hess_reduce_bcast_kernel<<<grid, 256, 0, st>>>(inbox, pull, flags, out, ldh, n, T, me, R, slots, epoch, P, ldp, tP,
                                                 ipm_internal_fault_word());
  IPM_LAUNCH_CHECK();
  hess_wait_kernel<<<1, 1, 0, st>>>((const unsigned int*)peer_done[me], done_target, ipm_internal_fault_word());
  IPM_LAUNCH_CHECK();
  return IPM_OK;
}

extern "C" int ipm_hess_reduce_bcast_f64(const double* inbox, const unsigned int* flags, void* const* peer_H,
                                         void* const* peer_done, int ldh, int n, int me, int R, int slots,
                                         unsigned int epoch, unsigned int done_target, const double* P, int ldp,
                                         double tP, void* stream) {
  if (!inbox) return IPM_ERR_ARG;
  return reduce_bcast(inbox, nullptr, flags, peer_H, peer_done, ldh, n, me, R, slots, epoch, done_target, P, ldp, tP,
                      stream);
}

// The same reduction when the partial tiles were left in the producers' buffers (ipm_hess_i8_scatter_f64): rank src
// keeps the tiles it computed for owner o at peer_inbox[src] + (o * slots + slot) tiles, and raised o's flag; the owner
// pulls them over NVLink (512-byte coalesced loads) instead of receiving 16-byte remote stores from the epilogue.
extern "C" int ipm_hess_reduce_bcast_pull_f64(void* const* peer_inbox, const unsigned int* flags, void* const* peer_H,
                                              void* const* peer_done, int ldh, int n, int me, int R, int slots,
                                              unsigned int epoch, unsigned int done_target, const double* P, int ldp,
                                              double tP, void* stream) {
  if (!peer_inbox) return IPM_ERR_ARG;
  return reduce_bcast(nullptr, peer_inbox, flags, peer_H, peer_done, ldh, n, me, R, slots, epoch, done_target, P, ldp, tP,
                      stream);
}
