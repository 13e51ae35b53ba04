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
    static_assert(MI == 8 && NI == 4, "PlainEpilogue is written for the 128x128 CTA tile");
    const bool vec_ok = ((ldd & 1) == 0) && ((((uintptr_t)D) & 15) == 0);
    // fast path: the warp tile is fully inside the matrix, fully on/above the diagonal, 16-byte aligned
    const bool interior = vec_ok && (m_base + 64 <= M) && (n_base + 32 <= N) && (!tri || n_base >= m_base + 63);
    if (interior) {
#pragma unroll
      for (int ib = 0; ib < 8; ib += 2) {
        double2 old[2][4];
        if (beta != 0.0) {
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int jn = 0; jn < 4; ++jn)
              old[i][jn] = *reinterpret_cast<const double2*>(D + (long long)(m_base + (ib + i) * 8 + g8) * ldd +
                                                             n_base + jn * 8 + 2 * l4);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int jn = 0; jn < 4; ++jn) {
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
    for (int i = 0; i < 8; ++i) {
      const int row = m_base + i * 8 + g8;
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) {
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
constexpr int kMaxDev = 16, kSlotsPerDev = 4;
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

  // Long-K contraction with more tiles than SMs: persistent CTAs + stream-K remainder (no partial last wave).
  int dev = 0;
  IPM_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < kMaxDev && !g_num_sms[dev])
    IPM_CUDA_CHECK(cudaDeviceGetAttribute(&g_num_sms[dev], cudaDevAttrMultiProcessorCount, dev));
  const int G = (dev >= 0 && dev < kMaxDev) ? g_num_sms[dev] : 0;
  if (G > 0 && ktiles >= 64 && tiles > G) {
    SkSlot* slot = get_sk_slot(dev, st, &rc);
    if (rc) return rc;
    if (slot) {
      const int rem = tiles % G;
      const long long U = (long long)rem * ktiles;
      long long P = U / 8;  // at least 8 k-tiles (K = 128) per stream-K slice
      if (P < rem) P = rem;
      if (P > G) P = G;
      if (P < 1) P = 1;
      unsigned epoch = ++g_sk_epoch;
      if (epoch == 0) epoch = ++g_sk_epoch;
      gemm::StreamK sk{slot->partials, slot->flags, epoch, (int)P};
      if (w) {
        auto kern = gemm::gemm_tn_persistent_kernel<true, PlainEpilogue>;
        IPM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES));
        kern<<<G, gemm::THREADS, gemm::SMEM_BYTES, st>>>(tmA, tmB, M, N, K, w, tri_tiles, epi, sk);
      } else {
        auto kern = gemm::gemm_tn_persistent_kernel<false, PlainEpilogue>;
        IPM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm::SMEM_BYTES));
        kern<<<G, gemm::THREADS, gemm::SMEM_BYTES, st>>>(tmA, tmB, M, N, K, nullptr, tri_tiles, epi, sk);
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
