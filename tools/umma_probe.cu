// Probe of the INT8 UMMA on sm_100a (tools only): what bounds tcgen05.mma.kind::i8 when operands come from shared memory
// (SS) or A comes from tensor memory (TS), and is the TMEM layout of an INT8 A operand "lane = row, 32-bit column c =
// k 4c..4c+3"?
//
//   umma_probe_rate(mode, N, iters, out)   1 CTA per SM; one thread issues `iters` MMAs (M128 x N x K32) back to back on
//                                          resident operands; out[cta] = SM cycles per MMA.  mode 0: SS, 1: TS.
//   umma_probe_ts(A, B, D)                 one TS MMA: A[128][32] int8 stored to TMEM with tcgen05.st.32x32b.x8 (thread =
//                                          row), B[64][32] int8 placed in shared memory in the 128-byte-swizzle K-major
//                                          layout, D[128][64] int32 read back with tcgen05.ld.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {
constexpr uint64_t DESC = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);  // K-major SW128, SBO 1024 B
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (long long spins = 0; spins < (1ll << 24); ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) { return DESC | (uint64_t)((addr & 0x3FFFFu) >> 4); }

__device__ __forceinline__ uint32_t tmem_setup(uint32_t tptr, int warp) {
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
  return tmem;
}
__device__ __forceinline__ void tmem_free(uint32_t tmem, int warp) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// mode 0: SS (A = 16 KB slot, B = 64 KB stage, both in smem), 1: TS (A = TMEM columns 256.., B in smem)
// split > 0: alternate between two accumulator ranges so that consecutive MMAs are independent
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int N, int iters, int ksteps, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smB = base, smA = base + 65536, bar = smA + 16384, tptr = bar + 8;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (65536 + 16384) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x01010101u * (i & 3);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t tmem = tmem_setup(tptr, warp);
  if (threadIdx.x == 0) {
    const uint32_t idesc = IDESC | ((uint32_t)(N >> 3) << 17);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int j = it % ksteps;  // k-step inside a 128-byte row
      const uint64_t db = smem_desc(smB + j * 32);
      if (mode == 0)
        umma_ss(tmem, smem_desc(smA + j * 32), db, idesc, 1u);
      else
        umma_ts(tmem, tmem + 256 + 8 * j, db, idesc, 1u);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    out[blockIdx.x] = (float)(clock64() - t0) / iters;
  }
  tmem_free(tmem, warp);
}

__global__ void __launch_bounds__(128, 1) ts_kernel(const int8_t* A, const int8_t* B, int* D, int use_ss) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t smB = base, smA = base + 8192, bar = base + 8192 + 16384, tptr = bar + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
  // K-major, 128-byte rows, 128-byte swizzle: 16-byte chunk c of row r sits at r*128 + ((c ^ (r & 7)) * 16)
  for (int i = threadIdx.x; i < 64 * 2; i += blockDim.x) {
    const int r = i >> 1, c = i & 1;
    *reinterpret_cast<uint4*>(sm + r * 128 + ((c ^ (r & 7)) * 16)) = *reinterpret_cast<const uint4*>(B + r * 32 + c * 16);
  }
  for (int i = threadIdx.x; i < 128 * 2; i += blockDim.x) {
    const int r = i >> 1, c = i & 1;
    *reinterpret_cast<uint4*>(sm + 8192 + r * 128 + ((c ^ (r & 7)) * 16)) =
        *reinterpret_cast<const uint4*>(A + r * 32 + c * 16);
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t tmem = tmem_setup(tptr, warp);
  // A -> TMEM columns 256..263: thread = row (TMEM lane 32 * warp + lane), register c = bytes 4c..4c+3 of the row
  uint32_t a[8];
  for (int c = 0; c < 8; ++c) a[c] = *reinterpret_cast<const uint32_t*>(A + row * 32 + 4 * c);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                   tmem + ((uint32_t)(warp * 32) << 16) + 256),
               "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t idesc = IDESC | ((uint32_t)(64 >> 3) << 17);
    if (use_ss)
      umma_ss(tmem, smem_desc(smA), smem_desc(smB), idesc, 0u);
    else
      umma_ts(tmem, tmem + 256, smem_desc(smB), idesc, 0u);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 64; c0 += 8) {
    int v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 8; ++c) D[row * 64 + c0 + c] = v[c];
  }
  tmem_free(tmem, warp);
  (void)lane;
}
}  // namespace

extern "C" int umma_probe_rate(int mode, int N, int iters, int ksteps, float* out, int ctas) {
  const int smem = 65536 + 16384 + 1024 + 64;
  if (cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
  rate_kernel<<<ctas, 128, smem>>>(mode, N, iters, ksteps, out);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}

extern "C" int umma_probe_ts(const int8_t* A, const int8_t* B, int* D, int use_ss) {
  const int smem = 8192 + 16384 + 1024 + 64;
  if (cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
  ts_kernel<<<1, 128, smem>>>(A, B, D, use_ss);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}
