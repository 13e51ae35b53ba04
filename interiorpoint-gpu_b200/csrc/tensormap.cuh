// Host-side TMA descriptor construction (cuTensorMapEncodeTiled resolved through the runtime, so the
// library has no link-time dependency on libcuda and still loads on a CPU-only box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace ipm {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Operand X: `rows` x `cols` doubles, row-major with leading dimension ld.  Box = [box_rows][16 doubles],
// 128B swizzle; out-of-bounds elements are zero-filled, which handles ragged K / M / N for free.
inline int make_operand_map(CUtensorMap* tm, const double* base, long long ld, long long rows, long long cols,
                            int box_rows = 32) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return IPM_ERR_NO_DEVICE;
  if (((uintptr_t)base & 15) || (ld & 1)) return IPM_ERR_ARG;  // 16-byte base and row stride
  if (rows == 0) rows = 1;  // K == 0: the kernel never issues a load
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? IPM_OK : IPM_ERR_ARG;
}

}  // namespace ipm
