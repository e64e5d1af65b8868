// Host-side TMA tensor-map encoding.  cuTensorMapEncodeTiled is resolved through the runtime
// (cudaGetDriverEntryPoint) so libm2b200.so has no link-time dependency on libcuda.so and
// loads on a box without a driver (the C-ABI symbol test runs on CPU).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace m2 {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = []() -> PFN_encodeTiled {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<PFN_encodeTiled>(p);
  }();
  return fn;
}

// Row-major bf16 matrix [rows][cols] with leading dimension ld (elements); box = box_rows x box_cols,
// 128-byte swizzle (box_cols * 2 bytes must be <= 128), out-of-bounds elements read as zero.
inline int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                          uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return M2_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15)) return M2_ERR_ALIGN;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? M2_OK : M2_ERR_DRIVER;
}


// Output tile map for TMA stores: row-major [batch][rows][cols] (element size 2 = bf16 or 4 = fp32), leading dimension ld
// and batch stride in elements, box = 1 x box_rows x box_cols with box_cols * elem_bytes == 128 (128-byte swizzle).
// Stores clip at the tensor bounds, so ragged edge tiles need no predicates.
inline int make_tmap_store3d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t batch, uint64_t rows, uint64_t cols,
                             uint64_t ld, uint64_t batch_stride, uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return M2_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * elem_bytes) & 15) || ((batch_stride * elem_bytes) & 15)) return M2_ERR_ALIGN;
  cuuint64_t gdim[3] = {cols, rows, batch};
  cuuint64_t gstride[2] = {ld * elem_bytes, (batch > 1 ? batch_stride : rows * ld) * elem_bytes};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? M2_OK : M2_ERR_DRIVER;
}

}  // namespace m2
