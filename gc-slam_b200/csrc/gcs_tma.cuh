// gcs_tma.cuh -- Tensor Memory Accelerator plumbing (sm_100a): tensor-map encoding on the host through the driver entry
// point (no link-time dependency on libcuda), bulk tensor loads global -> shared memory completing on an mbarrier.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gcs_tc.cuh"   // mbarrier helpers

namespace gcs {
namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D row-major float64 tensor (outer, inner) with a (box_outer, box_inner) box, no swizzle.  false: no TMA (the caller
// falls back to plain loads).  base must be 16-byte aligned, inner * 8 a multiple of 16, boxes <= 256 per dimension.
inline bool encode_2d_f64(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner,
                          uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn || ((uintptr_t)base & 15) || box_inner > 256 || box_outer > 256 || box_inner == 0 || box_outer == 0) return false;
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {inner * 8};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
// box at (x = inner coordinate, y = outer coordinate) -> shared memory (128-byte aligned); completes `bytes` on `bar`
__device__ __forceinline__ void load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          tc::smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(tc::smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace tma
}  // namespace gcs
