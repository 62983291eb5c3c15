// Tensor-map (TMA descriptor) helpers shared by the kernels that stage tiles with cp.async.bulk.tensor.
#pragma once
#include <cuda.h>   // CUtensorMap + the cuTensorMapEncodeTiled prototype (resolved at run time, no link dependency)
#include <cuda_runtime.h>
#include <stdint.h>

namespace kt {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// driver entry point through the runtime: libkantts has no link-time libcuda dependency
inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

namespace tc {

// one 5-D box global -> shared (SASS UTMALDG.5D), completion (box bytes) on an mbarrier; out-of-range elements arrive as zeros
__device__ __forceinline__ void tma_load_5d(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          (uint32_t)__cvta_generic_to_shared(dst_smem)),
      "l"(map), "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// one 3-D box global -> shared
__device__ __forceinline__ void tma_load_3d(void* dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(dst_smem)),
               "l"(map), "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// one 3-D box shared -> global (SASS UTMASTG), bulk async-group completion; elements outside the tensor are not written
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"((uint32_t)__cvta_generic_to_shared(src_smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the N most recent bulk groups of this thread have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

}  // namespace tc
}  // namespace kt
