// TMA (cp.async.bulk.tensor) tile loads for the fused multigrid kernels: one elected
// thread issues one instruction per tile, the copy engine writes the tile into shared
// memory (zero-filling what lies outside the array) and signals an mbarrier -- instead
// of one LDGSTS per 8 bytes through the SM's memory-instruction queue.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace f2d {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned phase) {
  unsigned ok;
  do {
    asm volatile(
        "{\n"
        "  .reg .pred p;\n"
        "  mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "  selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  } while (!ok);
}
// descriptor fetch ahead of the first copy that uses it (a kernel parameter: legal before
// griddepcontrol.wait, the tensor map is not written by any kernel)
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)tm) : "memory");
}
// box whose first element is (row y, column x) of the 2-D array described by tm -> dst.
// x * 8 bytes must be a multiple of 16 (x even): an odd innermost coordinate raises
// "illegal instruction" on sm_100 (measured, tools/tma_probe.cu); y is unconstrained.
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, uint64_t *bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((unsigned long long)tm), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}

// tensor map of a row-major [ny][nx] fp64 array with a boxh x boxw box (no swizzle;
// out-of-bounds elements read as zero).  Requirements (cuda.h, cuTensorMapEncodeTiled):
// base 16-byte aligned, nx*8 a multiple of 16, boxw even, box sides <= 256.
inline int make_tmap_2d(CUtensorMap *tm, const double *base, int ny, int nx, int boxh, int boxw) {
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return 1;
    encode = (encode_fn)fn;
  }
  if (((uintptr_t)base & 15) || (nx & 1) || (boxw & 1) || boxw > 256 || boxh > 256) return 2;
  cuuint64_t dims[2] = {(cuuint64_t)nx, (cuuint64_t)ny};
  cuuint64_t strides[1] = {(cuuint64_t)nx * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)boxw, (cuuint32_t)boxh};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 3;
}

}  // namespace f2d
