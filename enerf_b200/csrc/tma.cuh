// tma.cuh -- TMA tensor maps (cp.async.bulk.tensor) for the channels-last activation tensors.
//
// Host: cuTensorMapEncodeTiled is reached through cudaGetDriverEntryPoint (the library links only
// against cudart; there is no libcuda on the build box).  Device: thin inline-PTX wrappers.
//
// Layout contract used by tc_conv2.cu: a (D,H,W,C) fp32 channels-last tensor is described as the 4-D
// tensor {C, W, H, D} (innermost first) with a box {C, IX, IY, IZ} and swizzle = C*4 bytes
// (32 / 64 / 128).  One box load therefore lands the halo tile as rows of C*4 bytes, one row per
// pixel in linear (z,y,x) halo order, 16-byte chunks XOR-swizzled by the row index -- exactly the
// K-major SWIZZLE_{32,64,128}B operand layout of tcgen05.mma with 8-row groups 8*C*4 bytes apart
// (SBO), so "pixel p" is operand row p and a filter tap is a start-address offset of whole rows.
// Out-of-bounds box elements are filled with zeros = the convolution's zero padding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc.cuh"

namespace enerf {
namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static inline CUtensorMapSwizzle swizzle_for_bytes(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
         : row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
}

// fp32 tensor of `rank` dims (innermost first: dims[0] contiguous).  strides_bytes[i] = byte stride of dim i+1.
// Returns 0 on success, the CUresult otherwise (-1: entry point unavailable).
static inline int encode_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                             const uint32_t* box, const uint32_t* elem_strides, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) gd[i] = dims[i], bx[i] = box[i], es[i] = elem_strides ? elem_strides[i] : 1;
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

// ---- device ---------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   tc::smem_u32(smem_dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void load_4d(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(tc::smem_u32(bar))
      : "memory");
}

// K-major SWIZZLED shared-memory matrix descriptor: rows of `row_bytes` (= the swizzle span), 8-row groups
// 8*row_bytes apart (SBO); layout_type 2 / 4 / 6 for 128 / 64 / 32-byte swizzle; base_offset (bits 49-51)
// re-phases the XOR pattern when the start address is not aligned to the pattern's repeat.
__device__ __forceinline__ uint64_t smem_desc_swz(uint32_t saddr, uint32_t row_bytes, uint32_t base_offset) {
  const uint64_t layout = row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(((8u * row_bytes) >> 4) & 0x3FFFu) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)(base_offset & 7u) << 49) | (layout << 61);
}

}  // namespace tma
}  // namespace enerf
