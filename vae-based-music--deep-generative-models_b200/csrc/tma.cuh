// tma.cuh — Tensor Memory Accelerator plumbing for [B, L, 32] fp32 activations (channels-last rows of 128 bytes).
// Host: a 3-D tensor map {32 channels, L rows, B items} whose box is {32, rows, 1} with the 128-byte swizzle, so a tile of
// rows moves between global memory and a dense shared-memory image with ONE instruction issued by ONE thread; rows outside
// [0, L) are clipped on stores and zero-filled on loads by the hardware (Keras' SAME padding for free).
// Device: bulk-tensor store / load, bulk async-group commit / wait, and the address of a 16-byte chunk in the swizzled
// image (chunk index XOR row % 8: a warp whose lanes own consecutive rows writes all 32 banks, no conflicts).
// The driver entry point is resolved at run time (cudaGetDriverEntryPoint): libvqvae_b200.so does not link libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc.cuh"

namespace vqb {
namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}
// tensor map of a [B, L, 32] fp32 tensor with box {32, box_rows, 1}; false if the driver refuses
static inline bool make_rows_map(CUtensorMap* m, const float* base, int B, int L, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn || !base) return false;
  const cuuint64_t dims[3] = {32, (cuuint64_t)L, (cuuint64_t)B};
  const cuuint64_t strides[2] = {128, (cuuint64_t)L * 128};
  const cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// byte offset of 16-byte chunk `c` (0..7) of row `j` in a 1024-byte-aligned, 128-byte-swizzled image of 128-byte rows
__device__ __forceinline__ uint32_t swz(int j, int c) { return (uint32_t)(j * 128 + ((c ^ (j & 7)) << 4)); }

// shared -> global, rows [c1, c1 + box_rows) of item c2 (rows >= L are not written); joins the thread's current bulk group
__device__ __forceinline__ void store_rows(const CUtensorMap* tm, const void* smem_src, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm),
               "r"(tc::smem_u32(smem_src)), "r"(0), "r"(c1), "r"(c2)
               : "memory");
}
// global -> shared (rows outside [0, L) arrive as zeros); completion is signalled on `bar` as box_rows * 128 bytes
__device__ __forceinline__ void load_rows(const CUtensorMap* tm, void* smem_dst, uint64_t* bar, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          tc::smem_u32(smem_dst)),
      "l"(tm), "r"(tc::smem_u32(bar)), "r"(0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the issuing thread's bulk stores have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... all but the most recently committed group
__device__ __forceinline__ void wait_read_but_last() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// ... have completed (required before the CTA exits)
__device__ __forceinline__ void wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

}  // namespace tma
}  // namespace vqb
