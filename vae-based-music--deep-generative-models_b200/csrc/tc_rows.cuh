// tc_rows.cuh — helpers shared by the tcgen05 convolution kernels (resblock_tc.cu, conv_tc.cu): L2 prefetch hints and the
// per-warp staged row I/O of the TMEM epilogues.
#pragma once
#include "tc.cuh"

namespace vqb {
using namespace tc;

// asks the memory system to bring [p, p + bytes) into L2 (no destination: a hint that hides the DRAM latency of the next tile)
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// rows [r0, r1) clipped to [0, L) of batch item b of a [B, L, 32] fp32 tensor
__device__ __forceinline__ void prefetch_rows(const float* base, long b, int L, int r0, int r1) {
  r0 = max(r0, 0); r1 = min(r1, L);
  if (base && r1 > r0) prefetch_l2(base + ((size_t)b * L + r0) * 32, (uint32_t)(r1 - r0) * 128u);
}

// ---- warp-cooperative row I/O --------------------------------------------------------------------------------------
// In the epilogues thread `lane` of a warp owns tile row i0 + lane (its TMEM lane).  Reading / writing its 64-byte half
// row straight from global memory would make every warp instruction touch 32 different lines, so rows move through a
// per-warp 2 KB staging area (32 rows x 64 B, 16-byte chunks XOR-swizzled by (row >> 1) & 3: conflict-free both ways)
// and the global side is done with lane -> (row = lane / 4, chunk = lane % 4): 64 contiguous bytes per row.
__device__ __forceinline__ uint32_t stg_off(int row, int q) { return (uint32_t)(row * 64 + ((q ^ ((row >> 1) & 3)) << 4)); }

// f <- the warp's 32 rows x 64 B (rows g0.., channels half*16..+15), coalesced: lane -> (row = idx / 4, chunk = idx % 4);
// zero where the row is outside [0, L).  Split from warp_unpack_rows so that the loads can be issued long before use.
__device__ __forceinline__ void warp_fetch_rows(const float* __restrict__ src, long batch_off, int g0, int L, int half,
                                                int lane, float4 (&f)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = k * 32 + lane, row = idx >> 2, q = idx & 3;
    const int g = g0 + row;
    f[k] = (g >= 0 && g < L) ? *reinterpret_cast<const float4*>(src + (batch_off + g) * 32 + half * 16 + q * 4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// v[16] <- the 16 channels of the row this thread owns (row = lane), through the warp's staging area
__device__ __forceinline__ void warp_unpack_rows(const float4 (&f)[4], uint8_t* stg, int lane, float* v) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = k * 32 + lane, row = idx >> 2, q = idx & 3;
    *reinterpret_cast<float4*>(stg + stg_off(row, q)) = f[k];
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(stg + stg_off(lane, q));
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
  __syncwarp();
}

// dst[(g0 + row) * 32 + half * 16 + ..] <- v of the thread owning `row`, for tile rows i0 + row in [own_lo, own_hi), g < L
__device__ __forceinline__ void warp_store_rows(float* __restrict__ dst, long batch_off, int g0, int L, int half, int i0,
                                                int own_lo, int own_hi, uint8_t* stg, int lane, const float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<float4*>(stg + stg_off(lane, q)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = k * 32 + lane, row = idx >> 2, q = idx & 3;
    const int g = g0 + row, i = i0 + row;
    if (i >= own_lo && i < own_hi && g >= 0 && g < L)
      *reinterpret_cast<float4*>(dst + (batch_off + g) * 32 + half * 16 + q * 4) =
          *reinterpret_cast<const float4*>(stg + stg_off(row, q));
  }
  __syncwarp();
}

}  // namespace vqb
