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

// All three helpers use lane -> (row r0 + 8k, chunk q) with r0 = lane / 4, q = lane % 4 on the global side: the address and
// the bounds of the four accesses then differ by constants (row stride 8 * 128 B, staging stride 512 B), so each call costs
// one 64-bit address, one pair of bounds and four 1-2 instruction accesses.  (The kernels that use them are instruction-issue
// bound: written naively, the per-access index / bounds / swizzle arithmetic was a quarter of their instructions.)
// staging offset of (row r0 + 8k, chunk q) = stg_base_off(lane) + 512 k   [8k rows do not change the swizzle: (8k >> 1) & 3 = 0]
__device__ __forceinline__ uint32_t stg_base_off(int lane) { return stg_off(lane >> 2, lane & 3); }

// f <- the warp's 32 rows x 64 B (rows g0.., channels half*16..+15), coalesced; zero where the row is outside [0, L).
// Split from warp_unpack_rows so that the loads can be issued long before use.
__device__ __forceinline__ void warp_fetch_rows(const float* __restrict__ src, long batch_off, int g0, int L, int half,
                                                int lane, float4 (&f)[4]) {
  const int r0 = lane >> 2;
  const float* p = src + (batch_off + g0 + r0) * 32 + half * 16 + (lane & 3) * 4;
  const unsigned lo = (unsigned)(g0 + r0), n = (unsigned)L;  // row g0 + r0 + 8k is valid iff (unsigned)(lo + 8k) < L
#pragma unroll
  for (int k = 0; k < 4; ++k)
    f[k] = (lo + 8u * k) < n ? *reinterpret_cast<const float4*>(p + k * 8 * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
}
// v[16] <- the 16 channels of the row this thread owns (row = lane), through the warp's staging area
__device__ __forceinline__ void warp_unpack_rows(const float4 (&f)[4], uint8_t* stg, int lane, float* v) {
  uint8_t* w = stg + stg_base_off(lane);
#pragma unroll
  for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(w + 512 * k) = f[k];
  __syncwarp();
  const uint8_t* r = stg + lane * 64;
  const int x = (lane >> 1) & 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(r + ((q ^ x) << 4));
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
  __syncwarp();
}

// dst[(g0 + row) * 32 + half * 16 + ..] <- v of the thread owning `row`, for tile rows i0 + row in [own_lo, own_hi), g < L
__device__ __forceinline__ void warp_store_rows(float* __restrict__ dst, long batch_off, int g0, int L, int half, int i0,
                                                int own_lo, int own_hi, uint8_t* stg, int lane, const float* v) {
  uint8_t* w = stg + lane * 64;
  const int x = (lane >> 1) & 3;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<float4*>(w + ((q ^ x) << 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  __syncwarp();
  const int r0 = lane >> 2;
  float* p = dst + (batch_off + g0 + r0) * 32 + half * 16 + (lane & 3) * 4;
  // row r0 + 8k is stored iff own_lo <= i0 + row < own_hi and 0 <= g0 + row < L  <=>  lo <= row < hi
  const int lo = max(own_lo - i0, -g0), hi = min(own_hi - i0, L - g0);
  const uint8_t* r = stg + stg_base_off(lane);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int row = r0 + 8 * k;
    if (row >= lo && row < hi) *reinterpret_cast<float4*>(p + k * 8 * 32) = *reinterpret_cast<const float4*>(r + 512 * k);
  }
  __syncwarp();
}

}  // namespace vqb
