// tc.cuh — sm_100a tensor-core plumbing used by the tcgen05 kernels: shared-memory matrix descriptors, instruction
// descriptors, TMEM allocation / loads, mbarrier and proxy fences.  Inline PTX only (no CUTLASS dependency); field
// layouts follow the PTX ISA "tcgen05 matrix / instruction descriptor" tables (cross-checked against
// cute/arch/mma_sm100_desc.hpp: SmemDescriptor, InstrDescriptor).
//
// OPERAND LAYOUT used throughout ("plane layout", no swizzle): a [rows x K] operand tile is stored as K/T planes
// (T = 16 bytes / sizeof(element): 8 bf16 or 4 tf32), plane p holding for every row the 16 bytes of elements
// [p*T, (p+1)*T):      byte address(row, p) = p * PLANE_BYTES + row * 16.
//   * as a K-major operand (rows = M or N, K = channels): core matrix = 8 rows x 16 B, contiguous 128 B;
//     SBO (next 8 rows) = 128 B, LBO (next 16-byte K chunk) = PLANE_BYTES;
//   * as an MN-major operand (M or N = channels, K = rows): core matrix = 8 rows(k) x 16 B; LBO (next 8 k) = 128 B,
//     SBO (next 16-byte chunk of M/N) = PLANE_BYTES.
// Because consecutive rows are 16 B apart in every plane, a view shifted by r rows is the same descriptor with
// start address + 16*r — which is how the dilated taps of a Conv1D are addressed without copying.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace vqb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- descriptors -----------------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);             // start address  [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // leading dimension byte offset [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;   // stride dimension byte offset  [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version = 1        [46,48)
  return d;                                           // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

enum { FMT_F16 = 0, FMT_BF16 = 1, FMT_TF32 = 2 };

// instruction descriptor (kind::f16 / kind::tf32), fp32 accumulate
__host__ __device__ constexpr uint32_t instr_desc(int fmt, int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                          // D format: F32      [4,6)
         | ((uint32_t)fmt << 7)             // A format           [7,10)
         | ((uint32_t)fmt << 10)            // B format           [10,13)
         | ((a_mn_major ? 1u : 0u) << 15)   // A major            [15]
         | ((b_mn_major ? 1u : 0u) << 16)   // B major            [16]
         | ((uint32_t)(N >> 3) << 17)       // N >> 3             [17,23)
         | ((uint32_t)(M >> 4) << 24);      // M >> 4             [24,29)
}

// One lane of a CONVERGED warp (elect.sync).  Single-thread tcgen05 work should be guarded by this rather than by a
// lane-id comparison: under `if (lane == 0)` the compiler cannot prove the descriptor operands warp-uniform and wraps every
// tcgen05.mma in a waterfall loop (ELECT + 4 x R2UR.BROADCAST + branch, ~15 instructions per MMA); with elect.sync the
// descriptors stay in uniform registers and an MMA costs 2-3 issue slots.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- MMA issue (one thread) ------------------------------------------------------------------------------
template <bool TF32>
__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ------------------------------------------------------------------------------------------------
// one full warp; ncols power of two >= 32; the base address is written to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// lane quadrant of this warp (a warp may only touch TMEM lanes 32*(warpid%4) .. +31)
__device__ __forceinline__ uint32_t tmem_lane_base() { return ((threadIdx.x >> 5) & 3u) * 32u; }

// 32 consecutive fp32 columns of this thread's lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 consecutive fp32 columns of this thread's lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier --------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Waits for the phase with the given parity.  try_wait suspends in hardware for a bounded time; a wait that lasts longer
// than ~2 s of SM clocks is a pipeline bug, and trapping turns it into a launch error instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok)
      : "r"(a), "r"(parity)
      : "memory");
  if (ok) return;
  // Not there yet: wait with a suspend-time hint, so that the thread sleeps in hardware until the phase completes (or the
  // hint expires) instead of spinning — a spinning warp takes issue slots from the warps it is waiting for (in the residual
  // stack kernel almost half of all executed instructions were such polls before the hint was added).
  // The watchdog (a wait longer than ~2 s of SM clocks is a pipeline bug: trap instead of hanging the GPU) reads the clock only
  // every 4096 polls: the poll loop itself is 3 instructions, and polling warps share issue slots with the working ones.
  long long t0 = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity), "r"(0x989680u)
        : "memory");
    if (ok) return;
    if ((++spins & 0xfffu) == 0u) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000ll) __trap();
    }
  }
}
// One non-blocking poll of the phase with the given parity (1 = complete).  The result arrives ~150 clk after issue whatever the
// barrier's state (tools/mma_probe.cu, mbarrier rows): callers put independent work between this call and the first use of its result.
__device__ __forceinline__ uint32_t mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// one arrival (release semantics: this thread's earlier shared-memory writes are visible to whoever observes the phase)
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- element packing -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// fp32 -> tf32 (round to nearest, ties away) kept in a 32-bit container
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// x = p0 + p1 + p2 with every piece a bf16 (8 mantissa bits each: together the 24 bits of an fp32)
template <int S>
__device__ __forceinline__ void split_bf16(float x, float* pc) {
  pc[0] = __bfloat162float(__float2bfloat16_rn(x));
  if (S > 1) { pc[1] = __bfloat162float(__float2bfloat16_rn(x - pc[0])); }
  if (S > 2) { pc[2] = __bfloat162float(__float2bfloat16_rn(x - pc[0] - pc[1])); }
}


// 8 consecutive fp32 channels -> S 16-byte chunks of bf16 pieces (piece s of channel c in half-word c of out[s]);
// x = p0 + p1 + p2 exactly as in split_bf16, but with the packed two-at-a-time conversion (cvt.rn.bf16x2.f32)
template <int S>
__device__ __forceinline__ void split8(const float4& a, const float4& b, uint4* out) {
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int s = 0; s < S; ++s) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pk[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
      if (s + 1 < S) {  // exact remainders (the piece is the leading 8 bits of the value)
        v[2 * i] -= __uint_as_float(pk[i] << 16);
        v[2 * i + 1] -= __uint_as_float(pk[i] & 0xffff0000u);
      }
    }
    out[s] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}


// ---- fp16x2: scaled two-piece fp16 split -------------------------------------------------------------------
// fp16 carries 11 significant bits but only a 5-bit exponent, so operands are first multiplied by a power of two that
// brings the largest magnitude of the tile / weight tensor into [2^14, 2^15) (exact; undone in the epilogue).  Then
// hi = rn_f16(x), lo = rn_f16(x - hi) reproduce x to 2^-22 relative for every element within 2^-11 of the largest
// one and to 2^-25 * 2^-15 of the largest one below that (lo becomes subnormal: absolute step 2^-25).
// pow2_scale: power of two s with maxabs * s in [2^14, 2^15); 1 for zero / subnormal / non-finite maxabs.
__device__ __forceinline__ float pow2_scale(float maxabs) {
  const int e = (int)((__float_as_uint(maxabs) >> 23) & 0xffu);
  if (e == 0 || e == 255) return 1.f;
  const int se = min(max(268 - e, 4), 250);  // biased exponent of 2^(14 - (e - 127)), kept invertible
  return __uint_as_float((uint32_t)se << 23);
}
// exact reciprocal of a pow2_scale() value
__device__ __forceinline__ float pow2_inv(float s) { return __uint_as_float((254u << 23) - __float_as_uint(s)); }
__device__ __forceinline__ uint32_t absbits(float x) { return __float_as_uint(x) & 0x7fffffffu; }

// 8 consecutive fp32 channels * scale -> two 16-byte chunks of fp16 pieces (hi, lo)
__device__ __forceinline__ void split8_f16(const float4& a, const float4& b, float scale, uint4* out) {
  const float v[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b.x * scale, b.y * scale, b.z * scale, b.w * scale};
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);  // exact remainders
    hi[i] = *reinterpret_cast<const uint32_t*>(&h);
    lo[i] = *reinterpret_cast<const uint32_t*>(&l);
  }
  out[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  out[1] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// Compiler-level fence on a register value: arithmetic on `a` cannot be scheduled above this point.  Used after an
// mbarrier wait so that work on prefetched global data is not hoisted to right behind its loads (which would expose the
// DRAM latency the prefetch is meant to hide); volatile asm statements keep their order relative to one another.
__device__ __forceinline__ void reg_fence(float4& a) { asm volatile("" : "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w)); }

}  // namespace tc
}  // namespace vqb
