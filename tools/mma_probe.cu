// mma_probe.cu — issue-throughput probe for tcgen05.mma (kind::f16, M = 128, K = 16) with both operands in shared memory
// in the no-swizzle "plane layout" of csrc/tc.cuh: how many SM clocks does one MMA cost as a function of N, of the operand
// majors (K-major as in resblock_tc.cu / conv_tc.cu, MN-major as in wgrad_tc.cu) and of the plane pitch?  The numbers
// feed the kernel cost models in DESIGN.md.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I<package>/csrc -Iinclude -o tools/mma_probe tools/mma_probe.cu; tools/run_profiles.sh builds it when missing and runs it on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../vae-based-music--deep-generative-models_b200/csrc/tc.cuh"

using namespace vqb::tc;

struct Probe {
  int a_mn, b_mn, N, plane_a, plane_b, nacc, reps, shift;
  int swz;     // 1: both operands K-major in the 128-byte-swizzled layout (rows of 128 B, 8-row groups of 1024 B)
  int a_tmem;  // 1: A operand read from tensor memory (.ts form: tcgen05.mma [d], [a], bdesc, idesc)
  int M;       // 0 = 128
};

// K-major operand in the SWIZZLE_128B layout: SBO = 1024 (next 8 rows), LBO unused, layout_type 2 (bits 61..63)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// A from tensor memory
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Probe p, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x;
  for (int e = tid; e < 200 * 1024 / 16; e += blockDim.x) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 32) tmem_alloc(&tslot, 512);
  if (tid == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tslot;
  if (tid == 0) {
    uint8_t* A = smem;
    uint8_t* B = smem + 100 * 1024;
    // K-major: LBO = plane pitch (next 16 bytes of K), SBO = 128 (next 8 rows);  MN-major: LBO = 128, SBO = plane pitch
    const uint32_t A1k = (smem_u32(A) + 1023u) & ~1023u, B1k = (smem_u32(B) + 1023u) & ~1023u;  // swizzle atoms: 1024-byte aligned
    const uint64_t ad0 = p.swz ? smem_desc_sw128(A1k) : p.a_mn ? smem_desc(smem_u32(A), 128, p.plane_a) : smem_desc(smem_u32(A), p.plane_a, 128);
    const uint64_t bd0 = p.swz ? smem_desc_sw128(B1k) : p.b_mn ? smem_desc(smem_u32(B), 128, p.plane_b) : smem_desc(smem_u32(B), p.plane_b, 128);
    const uint32_t idesc = instr_desc(FMT_BF16, p.M ? p.M : 128, p.N, p.a_mn != 0, p.b_mn != 0);
    if (p.a_tmem) {  // A = 128 lanes x 8 columns of tensor memory (K = 16 halves), accumulators behind it
      const uint32_t ta = tmem + 480;
      for (int i = 0; i < 16; ++i) mma_ts(tmem, ta, bd0, idesc, 1);
      commit(&bar);
      mbar_wait(&bar, 0);
      const long long t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < p.reps; i += 6) {
        mma_ts(tmem + (0 % p.nacc) * p.N, ta, bd0, idesc, 1);
        mma_ts(tmem + (1 % p.nacc) * p.N, ta, bd0, idesc, 1);
        mma_ts(tmem + (2 % p.nacc) * p.N, ta, bd0, idesc, 1);
        mma_ts(tmem + (3 % p.nacc) * p.N, ta, bd0, idesc, 1);
        mma_ts(tmem + (4 % p.nacc) * p.N, ta, bd0, idesc, 1);
        mma_ts(tmem + (5 % p.nacc) * p.N, ta, bd0, idesc, 1);
      }
      commit(&bar);
      mbar_wait(&bar, 1);
      const long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
    } else {
    // warm-up
    for (int i = 0; i < 16; ++i) mma<false>(tmem, ad0, bd0, idesc, 1);
    commit(&bar);
    mbar_wait(&bar, 0);
    // descriptors and accumulator addresses are precomputed: the timed loop is nothing but tcgen05.mma instructions
    uint64_t ad[6];
    uint32_t dst[6];
    for (int i = 0; i < 6; ++i) { ad[i] = ad0 + (uint64_t)((i % 3) * p.shift); dst[i] = tmem + (i % p.nacc) * p.N; }
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < p.reps; i += 6) {
      mma<false>(dst[0], ad[0], bd0, idesc, 1);
      mma<false>(dst[1], ad[1], bd0, idesc, 1);
      mma<false>(dst[2], ad[2], bd0, idesc, 1);
      mma<false>(dst[3], ad[3], bd0, idesc, 1);
      mma<false>(dst[4], ad[4], bd0, idesc, 1);
      mma<false>(dst[5], ad[5], bd0, idesc, 1);
    }
    commit(&bar);
    mbar_wait(&bar, 1);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  }
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 512);
}

// Is the ~50 clk per small-N MMA a property of the tensor pipe or of the issuing thread?  NI threads (lane 0 of NI different warps)
// issue reps / NI MMAs each into their own accumulators at the same time; the clock runs until all of them have completed.
__global__ void __launch_bounds__(128, 1) probe_multi_kernel(int N, int reps, int ni, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tslot;
  __shared__ long long tstart[4], tend[4];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 200 * 1024 / 16; e += blockDim.x) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 32) tmem_alloc(&tslot, 512);
  if (tid == 32) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tslot;
  const int PK = (256 + 64) * 16 + 32, PW = 96 * 16;
  if ((tid & 31) == 0 && warp < ni) {
    uint8_t* A = smem + warp * 24 * 1024;
    uint8_t* B = smem + 100 * 1024 + warp * 8 * 1024;
    const uint64_t ad0 = smem_desc(smem_u32(A), PK, 128), bd0 = smem_desc(smem_u32(B), PW, 128);
    const uint32_t idesc = instr_desc(FMT_BF16, 128, N, false, false);
    const uint32_t dst = tmem + warp * 128;
    for (int i = 0; i < 8; ++i) mma<false>(dst, ad0, bd0, idesc, 1);
    commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    tstart[warp] = clock64();
#pragma unroll 1
    for (int i = 0; i < reps / ni; i += 4) {
      mma<false>(dst, ad0, bd0, idesc, 1);
      mma<false>(dst, ad0 + 1, bd0, idesc, 1);
      mma<false>(dst, ad0 + 2, bd0, idesc, 1);
      mma<false>(dst, ad0, bd0, idesc, 1);
    }
    commit(&bar[warp]);
    mbar_wait(&bar[warp], 1);
    tend[warp] = clock64();
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) {
    long long t0 = tstart[0], t1 = tend[0];
    for (int i = 1; i < ni; ++i) { t0 = t0 < tstart[i] ? t0 : tstart[i]; t1 = t1 > tend[i] ? t1 : tend[i]; }
    out[0] = t1 - t0;
  }
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 512);
}

// Tensor-memory read rate: nw warps (lane quadrant = warp % 4) each issue reps tcgen05.ld.32x32b.x32 (4 KB per instruction: 32
// lanes x 32 columns x 4 B), waited for in groups of 4.  Bytes per clock per SM.
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
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

template <int SHAPE>
__global__ void __launch_bounds__(512, 1) probe_tmem_ld_kernel(int reps, long long* out) {
  __shared__ uint32_t tslot;
  __shared__ long long t0s[16], t1s[16];
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tslot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t taddr = tslot + ((uint32_t)(warp & 3) * 32u << 16) + (uint32_t)((warp >> 2) * 32);
  float acc = 0.f;
  float v[32];
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < reps; ++i) {
    if (SHAPE == 0) tmem_ld32(taddr, v); else tmem_ld_16x256b_x8(taddr, v);
    acc += v[i & 31];
  }
  const long long t1 = clock64();
  if ((tid & 31) == 0) { t0s[warp] = t0; t1s[warp] = t1; }
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) {
    long long a = t0s[0], b = t1s[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { a = a < t0s[w] ? a : t0s[w]; b = b > t1s[w] ? b : t1s[w]; }
    out[0] = b - a;
    out[1] = (long long)acc;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tslot, 512);
}

// Round-2 follow-up: the same read with G loads in flight per wait (no wait after every instruction, no array indexing that
// could send the registers through local memory): is 32 B/clk the pipe's rate or the latency of one load + wait?
template <int G>
__global__ void __launch_bounds__(512, 1) probe_tmem_ld_inflight_kernel(int reps, long long* out) {
  __shared__ uint32_t tslot;
  __shared__ long long t0s[16], t1s[16];
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tslot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t taddr = tslot + ((uint32_t)(warp & 3) * 32u << 16) + (uint32_t)((warp >> 2) * 128);
  uint32_t acc = 0u;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < reps; ++i) {
    uint32_t r[G][16];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
          : "=r"(r[g][0]), "=r"(r[g][1]), "=r"(r[g][2]), "=r"(r[g][3]), "=r"(r[g][4]), "=r"(r[g][5]), "=r"(r[g][6]), "=r"(r[g][7]),
            "=r"(r[g][8]), "=r"(r[g][9]), "=r"(r[g][10]), "=r"(r[g][11]), "=r"(r[g][12]), "=r"(r[g][13]), "=r"(r[g][14]), "=r"(r[g][15])
          : "r"(taddr + g * 16)
          : "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int g = 0; g < G; ++g) acc ^= r[g][0] ^ r[g][15];
  }
  const long long t1 = clock64();
  if ((tid & 31) == 0) { t0s[warp] = t0; t1s[warp] = t1; }
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) {
    long long a = t0s[0], b = t1s[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { a = a < t0s[w] ? a : t0s[w]; b = b > t1s[w] ? b : t1s[w]; }
    out[0] = b - a;
    out[1] = (long long)acc;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tslot, 512);
}

// Do tcgen05.mma and concurrent tcgen05.ld slow one another down?  Warp 0 issues `reps` MMAs (M = 128, N = 256, K-major planes of
// the given pitches) into accumulator columns [0, 256); `nld` further warps read columns [256, 512) in a loop (2 x 32 columns per
// wait) until the MMAs have completed.  Prints clk per MMA and the loaders' rate.
__global__ void __launch_bounds__(544, 1) probe_mma_ld_kernel(int plane_a, int plane_b, int reps, int nld, int same_acc, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  __shared__ volatile int stop;
  __shared__ long long nloads[16];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 200 * 1024 / 16; e += blockDim.x) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 32) tmem_alloc(&tslot, 512);
  if (tid == 32) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
  if (tid < 16) nloads[tid] = 0;
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tslot;
  if (warp == 16) {
    if (elect_one()) {
      const uint64_t ad0 = smem_desc(smem_u32(smem), plane_a, 128), bd0 = smem_desc(smem_u32(smem + 100 * 1024), plane_b, 128);
      const uint32_t idesc = instr_desc(FMT_BF16, 128, 256, false, false);
      const long long t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < reps; i += 4) {
        mma<false>(tmem, ad0, bd0, idesc, 1);
        mma<false>(tmem, ad0 + (uint64_t)((2 * plane_a) >> 4), bd0 + (uint64_t)((2 * plane_b) >> 4), idesc, 1);
        mma<false>(tmem, ad0 + (uint64_t)((4 * plane_a) >> 4), bd0 + (uint64_t)((4 * plane_b) >> 4), idesc, 1);
        mma<false>(tmem, ad0 + (uint64_t)((6 * plane_a) >> 4), bd0 + (uint64_t)((6 * plane_b) >> 4), idesc, 1);
      }
      commit(&bar);
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      stop = 1;
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncwarp();
  } else if (warp < nld) {
    const uint32_t taddr = tmem + ((uint32_t)(warp & 3) * 32u << 16) + (uint32_t)((same_acc ? 0 : 256) + (warp >> 2) * 64);
    uint32_t acc = 0u;
    long long n = 0;
    while (!stop) {
      uint32_t r[2][32];
#pragma unroll
      for (int g = 0; g < 2; ++g)
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[g][0]), "=r"(r[g][1]), "=r"(r[g][2]), "=r"(r[g][3]), "=r"(r[g][4]), "=r"(r[g][5]), "=r"(r[g][6]), "=r"(r[g][7]), "=r"(r[g][8]),
              "=r"(r[g][9]), "=r"(r[g][10]), "=r"(r[g][11]), "=r"(r[g][12]), "=r"(r[g][13]), "=r"(r[g][14]), "=r"(r[g][15]), "=r"(r[g][16]),
              "=r"(r[g][17]), "=r"(r[g][18]), "=r"(r[g][19]), "=r"(r[g][20]), "=r"(r[g][21]), "=r"(r[g][22]), "=r"(r[g][23]), "=r"(r[g][24]),
              "=r"(r[g][25]), "=r"(r[g][26]), "=r"(r[g][27]), "=r"(r[g][28]), "=r"(r[g][29]), "=r"(r[g][30]), "=r"(r[g][31])
            : "r"(taddr + g * 32)
            : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r[0][0] ^ r[1][31];
      ++n;
    }
    if ((tid & 31) == 0) nloads[warp] = n + (acc == 0x12345u);
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) {
    long long t = 0;
    for (int i = 0; i < 16; ++i) t += nloads[i];
    out[1] = t;
  }
  if (tid < 32) tmem_dealloc(tmem, 512);
}

// mbarrier costs: (a) a wait on a phase that has already completed (test_wait / try_wait / the kernels' mbar_wait), (b) the wake-up
// latency of a waiting thread after another warp's arrive, for a plain try_wait loop and for mbar_wait (try_wait with a suspend hint).
__device__ __forceinline__ uint32_t test_wait_once(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ uint32_t try_wait_once(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok;
}
__global__ void __launch_bounds__(64, 1) probe_mbar_kernel(long long* out) {
  __shared__ uint64_t bar[4];
  __shared__ volatile long long t_arrive;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  __syncthreads();
  if (tid == 0) mbar_arrive(&bar[0]);  // phase 0 of bar[0] is complete from here on
  __syncthreads();
  if (warp == 0 && lane == 0) {
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < 100; ++i) acc += test_wait_once(&bar[0], 0);
    long long t1 = clock64();
    for (int i = 0; i < 100; ++i) acc += try_wait_once(&bar[0], 0);
    long long t2 = clock64();
    for (int i = 0; i < 100; ++i) mbar_wait(&bar[0], 0);
    long long t3 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[7] = acc;
  }
  __syncthreads();
  // wake-up latency, 3 variants: warp 0 waits on bar[1 + v]; warp 1 arrives ~20000 clk later
  for (int v = 0; v < 3; ++v) {
    __syncthreads();
    if (warp == 0) {
      if (lane == 0) {
        if (v == 0) { while (!test_wait_once(&bar[1 + v], 0)) {} }
        else if (v == 1) { while (!try_wait_once(&bar[1 + v], 0)) {} }
        else mbar_wait(&bar[1 + v], 0);
        const long long t = clock64();
        out[3 + v] = t - t_arrive;
      }
      __syncwarp();
    } else {
      if (lane == 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < 20000) {}
        t_arrive = clock64();
        __threadfence_block();
        mbar_arrive(&bar[1 + v]);
      }
      __syncwarp();
    }
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 600;
  struct Case { const char* name; Probe p; };
  const int PK = (256 + 64) * 16 + 32;   // resblock_tc plane pitch (K-major A tile)
  const int PW = 96 * 16;                // its weight planes
  const int PX = (128 + 64) * 16 + 32;   // wgrad_tc act(x) plane pitch (MN-major)
  const int PD = 128 * 16 + 32;          // wgrad_tc dy plane pitch
  Case cases[] = {
      {"K-major A / K-major B  N=32  (rb bf16)", {0, 0, 32, PK, PW, 1, reps, 1}},
      {"K-major A / K-major B  N=64", {0, 0, 64, PK, PW, 1, reps, 1}},
      {"K-major A / K-major B  N=96  (rb bf16x3 sa=0)", {0, 0, 96, PK, PW, 1, reps, 1}},
      {"K-major A / K-major B  N=96, 2 accumulators", {0, 0, 96, PK, PW, 2, reps, 1}},
      {"K-major A / K-major B  N=128", {0, 0, 128, PK, 128 * 16, 1, reps, 1}},
      {"K-major A / K-major B  N=256", {0, 0, 256, PK, 256 * 16, 1, reps, 1}},
      {"K-major A / K-major B  N=256, 2 accumulators", {0, 0, 256, PK, 256 * 16, 2, reps, 1}},
      {"MN-major A / MN-major B N=32  (wgrad bf16)", {1, 1, 32, PX, PD, 3, reps, 3}},
      {"MN-major A / MN-major B N=64  (wgrad bf16x2)", {1, 1, 64, PX, PD, 3, reps, 3}},
      {"MN-major A / MN-major B N=96  (wgrad bf16x3)", {1, 1, 96, PX, PD, 3, reps, 3}},
      {"MN-major A / MN-major B N=96, pitch 2048+16", {1, 1, 96, 2048 + 16, 2048 + 16, 3, reps, 3}},
      {"MN-major A / MN-major B N=96, pitch 2048+128", {1, 1, 96, 2048 + 128, 2048 + 128, 3, reps, 3}},
      {"MN-major A / MN-major B N=96, pitch 4096", {1, 1, 96, 4096, 4096, 3, reps, 3}},
      {"MN-major A / K-major B  N=96", {1, 0, 96, PX, PW, 3, reps, 3}},
      {"K-major A / MN-major B  N=96", {0, 1, 96, PK, PD, 3, reps, 1}},
      {"MN-major A / MN-major B N=256", {1, 1, 256, PX, PD, 1, reps, 3}},
      // round 2: 128-byte-swizzled K-major operands (the TMA-native layout) ...
      {"SW128 K-major A / SW128 K-major B  N=32", {0, 0, 32, 0, 0, 1, reps, 2, 1, 0, 0}},
      {"SW128 K-major A / SW128 K-major B  N=64", {0, 0, 64, 0, 0, 1, reps, 2, 1, 0, 0}},
      {"SW128 K-major A / SW128 K-major B  N=96", {0, 0, 96, 0, 0, 1, reps, 2, 1, 0, 0}},
      {"SW128 K-major A / SW128 K-major B  N=128", {0, 0, 128, 0, 0, 1, reps, 2, 1, 0, 0}},
      {"SW128 K-major A / SW128 K-major B  N=256", {0, 0, 256, 0, 0, 1, reps, 2, 1, 0, 0}},
      // ... M = 64 (half the A bytes per instruction) ...
      {"K-major A / K-major B  M=64 N=32", {0, 0, 32, PK, PW, 1, reps, 1, 0, 0, 64}},
      {"K-major A / K-major B  M=64 N=64", {0, 0, 64, PK, PW, 1, reps, 1, 0, 0, 64}},
      {"K-major A / K-major B  M=64 N=256", {0, 0, 256, PK, 256 * 16, 1, reps, 1, 0, 0, 64}},
      // ... and the A operand in tensor memory (.ts): no shared-memory read of A at all
      {"A in TMEM / K-major B  N=32", {0, 0, 32, PK, PW, 1, reps, 1, 0, 1, 0}},
      {"A in TMEM / K-major B  N=64", {0, 0, 64, PK, PW, 1, reps, 1, 0, 1, 0}},
      {"A in TMEM / K-major B  N=96", {0, 0, 96, PK, PW, 1, reps, 1, 0, 1, 0}},
      {"A in TMEM / K-major B  N=128", {0, 0, 128, PK, 128 * 16, 1, reps, 1, 0, 1, 0}},
      {"A in TMEM / K-major B  N=256", {0, 0, 256, PK, 256 * 16, 1, reps, 1, 0, 1, 0}},
      {"A in TMEM / SW128 K-major B  N=64", {0, 0, 64, 0, 0, 1, reps, 1, 1, 1, 0}},
      {"A in TMEM / K-major B  M=64 N=64", {0, 0, 64, PK, PW, 1, reps, 1, 0, 1, 64}},
  };
  for (auto& c : cases) {
    probe_kernel<<<4, 128, 200 * 1024>>>(c.p, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    printf("%-52s %8.1f clk / MMA   (%s)\n", c.name, (double)cyc / reps, cudaGetErrorString(e));
  }
  for (int shape : {0, 1})
    for (int nw : {1, 4, 8, 16}) {
      if (shape == 0) probe_tmem_ld_kernel<0><<<4, nw * 32>>>(2000, d); else probe_tmem_ld_kernel<1><<<4, nw * 32>>>(2000, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("tcgen05.ld.%s, %2d warps x 2000 loads (4 KB each)                        %8.1f B / clk / SM   (%s)\n",
             shape ? "16x256b.x8 " : "32x32b.x32 ", nw, (double)nw * 2000 * 4096 / (double)cyc, cudaGetErrorString(e));
    }
  for (int G : {1, 2, 4, 8})
    for (int nw : {1, 4, 8, 16}) {
      if (G == 1) probe_tmem_ld_inflight_kernel<1><<<4, nw * 32>>>(2000, d);
      else if (G == 2) probe_tmem_ld_inflight_kernel<2><<<4, nw * 32>>>(2000, d);
      else if (G == 4) probe_tmem_ld_inflight_kernel<4><<<4, nw * 32>>>(2000, d);
      else probe_tmem_ld_inflight_kernel<8><<<4, nw * 32>>>(2000, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("tcgen05.ld.32x32b.x16, %d in flight per wait, %2d warps x 2000 rounds (%d KB each)   %8.1f B / clk / SM   (%s)\n", G, nw,
             2 * G, (double)nw * 2000 * 2048 * G / (double)cyc, cudaGetErrorString(e));
    }
  cudaFuncSetAttribute(probe_mma_ld_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  {
    long long* d2;
    cudaMalloc(&d2, 16);
    struct { const char* name; int pa, pb; } lay[] = {{"pitch 2048 / 4096", 2048, 4096}, {"pitch 2048+64 / 4096+64 (vq2_kernel)", 2048 + 64, 4096 + 64},
                                                       {"pitch 2048+128 / 4096+128", 2048 + 128, 4096 + 128}};
    for (auto& l : lay)
      for (int same : {0, 1})
        for (int nld : {0, 4, 8, 16}) {
          if (same && nld == 0) continue;
          probe_mma_ld_kernel<<<4, 544, 200 * 1024>>>(l.pa, l.pb, 400, nld, same, d2);
          cudaError_t e = cudaDeviceSynchronize();
          long long r[2] = {0, 0};
          cudaMemcpy(r, d2, 16, cudaMemcpyDeviceToHost);
          printf("MMA N=256 K-major, %-38s + %2d warps of tcgen05.ld on %s accumulator: %7.1f clk / MMA, loads %7.1f B / clk   (%s)\n", l.name, nld,
                 same ? "the SAME " : "the other", (double)r[0] / 400, (double)r[1] * 8192 / (double)r[0], cudaGetErrorString(e));
        }
  }
  {
    long long* d3;
    cudaMalloc(&d3, 64);
    probe_mbar_kernel<<<1, 64>>>(d3);
    cudaError_t e = cudaDeviceSynchronize();
    long long r[8];
    cudaMemcpy(r, d3, 64, cudaMemcpyDeviceToHost);
    printf("mbarrier, completed phase: test_wait %.1f clk, try_wait %.1f clk, mbar_wait (try_wait, then try_wait + suspend hint) %.1f clk   (%s)\n",
           r[0] / 100.0, r[1] / 100.0, r[2] / 100.0, cudaGetErrorString(e));
    printf("mbarrier, wake-up after another warp's arrive: test_wait spin %lld clk, try_wait loop %lld clk, mbar_wait %lld clk\n", r[3], r[4], r[5]);
  }
  cudaFuncSetAttribute(probe_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int N : {32, 64}) {
    for (int ni : {1, 2, 4}) {
      probe_multi_kernel<<<4, 128, 200 * 1024>>>(N, 480, ni, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("K-major N=%d, %d issuing thread(s) (different warps), 480 MMAs in total      %8.1f clk / MMA   (%s)\n", N, ni,
             (double)cyc / 480, cudaGetErrorString(e));
    }
  }
  return 0;
}
