// mma_probe.cu — issue-throughput probe for tcgen05.mma (kind::f16, M = 128, K = 16) with both operands in shared memory
// in the no-swizzle "plane layout" of csrc/tc.cuh: how many SM clocks does one MMA cost as a function of N, of the operand
// majors (K-major as in resblock_tc.cu / conv_tc.cu, MN-major as in wgrad_tc.cu) and of the plane pitch?  The numbers
// feed the kernel cost models in DESIGN.md.  Build + run: tools/run_mma_probe.sh (nvcc here, gpurun for the run).
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
