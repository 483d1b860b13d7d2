// vq_tc.cu — tensor-core nearest-code search (VectorQuantizer.get_code_indices, VectorQuantizer.py:170-186) for D = 64.
//
// The [N, K] similarity matrix X.E of the reference is never materialised: per tile of 128 latents the dot products
// against 256 codes at a time are one tcgen05 accumulator (fp32, 128 lanes x 256 columns in TMEM) produced by 4 bf16 MMAs
// (M = 128, N = 256, K = 16 each) from the x tile and the codebook staged in the plane layout of tc.cuh.
// bf16 operands make those dot products approximate, and the parity rule wants the EXACT fp32 argmin, so the scan is
// "approximate, then verify":
//   pass 1  every thread (one latent = one TMEM lane) scans its 256 columns: score = ee[k] - 2 dot, running minimum;
//   pass 2  it re-reads the columns and, for every code whose score is within `margin` of the running minimum — margin
//           = 2 * 2^-7 * sqrt(||x||^2 * max||e||^2), a rigorous bound on the bf16 rounding of both scores — evaluates the
//           distance exactly as the fp32 kernel does ((xx + ee) - 2 x.e with sequential fp32 FMAs) and keeps the exact
//           minimum (first index on ties).
// Any code that could be the exact argmin passes the margin test against a minimum that only decreases, so the result is
// the exact-fp32 argmin; typically 1-3 codes per latent are re-evaluated.
// Persistent CTAs (2 per SM, 256 TMEM columns each): the codebook (K <= 512) stays resident in shared memory, so one
// CTA's MMAs overlap the other's scan.
#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace vqb {

using namespace tc;

constexpr int VT_ROWS = 128;    // latents per tile
constexpr int VT_CHUNK = 256;   // codes per accumulator
constexpr int VT_D = 64;
constexpr int VT_NP = 8;        // planes (8 bf16 per 16 B)
constexpr int VT_PLANE_A = VT_ROWS * 16 + 32;
constexpr int VT_PLANE_B = VT_CHUNK * 16 + 32;
constexpr int VT_CHUNK_BYTES = VT_NP * VT_PLANE_B;   // one packed 256-code chunk
constexpr int VT_MAX_RESIDENT = 2;                   // chunks kept in shared memory (K <= 512)
constexpr int VT_MAXC = 12;                          // candidate slots per latent

struct VqTcParams {
  const float* x;       // [N, 64]
  const float* Et;      // [K, 64] fp32 (exact re-evaluation)
  const float* ee;      // [K]
  const uint8_t* Epk;   // packed bf16 codebook: [K/256][8 planes][256 codes * 16 B (+pad)]
  int64_t* idx;
  long N;
  int K, nchunks, resident;
  float ee_max;         // filled on device: see ee_max_ptr
  const float* ee_max_ptr;
  long ntiles;
};

// codebook E [64, K] fp32 -> bf16 plane layout per 256-code chunk; also max_k ||e_k||^2
__global__ void vq_pack_codebook_kernel(const float* __restrict__ E, int K, const float* __restrict__ ee,
                                        uint8_t* __restrict__ Epk, float* __restrict__ ee_max) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) {
    const int ch = k / VT_CHUNK, r = k - ch * VT_CHUNK;
    uint8_t* base = Epk + (size_t)ch * VT_CHUNK_BYTES + r * 16;
    for (int p = 0; p < VT_NP; ++p) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        w[i] = pack_bf16(E[(size_t)(p * 8 + 2 * i) * K + k], E[(size_t)(p * 8 + 2 * i + 1) * K + k]);
      *reinterpret_cast<uint4*>(base + p * VT_PLANE_B) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  if (blockIdx.x == 0) {  // max ||e||^2 (single block pass, fixed order)
    __shared__ float red[32];
    float m = 0.f;
    for (int i = threadIdx.x; i < K; i += blockDim.x) m = fmaxf(m, ee[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < (int)(blockDim.x + 31) / 32; ++i) t = fmaxf(t, red[i]);
      ee_max[0] = t;
    }
  }
}

__global__ void __launch_bounds__(128, 2) vq_tc_kernel(const VqTcParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* As = smem;                                   // x tile, bf16 planes
  uint8_t* Bs = As + VT_NP * VT_PLANE_A;                // resident / streamed codebook chunks
  const int nbuf = p.resident ? p.nchunks : 1;
  float* ee_s = reinterpret_cast<float*>(Bs + (size_t)nbuf * VT_CHUNK_BYTES);  // [K]
  uint64_t* bar = reinterpret_cast<uint64_t*>(ee_s + p.K);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) tmem_alloc(tslot, 256);
  if (tid == 32) { mbar_init(bar, 1); fence_mbar_init(); }
  for (int i = tid; i < p.K; i += 128) ee_s[i] = p.ee[i];
  if (p.resident) {
    const uint4* src = reinterpret_cast<const uint4*>(p.Epk);
    uint4* dst = reinterpret_cast<uint4*>(Bs);
    for (int i = tid; i < p.nchunks * (VT_CHUNK_BYTES / 16); i += 128) dst[i] = src[i];
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tslot;
  const float ee_max = p.ee_max_ptr[0];
  const uint32_t idesc = instr_desc(FMT_BF16, 128, VT_CHUNK, false, false);
  const uint32_t taddr = tmem + (((uint32_t)warp * 32u) << 16);
  uint32_t phase = 0;

#pragma unroll 1
  for (long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const long n = tile * VT_ROWS + tid;   // this thread's latent
    const bool valid = n < p.N;
    // own row: fp32 registers (exact re-evaluation) and bf16 planes (MMA A operand)
    float xr[VT_D];
    float xx = 0.f;
#pragma unroll
    for (int q = 0; q < VT_D / 4; ++q) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) v = *reinterpret_cast<const float4*>(p.x + n * VT_D + q * 4);
      xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int d = 0; d < VT_D; ++d) xx = fmaf(xr[d], xr[d], xx);   // same order as vq_search_kernel
#pragma unroll
    for (int pl = 0; pl < VT_NP; ++pl)
      *reinterpret_cast<uint4*>(As + pl * VT_PLANE_A + tid * 16) =
          make_uint4(pack_bf16(xr[8 * pl], xr[8 * pl + 1]), pack_bf16(xr[8 * pl + 2], xr[8 * pl + 3]),
                     pack_bf16(xr[8 * pl + 4], xr[8 * pl + 5]), pack_bf16(xr[8 * pl + 6], xr[8 * pl + 7]));
    // rigorous bound on |(approx score_a - approx score_b) - (exact ...)|: each score is off by <= 2^-7 |x||e|
    const float margin = 2.f * 0.0078125f * sqrtf(xx * ee_max) * 1.01f + 1e-30f;
    float run_min = INFINITY;     // running minimum of the approximate scores
    int cand[VT_MAXC];            // codes within the margin of the running minimum (a superset of what the final minimum admits)
    float cscore[VT_MAXC];        // their approximate scores (to drop the ones a later, lower minimum rules out)
    int ncand = 0;
    bool overflow = false;

#pragma unroll 1
    for (int ch = 0; ch < p.nchunks; ++ch) {
      uint8_t* Bc = Bs + (p.resident ? (size_t)ch * VT_CHUNK_BYTES : 0);
      if (!p.resident) {  // stream this chunk (K > 512): the previous chunk's MMAs have completed (we waited on them)
        const uint4* src = reinterpret_cast<const uint4*>(p.Epk + (size_t)ch * VT_CHUNK_BYTES);
        uint4* dst = reinterpret_cast<uint4*>(Bc);
        for (int i = tid; i < VT_CHUNK_BYTES / 16; i += 128) dst[i] = src[i];
      }
      fence_proxy_async();
      fence_before_sync();
      __syncthreads();   // A tile (and streamed chunk) visible; every thread has finished reading the previous accumulator
      if (warp == 0 && elect_one()) {
        fence_after_sync();
        const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bc);
#pragma unroll
        for (int kk = 0; kk < VT_D / 16; ++kk)
          mma<false>(tmem, smem_desc(a0 + kk * 2 * VT_PLANE_A, VT_PLANE_A, 128), smem_desc(b0 + kk * 2 * VT_PLANE_B, VT_PLANE_B, 128),
                     idesc, kk != 0);
        commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, phase);
      phase ^= 1;
      fence_after_sync();
      const float* eec = ee_s + ch * VT_CHUNK;
      float v[32];
      // pass 1: running minimum of score = ee - 2 dot
#pragma unroll 1
      for (int c0 = 0; c0 < VT_CHUNK; c0 += 32) {
        tmem_ld32(taddr + c0, v);
#pragma unroll
        for (int c = 0; c < 32; ++c) run_min = fminf(run_min, fmaf(-2.f, v[c], eec[c0 + c]));
      }
      // pass 2: remember every code within the margin (ascending code order)
      const float thr = run_min + margin;
#pragma unroll 1
      for (int c0 = 0; c0 < VT_CHUNK; c0 += 32) {
        tmem_ld32(taddr + c0, v);
        uint32_t m = 0;
#pragma unroll
        for (int c = 0; c < 32; ++c) m |= (fmaf(-2.f, v[c], eec[c0 + c]) <= thr ? 1u : 0u) << c;
        while (m) {
          const int c = __ffs(m) - 1;
          m &= m - 1;
          if (ncand == VT_MAXC) {  // full: drop the entries the current (lower) minimum has ruled out, keeping code order
            int w = 0;
#pragma unroll
            for (int i = 0; i < VT_MAXC; ++i)
              if (cscore[i] <= thr) { cand[w] = cand[i]; cscore[w] = cscore[i]; ++w; }
            ncand = w;
          }
          if (ncand < VT_MAXC) {
            cand[ncand] = ch * VT_CHUNK + c0 + c;
            cscore[ncand] = fmaf(-2.f, v[c], eec[c0 + c]);
            ++ncand;
          } else {
            overflow = true;
          }
        }
      }
      fence_before_sync();
    }
    if (!overflow) {  // final pruning with the final minimum
      const float thr = run_min + margin;
      int w = 0;
#pragma unroll
      for (int i = 0; i < VT_MAXC; ++i)
        if (i < ncand && cscore[i] <= thr) { cand[w] = cand[i]; ++w; }
      ncand = w;
    }
    // exact fp32 evaluation (VectorQuantizer.py:175-182 op order, sequential FMAs as in vq_search_kernel).  A single
    // candidate needs none: the exact argmin is always inside the margin set.
    int best_k = cand[0];
    if (valid && (ncand > 1 || overflow)) {
      float best = INFINITY;
      const int cnt = overflow ? p.K : ncand;       // overflow (rare): more candidates than slots -> exact scan of every code
#pragma unroll 1
      for (int ci = 0; ci < cnt; ++ci) {
        const int k = overflow ? ci : cand[ci];
        const float* er = p.Et + (size_t)k * VT_D;
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < VT_D / 4; ++q) {
          const float4 e = *reinterpret_cast<const float4*>(er + q * 4);
          acc = fmaf(xr[4 * q], e.x, acc); acc = fmaf(xr[4 * q + 1], e.y, acc);
          acc = fmaf(xr[4 * q + 2], e.z, acc); acc = fmaf(xr[4 * q + 3], e.w, acc);
        }
        const float dist = __fsub_rn(__fadd_rn(xx, ee_s[k]), 2.f * acc);
        if (dist < best) { best = dist; best_k = k; }   // candidates are in ascending code order: first minimum wins
      }
    }
    if (valid) p.idx[n] = best_k;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// =====================================================================================================================
// Resident-codebook search (K = 256 or 512): bf16x2 scores, single scan, exact verification only where it matters.
//
//   score'[n][k] = x_n . e_k - ||e_k||^2 / 2      (argmax score' = argmin distance)
// comes out of the tensor core directly: x and E are split into two bf16 pieces each (hi + lo, 16 mantissa bits) and the
// three leading piece products hi.hi + hi.lo + lo.hi are accumulated into ONE fp32 accumulator by chaining the MMAs along
// K (no epilogue summation); -||e||^2/2 rides along as one more K step (a constant-ones A plane against a B plane holding
// its three bf16 pieces).  Error of a score: <= 2^-15 |x| |e|.
// Scan: TMEM reads are the scarce resource (64 B/clk), so every accumulator element is read exactly once; two threads per
// latent (128 columns each per 256-code chunk) track the largest score with its index and the second largest.  If the
// gap exceeds the error margin the index IS the exact-fp32 argmin (VectorQuantizer.py:173-185 evaluated with sequential
// fp32 FMAs); otherwise (rare) the row is re-evaluated exactly against all K codes by a whole warp.
// Pipeline (one persistent CTA per SM): 8 scan warps + 1 MMA warp; two 256-column accumulators alternate, so the MMAs of
// the next chunk / next tile run under the scan of the current one; the next tile's rows are prefetched into registers.
constexpr int V2_PLANE_A = VT_ROWS * 16 + 64;              // +64: the 8 planes' rows fall into distinct 16-byte bank groups
constexpr int V2_PLANE_B = VT_CHUNK * 16 + 64;
constexpr int V2_TILE_A = VT_NP * V2_PLANE_A;              // one bf16 piece of the x tile
constexpr int V2_CHUNK_B = (2 * VT_NP + 2) * V2_PLANE_B;   // hi planes, lo planes, 2 planes for -ee/2
constexpr int V2_NSCAN = 512;                              // scan / staging threads (4 per latent: 64 columns each per chunk)
constexpr int V2_PARTS = V2_NSCAN / 128;                   // column groups per accumulator
constexpr int V2_COLS = VT_CHUNK / V2_PARTS;
constexpr int V2_NLD = (VT_ROWS * 8) / V2_NSCAN;           // 8-channel units of the x tile per thread
constexpr float V2_MARGIN = 2.5f * 3.0517578125e-5f;       // 2 scores x 2^-15, x1.25 for the exact evaluation's own rounding

struct Vq2Params {
  const void* x;        // [N, 64] fp32, or bf16 in the BF instantiations (vqb_vq_fwd_bf16)
  const float* Et;      // [K, 64] fp32 (exact re-evaluation)
  const float* ee;      // [K]
  const uint8_t* Epk;   // packed codebook, V2_CHUNK_B bytes per 256-code chunk
  const float* ee_max_ptr;
  int64_t* idx;
  int* fcount;          // number of rows whose top-2 gap is inside the error margin
  int* flist;           // their row numbers (any order)
  long N, ntiles;
  int K, nchunks;
  // Codebooks beyond the resident 512 codes are searched in passes of <= 512 codes (same kernel, the x tiles are re-read): pass
  // `code0 / 512` scans codes [code0, code0 + 256 * nchunks), merges its (best, second, index) with the state the previous
  // passes left per row (st_*) and, unless it is the last pass, writes the state back; the last pass decides.
  int code0, first_pass, last_pass, Ktot;
  float* st_m1;
  float* st_m2;
  int* st_i1;
  float* fbest;         // exact pass over code blocks (vq2_exact_kernel): running exact minimum per flagged row
  int* fk;
  long long* trace;     // TRACE build (tools/trace_vq.py): clock64 stamps of CTA 0's tiles 3 and 4, [2][16]
};

// TRACE: clock64 stamp of event `ev` (CTA 0, tiles 3 and 4, one lane)
#define V2_TR(ev)                                                                                                   \
  do {                                                                                                              \
    if (TRACE && p.trace && blockIdx.x == 0 && (it == 3 || it == 4) && lane == 0) p.trace[(it - 3) * 24 + (ev)] = clock64(); \
  } while (0)

__global__ void vq2_pack_codebook_kernel(const float* __restrict__ E, int K, const float* __restrict__ ee,
                                         uint8_t* __restrict__ Epk, float* __restrict__ ee_max) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) {
    const int ch = k / VT_CHUNK, r = k - ch * VT_CHUNK;
    uint8_t* base = Epk + (size_t)ch * V2_CHUNK_B + r * 16;
    for (int pl = 0; pl < VT_NP; ++pl) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = E[(size_t)(pl * 8 + i) * K + k];
      uint4 pc[2];
      split8<2>(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), pc);
      *reinterpret_cast<uint4*>(base + pl * V2_PLANE_B) = pc[0];
      *reinterpret_cast<uint4*>(base + (VT_NP + pl) * V2_PLANE_B) = pc[1];
    }
    float h[3];
    split_bf16<3>(-0.5f * ee[k], h);
    *reinterpret_cast<uint4*>(base + (2 * VT_NP) * V2_PLANE_B) = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], 0.f), 0u, 0u);
    *reinterpret_cast<uint4*>(base + (2 * VT_NP + 1) * V2_PLANE_B) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (blockIdx.x == 0) {  // max ||e||^2 (single block pass, fixed order)
    __shared__ float red[32];
    float m = 0.f;
    for (int i = threadIdx.x; i < K; i += blockDim.x) m = fmaxf(m, ee[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < (int)(blockDim.x + 31) / 32; ++i) t = fmaxf(t, red[i]);
      ee_max[0] = t;
    }
  }
}

__device__ __forceinline__ void bar_scan() { asm volatile("bar.sync 1, %0;" ::"n"(V2_NSCAN) : "memory"); }

// Exact fp32 argmin for the rows the search could not decide (the arithmetic of vq_search_kernel: sequential FMAs over
// d, (xx + ee) - 2 x.e, first minimum).  The fp32 codebook Et [K, 64] is staged once per CTA in shared memory (row stride
// 65 words: conflict-free for one thread per code); the CTAs stride over the list four rows at a time (one codebook read
// serves four independent FMA chains), one thread per code.
constexpr int V2_EX_THREADS = 512;
constexpr int V2_EX_ROWS = 4;
template <bool BF>
__global__ void __launch_bounds__(V2_EX_THREADS, 1) vq2_exact_kernel(const Vq2Params p) {
  extern __shared__ __align__(16) float esm[];  // [KB][65] codebook block, [64][4] x rows (row-interleaved), [4][16] warp results
  const int KB = p.K;                           // codes per block staged in shared memory (<= 512); p.Ktot codes in all
  float* xs = esm + (((size_t)KB * 65 + 3) & ~(size_t)3);
  float* wbest = xs + VT_D * V2_EX_ROWS;
  int* wk = reinterpret_cast<int*>(wbest + V2_EX_ROWS * 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nf = p.fcount[0];
  if ((int)blockIdx.x * V2_EX_ROWS >= nf) return;
  const int nblocks = (p.Ktot + KB - 1) / KB;
  for (int cb = 0; cb < nblocks; ++cb) {
    const int c0 = cb * KB, kn = min(KB, p.Ktot - c0);
    __syncthreads();  // the previous block's codebook has been consumed
    for (int e = tid; e < kn * VT_D; e += V2_EX_THREADS) esm[(e >> 6) * 65 + (e & 63)] = p.Et[(size_t)c0 * VT_D + e];
    for (int f0 = blockIdx.x * V2_EX_ROWS; f0 < nf; f0 += gridDim.x * V2_EX_ROWS) {
      __syncthreads();  // codebook staged / previous rows' buffers consumed
      if (tid < VT_D * V2_EX_ROWS) {
        const int rr = tid >> 6, d = tid & 63;
        xs[d * V2_EX_ROWS + rr] = f0 + rr >= nf ? 0.f
                                  : BF ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.x)[(long)p.flist[f0 + rr] * VT_D + d])
                                       : reinterpret_cast<const float*>(p.x)[(long)p.flist[f0 + rr] * VT_D + d];
      }
      __syncthreads();
      float xx[V2_EX_ROWS] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
      for (int d = 0; d < VT_D; ++d) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + d * V2_EX_ROWS);
        xx[0] = fmaf(xv.x, xv.x, xx[0]); xx[1] = fmaf(xv.y, xv.y, xx[1]);
        xx[2] = fmaf(xv.z, xv.z, xx[2]); xx[3] = fmaf(xv.w, xv.w, xx[3]);
      }
      float best[V2_EX_ROWS] = {INFINITY, INFINITY, INFINITY, INFINITY};
      int bk[V2_EX_ROWS] = {0, 0, 0, 0};
      for (int k = tid; k < kn; k += V2_EX_THREADS) {
        const float* er = esm + (size_t)k * 65;
        float acc[V2_EX_ROWS] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
        for (int d = 0; d < VT_D; ++d) {
          const float4 xv = *reinterpret_cast<const float4*>(xs + d * V2_EX_ROWS);
          const float e = er[d];
          acc[0] = fmaf(xv.x, e, acc[0]); acc[1] = fmaf(xv.y, e, acc[1]);
          acc[2] = fmaf(xv.z, e, acc[2]); acc[3] = fmaf(xv.w, e, acc[3]);
        }
        const float e2 = p.ee[c0 + k];
#pragma unroll
        for (int rr = 0; rr < V2_EX_ROWS; ++rr) {
          const float dist = __fsub_rn(__fadd_rn(xx[rr], e2), 2.f * acc[rr]);
          if (dist < best[rr]) { best[rr] = dist; bk[rr] = c0 + k; }
        }
      }
#pragma unroll
      for (int rr = 0; rr < V2_EX_ROWS; ++rr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float od = __shfl_xor_sync(0xffffffffu, best[rr], o);
          const int ok = __shfl_xor_sync(0xffffffffu, bk[rr], o);
          if (od < best[rr] || (od == best[rr] && ok < bk[rr])) { best[rr] = od; bk[rr] = ok; }
        }
        if (lane == 0) { wbest[rr * 16 + warp] = best[rr]; wk[rr * 16 + warp] = bk[rr]; }
      }
      __syncthreads();
      if (tid < V2_EX_ROWS && f0 + tid < nf) {
        float b = wbest[tid * 16];
        int k = wk[tid * 16];
        for (int w = 1; w < V2_EX_THREADS / 32; ++w)
          if (wbest[tid * 16 + w] < b || (wbest[tid * 16 + w] == b && wk[tid * 16 + w] < k)) { b = wbest[tid * 16 + w]; k = wk[tid * 16 + w]; }
        if (nblocks > 1) {  // running exact minimum over the code blocks (earlier blocks = lower indices win ties)
          if (cb > 0 && !(b < p.fbest[f0 + tid])) { b = p.fbest[f0 + tid]; k = p.fk[f0 + tid]; }
          p.fbest[f0 + tid] = b; p.fk[f0 + tid] = k;
        }
        if (cb == nblocks - 1) p.idx[p.flist[f0 + tid]] = k;
      }
    }
  }
}

template <bool TRACE, bool BF>
__global__ void __launch_bounds__(V2_NSCAN + 32, 1) vq2_kernel(const Vq2Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* As = smem;                                 // [2 pieces][8 planes]
  uint8_t* Ac = As + 2 * V2_TILE_A;                   // constant planes: (1, 1, 1, 0, ...) and zeros
  uint8_t* Bs = Ac + 2 * V2_PLANE_A;                  // [nchunks][V2_CHUNK_B]
  uint8_t* misc = Bs + (size_t)p.nchunks * V2_CHUNK_B;
  uint64_t* bar = reinterpret_cast<uint64_t*>(misc);  // [0] A staged; [1],[2] accumulator 0/1 complete; [3],[4] accumulator 0/1 drained
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 5);
  float* xn_s = reinterpret_cast<float*>(misc + 64);  // [2][128] |x| per row (tile parity)
  float* pm1 = xn_s + 2 * VT_ROWS;                    // [2][PARTS-1][128] partial results of column groups 1.. (tile parity)
  float* pm2 = pm1 + 2 * (V2_PARTS - 1) * VT_ROWS;
  int* pi1 = reinterpret_cast<int*>(pm2 + 2 * (V2_PARTS - 1) * VT_ROWS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = blockIdx.x < p.ntiles ? (int)((p.ntiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;

  if (warp == 0) tmem_alloc(tslot, 512);
  if (tid == 32) {
    mbar_init(&bar[0], V2_NSCAN / 32); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1);
    mbar_init(&bar[3], V2_NSCAN / 32); mbar_init(&bar[4], V2_NSCAN / 32);
    fence_mbar_init();
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.Epk);
    uint4* dst = reinterpret_cast<uint4*>(Bs);
    for (int i = tid; i < p.nchunks * (V2_CHUNK_B / 16); i += V2_NSCAN + 32) dst[i] = src[i];
    for (int i = tid; i < 2 * (V2_PLANE_A / 16); i += V2_NSCAN + 32)
      reinterpret_cast<uint4*>(Ac)[i] = i < V2_PLANE_A / 16 ? make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tslot;

  if (warp == V2_NSCAN / 32) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc = instr_desc(FMT_BF16, 128, VT_CHUNK, false, false);
      const uint64_t a_hi = smem_desc(smem_u32(As), V2_PLANE_A, 128), a_lo = a_hi + (uint64_t)(V2_TILE_A >> 4);
      const uint64_t a_c = smem_desc(smem_u32(Ac), V2_PLANE_A, 128);
      for (int it = 0; it < ntiles; ++it) {
        mbar_wait(&bar[0], it & 1);
        fence_after_sync();
        V2_TR(0);
        for (int ch = 0; ch < p.nchunks; ++ch) {
          // accumulator `ch` was drained by the scan of the previous tile (phase it-1 of bar[3+ch])
          if (it > 0) { mbar_wait(&bar[3 + ch], (it - 1) & 1); fence_after_sync(); }
          V2_TR(1 + 2 * ch);
          const uint64_t b_hi = smem_desc(smem_u32(Bs + (size_t)ch * V2_CHUNK_B), V2_PLANE_B, 128);
          const uint64_t b_lo = b_hi + (uint64_t)((VT_NP * V2_PLANE_B) >> 4);
          const uint64_t b_ee = b_hi + (uint64_t)((2 * VT_NP * V2_PLANE_B) >> 4);
          const uint32_t d = tmem + ch * VT_CHUNK;
          mma<false>(d, a_c, b_ee, idesc, 0);  // -||e||^2 / 2 initialises the accumulator
#pragma unroll
          for (int kk = 0; kk < VT_D / 16; ++kk) {
            const uint64_t ka = (uint64_t)((kk * 2 * V2_PLANE_A) >> 4), kb = (uint64_t)((kk * 2 * V2_PLANE_B) >> 4);
            mma<false>(d, a_lo + ka, b_hi + kb, idesc, 1);  // small terms first
            mma<false>(d, a_hi + ka, b_lo + kb, idesc, 1);
            mma<false>(d, a_hi + ka, b_hi + kb, idesc, 1);
          }
          commit(&bar[1 + ch]);
          V2_TR(2 + 2 * ch);
          if (TRACE && p.trace && blockIdx.x == 0 && it == 4) {  // tile 4 only: spin (test_wait, no suspend) until the chunk has completed
            uint32_t ok = 0;
            while (!ok)
              asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
                           : "=r"(ok) : "r"(smem_u32(&bar[1 + ch])), "r"((uint32_t)(it & 1)) : "memory");
            V2_TR(19 + ch);
          }
        }
        if (TRACE && p.trace && blockIdx.x == 0 && (it == 3 || it == 4)) {  // when do the chunks really complete?  (perturbs the next tile slightly)
          mbar_wait(&bar[1], it & 1); V2_TR(12);
          mbar_wait(&bar[2], it & 1); V2_TR(15);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------------ staging + scan warps
    const int oct = tid & 7;                          // 8-channel unit of this thread in the staging loop
    const int qd = warp & 3, part = warp >> 2;        // TMEM lane quadrant, column group
    const int r = qd * 32 + lane;                     // tile row of this thread in the scan
    const float en_max = sqrtf(p.ee_max_ptr[0]);
    float4 ra[V2_NLD], rb[V2_NLD];
    auto load = [&](int it) {
      const long n0 = ((long)blockIdx.x + (long)it * gridDim.x) * VT_ROWS;
#pragma unroll
      for (int k = 0; k < V2_NLD; ++k) {
        const long n = n0 + ((tid + k * V2_NSCAN) >> 3);
        const bool ok = n < p.N;
        if (BF) {  // 8 bf16 channels = one 16-byte load; bf16 -> fp32 is a shift
          const uint4 u = ok ? *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.x) + n * VT_D + oct * 8) : make_uint4(0u, 0u, 0u, 0u);
          ra[k] = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
          rb[k] = make_float4(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xffff0000u), __uint_as_float(u.w << 16), __uint_as_float(u.w & 0xffff0000u));
        } else {
          const float* xf = reinterpret_cast<const float*>(p.x);
          ra[k] = ok ? *reinterpret_cast<const float4*>(xf + n * VT_D + oct * 8) : make_float4(0.f, 0.f, 0.f, 0.f);
          rb[k] = ok ? *reinterpret_cast<const float4*>(xf + n * VT_D + oct * 8 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    auto stage = [&](int it) {
#pragma unroll
      for (int k = 0; k < V2_NLD; ++k) { reg_fence(ra[k]); reg_fence(rb[k]); }
#pragma unroll
      for (int k = 0; k < V2_NLD; ++k) {
        const int row = (tid + k * V2_NSCAN) >> 3;
        uint4 pc[2];
        split8<2>(ra[k], rb[k], pc);
        *reinterpret_cast<uint4*>(As + oct * V2_PLANE_A + row * 16) = pc[0];
        *reinterpret_cast<uint4*>(As + V2_TILE_A + oct * V2_PLANE_A + row * 16) = pc[1];
        float s = ra[k].x * ra[k].x;
        s = fmaf(ra[k].y, ra[k].y, s); s = fmaf(ra[k].z, ra[k].z, s); s = fmaf(ra[k].w, ra[k].w, s);
        s = fmaf(rb[k].x, rb[k].x, s); s = fmaf(rb[k].y, rb[k].y, s); s = fmaf(rb[k].z, rb[k].z, s); s = fmaf(rb[k].w, rb[k].w, s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (oct == 0) xn_s[(it & 1) * VT_ROWS + row] = sqrtf(s);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[0]);
    };
    if (ntiles > 0) { load(0); stage(0); }
#pragma unroll 1
    for (int it = 0; it < ntiles; ++it) {
      const long n0 = ((long)blockIdx.x + (long)it * gridDim.x) * VT_ROWS;
      const bool has_next = it + 1 < ntiles;
      if (has_next) load(it + 1);
      if (warp == 0) { V2_TR(5); V2_TR(16); }
      // two independent (best, second, index) trackers over the even / odd columns halve the dependency chains
      float m1 = -INFINITY, m2 = -INFINITY, n1 = -INFINITY, n2 = -INFINITY;
      int i1 = 0, j1 = 0;
#pragma unroll 1
      for (int ch = 0; ch < p.nchunks; ++ch) {
        mbar_wait(&bar[1 + ch], it & 1);
        if (warp == 0) V2_TR(17 + ch);
        fence_after_sync();
        if (warp == 0) V2_TR(6 + 2 * ch);
        if (ch == p.nchunks - 1 && has_next) stage(it + 1);  // all MMAs of this tile have completed: the x tile is free
        if (warp == 0 && ch == p.nchunks - 1) V2_TR(10);
        const uint32_t taddr = tmem + (((uint32_t)qd * 32u) << 16) + (uint32_t)(ch * VT_CHUNK + part * V2_COLS);
        const int kbase = p.code0 + ch * VT_CHUNK + part * V2_COLS;
#pragma unroll 1
        for (int c0 = 0; c0 < V2_COLS; c0 += 32) {
          float v[32];
          tmem_ld32(taddr + c0, v);
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const float s = v[c], t = v[c + 1];
            const bool gs = s > m1, gt = t > n1;
            m2 = fmaxf(m2, fminf(s, m1)); n2 = fmaxf(n2, fminf(t, n1));
            m1 = fmaxf(m1, s); n1 = fmaxf(n1, t);
            i1 = gs ? kbase + c0 + c : i1; j1 = gt ? kbase + c0 + c + 1 : j1;
          }
        }
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[3 + ch]);
        if (warp == 0) V2_TR(7 + 2 * ch);
        if (warp == 15) V2_TR(13 + ch);
      }
      {  // fold the odd-column tracker into the even one (ties go to the lower index; they are re-evaluated exactly anyway)
        const float second = fmaxf(fminf(m1, n1), fmaxf(m2, n2));
        const bool take = n1 > m1 || (n1 == m1 && j1 < i1);
        m1 = take ? n1 : m1; i1 = take ? j1 : i1; m2 = second;
      }
      // merge the column groups of each row and decide; undecided rows go to the exact kernel's list
      const int par = (it & 1) * (V2_PARTS - 1) * VT_ROWS;
      if (part > 0) {
        const int o = par + (part - 1) * VT_ROWS + r;
        pm1[o] = m1; pm2[o] = m2; pi1[o] = i1;
      }
      bar_scan();
      if (warp == 0) V2_TR(11);
      if (part == 0) {
#pragma unroll
        for (int q = 0; q < V2_PARTS - 1; ++q) {
          const int o = par + q * VT_ROWS + r;
          const float o1 = pm1[o], o2 = pm2[o];
          const int oi = pi1[o];
          const float second = fmaxf(fminf(m1, o1), fmaxf(m2, o2));
          const bool take = o1 > m1;
          m1 = take ? o1 : m1; i1 = take ? oi : i1; m2 = second;
        }
        const long n = n0 + r;
        if (warp == 0) V2_TR(21);
        if (n < p.N) {
          if (!p.first_pass) {  // codes of the earlier passes (lower indices: they win ties)
            const float o1 = p.st_m1[n], o2 = p.st_m2[n];
            const int oi = p.st_i1[n];
            const float second = fmaxf(fminf(m1, o1), fmaxf(m2, o2));
            const bool take = o1 >= m1;
            m1 = take ? o1 : m1; i1 = take ? oi : i1; m2 = second;
          }
          if (!p.last_pass) {
            p.st_m1[n] = m1; p.st_m2[n] = m2; p.st_i1[n] = i1;
          } else {
            // score error <= 2^-15 |x||e| (operand pieces) + accumulation rounding of the -ee/2 term, for both scores
            const float margin = V2_MARGIN * xn_s[(it & 1) * VT_ROWS + r] * en_max + 1e-6f * en_max * en_max;
            p.idx[n] = i1;
            if (!(m1 - m2 > margin)) p.flist[atomicAdd(p.fcount, 1)] = (int)n;
          }
        }
        if (warp == 0) V2_TR(22);
      }
      if (warp == 0) V2_TR(23);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static size_t vq2_smem_bytes(int nchunks) {
  return (size_t)2 * V2_TILE_A + 2 * V2_PLANE_A + (size_t)nchunks * V2_CHUNK_B + 64 + (2 + 6 * (V2_PARTS - 1)) * VT_ROWS * 4 + 64;
}

static bool vq_tc_ok(const vqb_vq_desc* d) {
  return d->D == VT_D && d->K >= VT_CHUNK && d->K % VT_CHUNK == 0 && d->K <= 8192 &&
         (d->precision == VQB_PREC_BF16 || d->precision == VQB_PREC_TF32 || d->precision == VQB_PREC_BF16X2 ||
          d->precision == VQB_PREC_BF16X3 || d->precision == VQB_PREC_FP16X2);
}

bool vq_search_tc_supported(const vqb_vq_desc* d) { return vq_tc_ok(d); }

static bool vq2_ok(const vqb_vq_desc* d) { return d->K / VT_CHUNK <= 16; }  // <= 2 chunks resident per pass; up to 8 passes (K <= 4096)

size_t vq_search_tc_workspace_bytes(const vqb_vq_desc* d) {
  if (!vq_tc_ok(d)) return 0;
  if (vq2_ok(d))  // packed codebook + undecided-row list (+ exact running minima) + per-row search state of a multi-pass search
    return (size_t)(d->K / VT_CHUNK) * V2_CHUNK_B + 256 + 3 * ((size_t)d->N + 16) * sizeof(int) +
           (d->K / VT_CHUNK > 2 ? 3 * ((size_t)d->N + 16) * sizeof(float) : 0);
  return (size_t)(d->K / VT_CHUNK) * VT_CHUNK_BYTES + 256;
}

int vq_search_tc(const vqb_vq_desc* d, const void* x, int x_bf16, const float* E, const float* Et, const float* ee, int64_t* idx,
                 void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!vq_tc_ok(d))
    return set_err(VQB_ERR_UNIMPLEMENTED, "tensor-core VQ search needs D = 64 and K a multiple of 256 (got D=%d K=%d)", d->D, d->K);
  const size_t need = vq_search_tc_workspace_bytes(d);
  if (!ws || ws_bytes < need) return set_err(VQB_ERR_WORKSPACE, "VQ tensor-core workspace: need %zu bytes, got %zu", need, ws_bytes);
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    VQB_CUDA(cudaGetDevice(&dev));
    VQB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (vq2_ok(d)) {
    Vq2Params q{};
    q.x = x; q.Et = Et; q.ee = ee; q.idx = idx; q.N = d->N; q.Ktot = d->K;
    const int nchunks_tot = d->K / VT_CHUNK;
    uint8_t* Epk = (uint8_t*)ws;
    float* ee_max = (float*)(Epk + (size_t)nchunks_tot * V2_CHUNK_B);
    q.ee_max_ptr = ee_max;
    q.fcount = reinterpret_cast<int*>(ee_max + 16);
    q.flist = q.fcount + 16;
    q.fk = q.flist + d->N + 16;
    q.fbest = reinterpret_cast<float*>(q.fk + d->N + 16);
    q.st_m1 = q.fbest + d->N + 16; q.st_m2 = q.st_m1 + d->N + 16; q.st_i1 = reinterpret_cast<int*>(q.st_m2 + d->N + 16);
    VQB_REQUIRE(d->N < (1l << 31), "VQ tensor-core search: N must be below 2^31");
    VQB_CUDA(cudaMemsetAsync(q.fcount, 0, sizeof(int), st));
    q.ntiles = (d->N + VT_ROWS - 1) / VT_ROWS;
    vq2_pack_codebook_kernel<<<cdiv(d->K, 128), 128, 0, st>>>(E, d->K, ee, Epk, ee_max);
    VQB_LAUNCH_CHECK();
    const long grid = q.ntiles < num_sms ? q.ntiles : num_sms;
    for (int c0 = 0; c0 < nchunks_tot; c0 += 2) {  // passes of <= 2 resident chunks (512 codes)
      q.nchunks = nchunks_tot - c0 < 2 ? nchunks_tot - c0 : 2;
      q.K = q.nchunks * VT_CHUNK;
      q.code0 = c0 * VT_CHUNK;
      q.Epk = Epk + (size_t)c0 * V2_CHUNK_B;
      q.first_pass = c0 == 0; q.last_pass = c0 + 2 >= nchunks_tot;
      const size_t smem = vq2_smem_bytes(q.nchunks);
      static size_t smem_set2 = 0;
      if (smem > smem_set2) {
        VQB_CUDA((cudaFuncSetAttribute(vq2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        VQB_CUDA((cudaFuncSetAttribute(vq2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        VQB_CUDA((cudaFuncSetAttribute(vq2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
        smem_set2 = smem;
      }
      if (getenv("VQB_VQ_TRACE") && !x_bf16) {  // profiling aid: address of a device buffer of 48 int64 (tools/trace_vq.py)
        q.trace = reinterpret_cast<long long*>(strtoull(getenv("VQB_VQ_TRACE"), nullptr, 0));
        vq2_kernel<true, false><<<(int)grid, V2_NSCAN + 32, smem, st>>>(q);
      } else if (x_bf16)
        vq2_kernel<false, true><<<(int)grid, V2_NSCAN + 32, smem, st>>>(q);
      else
        vq2_kernel<false, false><<<(int)grid, V2_NSCAN + 32, smem, st>>>(q);
      VQB_LAUNCH_CHECK();
    }
    q.K = d->K < 512 ? d->K : 512;  // codes per block of the exact pass
    const size_t esmem = ((size_t)q.K * 65 + 4 + VT_D * V2_EX_ROWS + 2 * V2_EX_ROWS * 16) * sizeof(float);
    static size_t esmem_set = 0;
    if (esmem > esmem_set) {
      VQB_CUDA(cudaFuncSetAttribute(vq2_exact_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));
      VQB_CUDA(cudaFuncSetAttribute(vq2_exact_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));
      esmem_set = esmem;
    }
    if (x_bf16) vq2_exact_kernel<true><<<num_sms, V2_EX_THREADS, esmem, st>>>(q);
    else vq2_exact_kernel<false><<<num_sms, V2_EX_THREADS, esmem, st>>>(q);
    VQB_LAUNCH_CHECK();
    return VQB_OK;
  }
  if (x_bf16) return set_err(VQB_ERR_UNIMPLEMENTED, "vqb_vq_fwd_bf16: the tensor-core search for K = %d > 4096 takes fp32 activations only", d->K);
  VqTcParams p{};
  p.x = reinterpret_cast<const float*>(x); p.Et = Et; p.ee = ee; p.idx = idx; p.N = d->N; p.K = d->K;
  p.nchunks = d->K / VT_CHUNK;
  p.resident = p.nchunks <= VT_MAX_RESIDENT;
  uint8_t* Epk = (uint8_t*)ws;
  float* ee_max = (float*)(Epk + (size_t)p.nchunks * VT_CHUNK_BYTES);
  p.Epk = Epk; p.ee_max_ptr = ee_max;
  p.ntiles = (d->N + VT_ROWS - 1) / VT_ROWS;
  vq_pack_codebook_kernel<<<cdiv(d->K, 128), 128, 0, st>>>(E, d->K, ee, Epk, ee_max);
  VQB_LAUNCH_CHECK();
  const size_t smem = (size_t)VT_NP * VT_PLANE_A + (size_t)(p.resident ? p.nchunks : 1) * VT_CHUNK_BYTES + (size_t)d->K * 4 + 64;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    VQB_CUDA(cudaFuncSetAttribute(vq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const long grid = p.ntiles < 2L * num_sms ? p.ntiles : 2L * num_sms;
  vq_tc_kernel<<<(int)grid, 128, smem, st>>>(p);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

}  // namespace vqb
