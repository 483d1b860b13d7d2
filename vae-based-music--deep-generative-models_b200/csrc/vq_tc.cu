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
      if (tid == 0) {
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

static bool vq_tc_ok(const vqb_vq_desc* d) {
  return d->D == VT_D && d->K >= VT_CHUNK && d->K % VT_CHUNK == 0 && d->K <= 8192 &&
         (d->precision == VQB_PREC_BF16 || d->precision == VQB_PREC_TF32 || d->precision == VQB_PREC_BF16X2 ||
          d->precision == VQB_PREC_BF16X3);
}

bool vq_search_tc_supported(const vqb_vq_desc* d) { return vq_tc_ok(d); }

size_t vq_search_tc_workspace_bytes(const vqb_vq_desc* d) {
  if (!vq_tc_ok(d)) return 0;
  return (size_t)(d->K / VT_CHUNK) * VT_CHUNK_BYTES + 256;
}

int vq_search_tc(const vqb_vq_desc* d, const float* x, const float* E, const float* Et, const float* ee, int64_t* idx,
                 void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!vq_tc_ok(d))
    return set_err(VQB_ERR_UNIMPLEMENTED, "tensor-core VQ search needs D = 64 and K a multiple of 256 (got D=%d K=%d)", d->D, d->K);
  const size_t need = vq_search_tc_workspace_bytes(d);
  if (!ws || ws_bytes < need) return set_err(VQB_ERR_WORKSPACE, "VQ tensor-core workspace: need %zu bytes, got %zu", need, ws_bytes);
  VqTcParams p{};
  p.x = x; p.Et = Et; p.ee = ee; p.idx = idx; p.N = d->N; p.K = d->K;
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
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    VQB_CUDA(cudaGetDevice(&dev));
    VQB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long grid = p.ntiles < 2L * num_sms ? p.ntiles : 2L * num_sms;
  vq_tc_kernel<<<(int)grid, 128, smem, st>>>(p);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

}  // namespace vqb
