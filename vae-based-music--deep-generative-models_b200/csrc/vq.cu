// vq.cu — VectorQuantizer kernels (exact-fp32 search path) behind vqb_vq_* (include/vqb.h).
//
//   vq_prep_kernel     Et[K,D] = E^T, ee[k] = sum_d E[d,k]^2                 (VectorQuantizer.py:180)
//   vq_search_kernel   dist = (xx + ee) - 2 x.E, first-min argmin            (VectorQuantizer.py:173-185)
//   vq_finish_kernel   gather E[:,idx], straight-through output, commitment partial sums, per-code
//                      count / sum statistics                                 (:86-99,114,123-124)
//   vq_ema_kernel      EMA + dead-code restart, separately rounded ops        (:128-145)
//   vq_metrics_kernel  batch usage, running usage, entropy                    (:151-159)
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"

namespace vqb {

// implemented in vq_tc.cu (tcgen05 search); returns VQB_ERR_UNIMPLEMENTED for unsupported shapes
int vq_search_tc(const vqb_vq_desc* d, const void* x, int x_bf16, const float* E, const float* Et, const float* ee,
                 int64_t* idx, void* ws, size_t ws_bytes, cudaStream_t st);
size_t vq_search_tc_workspace_bytes(const vqb_vq_desc* d);
bool vq_search_tc_supported(const vqb_vq_desc* d);

// ---- activation I/O type: fp32 (the reference's) or bf16 (BASELINE configs[3] "bf16": x, q_st, q are bfloat16, everything the
// layer computes with them is fp32 arithmetic on float(x)) ----------------------------------------------------------------
typedef __nv_bfloat16 bf16_t;
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ float ld1(const bf16_t* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16_t* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(bf16_t* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16_t* p, const float4& v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

__global__ void vq_prep_kernel(const float* __restrict__ E, int D, int K, float* __restrict__ Et,
                               float* __restrict__ ee) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float s = 0.f;
  for (int d = 0; d < D; ++d) {
    const float v = E[(size_t)d * K + k];
    Et[(size_t)k * D + d] = v;
    s = fmaf(v, v, s);
  }
  ee[k] = s;
}

constexpr int VQ_RT = 32;   // rows per CTA
constexpr int VQ_CT = 128;  // codes per tile
constexpr int VQ_DC = 64;   // depth chunk

// 256 threads: tx = tid % 32 -> 4 codes, ty = tid / 32 -> 4 rows.
template <typename XT>
__global__ void __launch_bounds__(256) vq_search_kernel(const XT* __restrict__ x, const float* __restrict__ E,
                                                        const float* __restrict__ ee, long N, int D, int K,
                                                        int64_t* __restrict__ idx) {
  __shared__ __align__(16) float Xs[VQ_DC][VQ_RT + 4];   // transposed rows (+4: fewer store conflicts)
  __shared__ __align__(16) float Es[VQ_DC][VQ_CT];
  __shared__ float xx_s[VQ_RT];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const long n0 = (long)blockIdx.x * VQ_RT;
  const int ndc = (D + VQ_DC - 1) / VQ_DC;

  // ||x||^2 per row (sequential in d, one thread per row)
  if (tid < VQ_RT) {
    float s = 0.f;
    const long n = n0 + tid;
    if (n < N)
      for (int d = 0; d < D; ++d) { const float v = ld1(x + n * D + d); s = fmaf(v, v, s); }
    xx_s[tid] = s;
  }
  float best[4];
  int bidx[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { best[r] = INFINITY; bidx[r] = 0; }

  for (int k0 = 0; k0 < K; k0 += VQ_CT) {
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    for (int dc = 0; dc < ndc; ++dc) {
      const int d0 = dc * VQ_DC;
      __syncthreads();
      if (ndc > 1 || k0 == 0) {
        for (int e = tid; e < VQ_RT * VQ_DC; e += 256) {
          const int r = e / VQ_DC, d = e - r * VQ_DC;  // coalesced over d
          const long n = n0 + r;
          Xs[d][r] = (n < N && d0 + d < D) ? ld1(x + n * D + d0 + d) : 0.f;
        }
      }
      for (int e = tid; e < VQ_DC * VQ_CT; e += 256) {
        const int d = e / VQ_CT, c = e - d * VQ_CT;
        Es[d][c] = (d0 + d < D && k0 + c < K) ? E[(size_t)(d0 + d) * K + k0 + c] : 0.f;
      }
      __syncthreads();
      const int dmax = min(VQ_DC, D - d0);
#pragma unroll 4
      for (int d = 0; d < dmax; ++d) {
        const float4 xv = *(const float4*)&Xs[d][ty * 4];
        const float4 ev = *(const float4*)&Es[d][tx * 4];
        const float xr[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          acc[r][0] = fmaf(xr[r], ev.x, acc[r][0]);
          acc[r][1] = fmaf(xr[r], ev.y, acc[r][1]);
          acc[r][2] = fmaf(xr[r], ev.z, acc[r][2]);
          acc[r][3] = fmaf(xr[r], ev.w, acc[r][3]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int k = k0 + tx * 4 + c;
      if (k < K) {
        const float e2 = ee[k];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float dist = __fsub_rn(__fadd_rn(xx_s[ty * 4 + r], e2), 2.f * acc[r][c]);
          if (dist < best[r]) { best[r] = dist; bidx[r] = k; }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float bd = best[r];
    int bi = bidx[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    const long n = n0 + ty * 4 + r;
    if (tx == 0 && n < N) idx[n] = bi;
  }
}

// one warp per row, 8 warps x VQ_FR rows per CTA.  Per-code sums are accumulated in a [K, D] (code-major) scratch so that
// a row contributes D CONTIGUOUS floats: 16-byte vector reductions (red.global.add.v4.f32) when D % 4 == 0, i.e. 4x fewer
// L2 atomic operations than the [D, K] layout the reference keeps; vq_stats_transpose_kernel produces m_batch [D, K].
constexpr int VQ_FR = 4;
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <typename XT>
__global__ void __launch_bounds__(256) vq_finish_kernel(const XT* __restrict__ x, const float* __restrict__ Et,
                                                        const int64_t* __restrict__ idx, long N, int D, int K,
                                                        XT* __restrict__ q_st, XT* __restrict__ q,
                                                        float* __restrict__ m_kd, float* __restrict__ n_batch,
                                                        float* __restrict__ loss_partial) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float ls = 0.f;
  const bool vec = (D & 3) == 0;
  for (int i = 0; i < VQ_FR; ++i) {
    const long n = ((long)blockIdx.x * 8 + wid) * VQ_FR + i;
    if (n >= N) break;
    const int k = (int)idx[n];
    if (vec) {
      for (int d = lane * 4; d < D; d += 128) {
        const float4 xv = ld4(x + n * D + d);
        const float4 qv = *reinterpret_cast<const float4*>(Et + (size_t)k * D + d);
        const float4 df = make_float4(__fsub_rn(qv.x, xv.x), __fsub_rn(qv.y, xv.y), __fsub_rn(qv.z, xv.z), __fsub_rn(qv.w, xv.w));
        if (q) st4(q + n * D + d, qv);
        if (q_st) st4(q_st + n * D + d, make_float4(__fadd_rn(xv.x, df.x), __fadd_rn(xv.y, df.y), __fadd_rn(xv.z, df.z), __fadd_rn(xv.w, df.w)));
        ls = fmaf(df.x, df.x, ls); ls = fmaf(df.y, df.y, ls); ls = fmaf(df.z, df.z, ls); ls = fmaf(df.w, df.w, ls);
        if (m_kd) red_add_v4(m_kd + (size_t)k * D + d, xv.x, xv.y, xv.z, xv.w);
      }
    } else {
      for (int d = lane; d < D; d += 32) {
        const float xv = ld1(x + n * D + d);
        const float qv = Et[(size_t)k * D + d];
        const float diff = __fsub_rn(qv, xv);
        if (q) st1(q + n * D + d, qv);
        if (q_st) st1(q_st + n * D + d, __fadd_rn(xv, diff));
        ls = fmaf(diff, diff, ls);
        if (m_kd) atomicAdd(&m_kd[(size_t)k * D + d], xv);
      }
    }
    if (n_batch && lane == 0) atomicAdd(&n_batch[k], 1.0f);
  }
  const float s = block_sum(ls, red);
  if (threadIdx.x == 0) loss_partial[blockIdx.x] = s;
}

// Shared-memory privatised variant (K*D + K floats fit one CTA's shared memory, e.g. 512 x 64): persistent CTAs of 1024
// threads accumulate the per-code sums of their slab of rows with SHARED-memory reductions and write one [K, D] partial
// each; the partials are reduced in a fixed order.  No global atomics at all.
constexpr int VQ_FS_THREADS = 1024;
template <typename XT>
__global__ void __launch_bounds__(VQ_FS_THREADS, 1) vq_finish_smem_kernel(
    const XT* __restrict__ x, const float* __restrict__ Et, const int64_t* __restrict__ idx, long N, int D, int K,
    XT* __restrict__ q_st, XT* __restrict__ q, int stats, float* __restrict__ partial_kd,
    float* __restrict__ partial_n, float* __restrict__ loss_partial) {
  extern __shared__ __align__(16) float fsm[];  // [K*D] sums, [K] counts
  __shared__ float red[32];
  float* sm_n = fsm + (size_t)K * D;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (stats) {
    for (int e = tid; e < K * D + K; e += VQ_FS_THREADS) fsm[e] = 0.f;
    __syncthreads();
  }
  const long per = (N + gridDim.x - 1) / gridDim.x;
  const long r0 = (long)blockIdx.x * per, r1 = min(N, r0 + per);
  const int hl = lane & 15, hw = lane >> 4;  // half-warp per row: 16 lanes x float4 = 64 channels per pass
  float ls = 0.f;
  // VQ_FU rows per half-warp and iteration, every load issued before the first use: the kernel is a gather with a dependent
  // index load per row, so its bandwidth is set by how many rows are in flight per SM (Little: ~35 KB at 6.5 TB/s)
  constexpr int VQ_FU = 4;
  const bool perm = (D & 63) == 0;  // bank-permuted per-code rows (see the reductions below)
  constexpr int RS = (VQ_FS_THREADS / 32) * 2;  // rows per CTA pass
  for (long n0 = r0 + wid * 2 + hw; n0 < r1; n0 += RS * VQ_FU) {
    int k[VQ_FU];
#pragma unroll
    for (int u = 0; u < VQ_FU; ++u) k[u] = n0 + (long)u * RS < r1 ? (int)idx[n0 + (long)u * RS] : -1;
    for (int d = hl * 4; d < D; d += 64) {
      float4 xv[VQ_FU], qv[VQ_FU];
#pragma unroll
      for (int u = 0; u < VQ_FU; ++u)
        if (k[u] >= 0) xv[u] = ld4(x + (n0 + (long)u * RS) * D + d);
#pragma unroll
      for (int u = 0; u < VQ_FU; ++u)
        if (k[u] >= 0) qv[u] = *reinterpret_cast<const float4*>(Et + (size_t)k[u] * D + d);
#pragma unroll
      for (int u = 0; u < VQ_FU; ++u) {
        if (k[u] < 0) continue;
        const long n = n0 + (long)u * RS;
        const float4 df = make_float4(__fsub_rn(qv[u].x, xv[u].x), __fsub_rn(qv[u].y, xv[u].y), __fsub_rn(qv[u].z, xv[u].z),
                                      __fsub_rn(qv[u].w, xv[u].w));
        if (q) st4(q + n * D + d, qv[u]);
        if (q_st) st4(q_st + n * D + d, make_float4(__fadd_rn(xv[u].x, df.x), __fadd_rn(xv[u].y, df.y), __fadd_rn(xv[u].z, df.z), __fadd_rn(xv[u].w, df.w)));
        ls = fmaf(df.x, df.x, ls); ls = fmaf(df.y, df.y, ls); ls = fmaf(df.z, df.z, ls); ls = fmaf(df.w, df.w, ls);
        if (stats) {
          // channel d + c lives at word (d & ~63) + 16 * c + (d % 64) / 4 of the code's row: the 16 lanes of a half-warp hit
          // 16 consecutive banks per instruction, and the two half-warps take c in opposite pairs (0,1,2,3 / 1,0,3,2), i.e.
          // the other 16 banks: no conflicts (a [d, d+3] float4 per lane straight into the row would be 4-way conflicted)
          if (perm) {
            float* a = fsm + (size_t)k[u] * D + (d & ~63) + hl;
            const float e0 = hw ? xv[u].y : xv[u].x, e1 = hw ? xv[u].x : xv[u].y;
            const float e2 = hw ? xv[u].w : xv[u].z, e3 = hw ? xv[u].z : xv[u].w;
            atomicAdd(a + 16 * hw, e0); atomicAdd(a + 16 * (1 - hw), e1);
            atomicAdd(a + 32 + 16 * hw, e2); atomicAdd(a + 32 + 16 * (1 - hw), e3);
          } else {  // D not a multiple of 64: rows in natural order
            float* a = fsm + (size_t)k[u] * D + d;
            atomicAdd(a, xv[u].x); atomicAdd(a + 1, xv[u].y); atomicAdd(a + 2, xv[u].z); atomicAdd(a + 3, xv[u].w);
          }
        }
      }
    }
    if (stats && hl == 0) {
#pragma unroll
      for (int u = 0; u < VQ_FU; ++u)
        if (k[u] >= 0) atomicAdd(&sm_n[k[u]], 1.0f);
    }
  }
  const float s = block_sum(ls, red);
  if (tid == 0) loss_partial[blockIdx.x] = s;
  if (stats) {
    __syncthreads();
    float* pk = partial_kd + (size_t)blockIdx.x * K * D;
    for (int e = tid; e < K * D; e += VQ_FS_THREADS) {  // undo the bank permutation of the rows
      const int dd = e % D, c = dd & 3, q4 = (dd & 63) >> 2;
      pk[e] = perm ? fsm[e - dd + (dd & ~63) + 16 * c + q4] : fsm[e];
    }
    for (int e = tid; e < K; e += VQ_FS_THREADS) partial_n[(size_t)blockIdx.x * K + e] = sm_n[e];
  }
}

// m_batch[d, k] = m_kd[k, d]
__global__ void vq_stats_transpose_kernel(const float* __restrict__ m_kd, int D, int K, float* __restrict__ m_batch) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= D * K) return;
  const int d = e / K, k = e - d * K;
  m_batch[e] = m_kd[(size_t)k * D + d];
}

// out[0] = scale * sum(partial[0..n))  — single block, fixed order
__global__ void __launch_bounds__(256) final_sum_kernel(const float* __restrict__ partial, long n, float scale,
                                                        float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (long i = threadIdx.x; i < n; i += 256) s += partial[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s * scale;
}

__global__ void vq_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ x, const float* __restrict__ q,
                              long n, float c, float* __restrict__ dx) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dx[i] = fmaf(c, x[i] - q[i], dq ? dq[i] : 0.f);
}

// block (32, 8): x -> code, y strides the embedding dimension
__global__ void __launch_bounds__(256) vq_ema_kernel(int D, int K, float g, float omg, float thr,
                                                     const float* __restrict__ m_batch,
                                                     const float* __restrict__ n_batch,
                                                     const float* __restrict__ rows, float* __restrict__ E,
                                                     float* __restrict__ m_t, float* __restrict__ N_t) {
  const int k = blockIdx.x * 32 + threadIdx.x;
  float Nn = 0.f;
  if (k < K) {
    Nn = __fadd_rn(__fmul_rn(g, N_t[k]), __fmul_rn(omg, n_batch[k]));  // VectorQuantizer.py:131
    const float usage = Nn >= thr ? 1.f : 0.f;                         // :133 (uses the updated N_t)
    const float den = fminf(fmaxf(Nn, 1e-8f), 1e8f);                   // :144 clip_by_value
    for (int d = threadIdx.y; d < D; d += 8) {
      const size_t o = (size_t)d * K + k;
      const float mn = __fadd_rn(__fmul_rn(g, m_t[o]), __fmul_rn(omg, m_batch[o]));  // :128
      m_t[o] = mn;
      const float reset = __fmul_rn(__fsub_rn(1.f, usage), rows[(size_t)k * D + d]);  // :138
      E[o] = __fadd_rn(__fmul_rn(usage, __fdiv_rn(mn, den)), reset);                  // :144-145
    }
  }
  __syncthreads();
  if (k < K && threadIdx.y == 0) N_t[k] = Nn;
}

__global__ void __launch_bounds__(256) vq_metrics_kernel(int K, float thr, const float* __restrict__ n_batch,
                                                         const float* __restrict__ N_t, float* __restrict__ metrics) {
  __shared__ float red[32];
  __shared__ float tot_s;
  float tot = 0.f, bu = 0.f, u = 0.f;
  for (int k = threadIdx.x; k < K; k += 256) {
    tot += n_batch[k];
    bu += n_batch[k] >= thr ? 1.f : 0.f;
    u += N_t[k] >= thr ? 1.f : 0.f;
  }
  tot = block_sum(tot, red);
  if (threadIdx.x == 0) tot_s = tot;
  bu = block_sum(bu, red);
  u = block_sum(u, red);
  __syncthreads();
  const float total = tot_s;
  float ent = 0.f;
  for (int k = threadIdx.x; k < K; k += 256) {
    const float p = __fdiv_rn(n_batch[k], total);
    ent += p * logf(p + 1e-8f);
  }
  ent = block_sum(ent, red);
  if (threadIdx.x == 0) { metrics[0] = bu; metrics[1] = u; metrics[2] = -ent; }
}

__global__ void gather_rows_kernel(const float* __restrict__ x, long N, int D, const int64_t* __restrict__ ids, int n_ids,
                                   long Ntot, long off, float* __restrict__ rows) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)n_ids * D) return;
  const int i = (int)(e / D), d = (int)(e - (long)i * D);
  long r = ids[i] % Ntot;
  if (r < 0) r += Ntot;
  r -= off;
  rows[e] = (r >= 0 && r < N) ? x[r * D + d] : 0.f;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// K distinct pseudo-random rows of the (tiled) batch = the first K entries of a keyed pseudo-random PERMUTATION of [0, Nt):
// a 4-round Feistel network over the smallest even-width bit domain covering Nt, cycle-walked into range (a permutation
// restricted by cycle-walking stays a permutation).  Every id is an independent-looking draw, so the restart rows are a
// uniform sample of the batch like tf.random.shuffle(...)[:K] (VectorQuantizer.py:137) — not an arithmetic progression, whose
// rows are equally spaced and, for a small stride, neighbouring frames of one clip.  Same ids on every rank (seed, step).
__device__ __forceinline__ uint64_t feistel_perm(uint64_t v, int hb, uint64_t key) {
  const uint64_t mask = (1ull << hb) - 1ull;
  uint64_t L = v >> hb, R = v & mask;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const uint64_t t = L ^ (splitmix64(R ^ key ^ ((uint64_t)(r + 1) * 0x9E3779B97F4A7C15ull)) & mask);
    L = R;
    R = t;
  }
  return (L << hb) | R;
}

__global__ void restart_ids_kernel(long N, int K, uint64_t seed, const int64_t* __restrict__ step, int64_t* __restrict__ ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  const uint64_t Nt = N >= K ? (uint64_t)N : (uint64_t)N * (uint64_t)((K + N - 1) / N);
  const uint64_t key = splitmix64(seed ^ splitmix64((uint64_t)(step ? step[0] : 0)));
  int hb = 1;
  while ((1ull << (2 * hb)) < Nt) ++hb;  // domain 2^(2 hb) >= Nt, < 4 Nt: fewer than 4 walks expected
  uint64_t v = (uint64_t)i;
  do { v = feistel_perm(v, hb, key); } while (v >= Nt);
  ids[i] = (int64_t)v;
}

__global__ void gather_codes_kernel(const float* __restrict__ E, int D, int K, const int64_t* __restrict__ idx, long n,
                                    float* __restrict__ out) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * D) return;
  const long i = e / D;
  const int d = (int)(e - i * D);
  long k = idx[i];
  k = k < 0 ? 0 : (k >= K ? K - 1 : k);
  out[e] = E[(size_t)d * K + k];
}

__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ x, const float* __restrict__ r, long n, float gscale,
                                                  const float* __restrict__ dr_add, float* __restrict__ dr,
                                                  float* __restrict__ partial) {
  __shared__ float red[32];
  float s = 0.f;
  const long base = (long)blockIdx.x * 256 * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long e = base + i * 256 + threadIdx.x;
    if (e < n) {
      const float df = r[e] - x[e];
      s = fmaf(df, df, s);
      if (dr) dr[e] = fmaf(gscale, df, dr_add ? dr_add[e] : 0.f);
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long n, float lr, const float* __restrict__ lr_dev,
                                                   float b1, float b2, float eps,
                                                   float gs, const int64_t* __restrict__ step) {
  __shared__ float lr_s;
  if (threadIdx.x == 0) {
    const double t = (double)(step[0] + 1);
    if (lr_dev) lr = lr_dev[0];  // learning rate kept in device memory: survives CUDA-graph capture, follows schedules
    lr_s = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
  }
  __syncthreads();
  const float lr_t = lr_s;
  const float omb1 = 1.f - b1, omb2 = 1.f - b2;
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    const float gg = g[i] * gs;
    const float mm = m[i] + (gg - m[i]) * omb1;
    const float vv = v[i] + (gg * gg - v[i]) * omb2;
    m[i] = mm;
    v[i] = vv;
    p[i] = p[i] - lr_t * mm / (sqrtf(vv) + eps);
  }
}

__global__ void increment_kernel(int64_t* c) { c[0] += 1; }

struct LincombParams {
  const float* ptr[96];
  float coef[96];
  int start[33];
  int n_out;
};
__global__ void lincomb_kernel(const LincombParams p, float* __restrict__ out) {
  const int i = threadIdx.x;
  if (i >= p.n_out) return;
  float s = 0.f;
  for (int j = p.start[i]; j < p.start[i + 1]; ++j) s = __fadd_rn(s, __fmul_rn(p.coef[j], p.ptr[j][0]));
  out[i] = s;
}

}  // namespace vqb

using namespace vqb;

extern "C" {

size_t vqb_reduce_workspace_bytes(int64_t n) { return (size_t)cdiv(n, 2048) * sizeof(float) + 16; }

constexpr int VQ_FS_GRID = 148;
static bool vq_finish_smem_ok(const vqb_vq_desc* d) {
  return (d->D & 3) == 0 && ((size_t)d->K * d->D + d->K) * sizeof(float) <= 200 * 1024 && d->N >= 4096;
}
static size_t vq_base_ws_floats(const vqb_vq_desc* d) {
  const size_t nfin = (size_t)cdiv(d->N, 8 * VQ_FR) + VQ_FS_GRID;
  size_t n = 2 * (size_t)d->K * d->D + d->K + nfin + 64;  // Et, code-major statistics scratch, ee, loss partials
  if (vq_finish_smem_ok(d)) n += (size_t)VQ_FS_GRID * ((size_t)d->K * d->D + d->K);  // per-CTA partial sums / counts
  return n;
}

size_t vqb_vq_fwd_workspace_bytes(const vqb_vq_desc* d) {
  if (!d || d->N < 0 || d->D <= 0 || d->K <= 0) return 0;
  size_t b = vq_base_ws_floats(d) * sizeof(float);
  b = (b + 1023) & ~(size_t)1023;
  if (d->precision != VQB_PREC_FP32 && vq_search_tc_supported(d)) b += vq_search_tc_workspace_bytes(d);
  return b;
}

}  // extern "C"

template <typename XT>
static int vq_fwd_impl(const vqb_vq_desc* d, const XT* x, const float* E, int64_t* idx, XT* q_st, XT* q,
                       float* loss, float* m_batch, float* n_batch, void* workspace, size_t workspace_bytes, void* stream) {
  constexpr bool BF = !std::is_same<XT, float>::value;
  VQB_ARCH();
  VQB_REQUIRE(d && d->N >= 0 && d->D > 0 && d->K > 0, "vqb_vq_fwd: bad descriptor");
  VQB_REQUIRE(E && (d->N == 0 || (x && idx)), "vqb_vq_fwd: NULL pointer");
  VQB_REQUIRE((m_batch == nullptr) == (n_batch == nullptr), "vqb_vq_fwd: m_batch and n_batch go together");
  const size_t need = vqb_vq_fwd_workspace_bytes(d);
  if (!workspace || workspace_bytes < need)
    return set_err(VQB_ERR_WORKSPACE, "vqb_vq_fwd workspace: need %zu bytes, got %zu", need, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  float* Et = (float*)workspace;
  float* m_kd = Et + (size_t)d->K * d->D;
  float* ee = m_kd + (size_t)d->K * d->D;
  float* part = ee + d->K;
  const long N = d->N;
  const int D = d->D, K = d->K;
  vq_prep_kernel<<<cdiv(K, 128), 128, 0, st>>>(E, D, K, Et, ee);
  VQB_LAUNCH_CHECK();
  if (m_batch) {
    VQB_CUDA(cudaMemsetAsync(N == 0 ? m_batch : m_kd, 0, (size_t)D * K * sizeof(float), st));
    VQB_CUDA(cudaMemsetAsync(n_batch, 0, (size_t)K * sizeof(float), st));
  }
  if (N == 0) {
    if (loss) VQB_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    return VQB_OK;
  }
  if (d->precision == VQB_PREC_FP32 || !vq_search_tc_supported(d)) {  // shapes without a tensor-core kernel: exact fp32 search
    vq_search_kernel<XT><<<cdiv(N, VQ_RT), 256, 0, st>>>(x, E, ee, N, D, K, idx);
    VQB_LAUNCH_CHECK();
  } else {
    size_t off = (vq_base_ws_floats(d) * sizeof(float) + 1023) & ~(size_t)1023;
    int rc = vq_search_tc(d, x, BF ? 1 : 0, E, Et, ee, idx, (char*)workspace + off, workspace_bytes - off, st);
    if (rc != VQB_OK) return rc;
  }
  int nfin = cdiv(N, 8 * VQ_FR);
  if (vq_finish_smem_ok(d)) {
    nfin = VQ_FS_GRID;
    float* pkd = part + cdiv(N, 8 * VQ_FR) + VQ_FS_GRID;
    float* pn = pkd + (size_t)VQ_FS_GRID * K * D;
    const size_t smem = ((size_t)K * D + K) * sizeof(float);
    static size_t smem_set = 0;
    if (smem > smem_set) {
      VQB_CUDA(cudaFuncSetAttribute(vq_finish_smem_kernel<XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set = smem;
    }
    vq_finish_smem_kernel<XT><<<VQ_FS_GRID, VQ_FS_THREADS, smem, st>>>(x, Et, idx, N, D, K, q_st, q, m_batch ? 1 : 0, pkd, pn, part);
    VQB_LAUNCH_CHECK();
    if (m_batch) {
      reduce_chunks_strided(pkd, VQ_FS_GRID, (long)K * D, 0, K * D, m_kd, st);
      VQB_LAUNCH_CHECK();
      reduce_chunks_strided(pn, VQ_FS_GRID, K, 0, K, n_batch, st);
      VQB_LAUNCH_CHECK();
    }
  } else {
    vq_finish_kernel<XT><<<nfin, 256, 0, st>>>(x, Et, idx, N, D, K, q_st, q, m_batch ? m_kd : nullptr, n_batch, part);
    VQB_LAUNCH_CHECK();
  }
  if (m_batch) {
    vq_stats_transpose_kernel<<<cdiv((long)D * K, 256), 256, 0, st>>>(m_kd, D, K, m_batch);
    VQB_LAUNCH_CHECK();
  }
  if (loss) {
    final_sum_kernel<<<1, 256, 0, st>>>(part, nfin, d->beta / ((float)N * (float)D), loss);
    VQB_LAUNCH_CHECK();
  }
  return VQB_OK;
}

extern "C" {

int vqb_vq_fwd(const vqb_vq_desc* d, const float* x, const float* E, int64_t* idx, float* q_st, float* q,
               float* loss, float* m_batch, float* n_batch, void* workspace, size_t workspace_bytes, void* stream) {
  return vq_fwd_impl<float>(d, x, E, idx, q_st, q, loss, m_batch, n_batch, workspace, workspace_bytes, stream);
}

int vqb_vq_fwd_bf16(const vqb_vq_desc* d, const uint16_t* x, const float* E, int64_t* idx, uint16_t* q_st, uint16_t* q,
                    float* loss, float* m_batch, float* n_batch, void* workspace, size_t workspace_bytes, void* stream) {
  return vq_fwd_impl<bf16_t>(d, reinterpret_cast<const bf16_t*>(x), E, idx, reinterpret_cast<bf16_t*>(q_st),
                             reinterpret_cast<bf16_t*>(q), loss, m_batch, n_batch, workspace, workspace_bytes, stream);
}

int vqb_vq_bwd(const vqb_vq_desc* d, const float* dq_out, const float* x, const float* q, float loss_scale,
               float* dx, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(d && x && q && dx, "vqb_vq_bwd: NULL pointer");
  const long n = d->N * d->D;
  if (n == 0) return VQB_OK;
  const float c = loss_scale * 2.f * d->beta / ((float)d->N * (float)d->D);
  vq_bwd_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(dq_out, x, q, n, c, dx);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

int vqb_vq_ema_update(int32_t D, int32_t K, double gamma, float threshold, const float* m_batch,
                      const float* n_batch, const float* restart_rows, float* E, float* m_t, float* N_t,
                      float* metrics, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(D > 0 && K > 0, "vqb_vq_ema_update: bad shape");
  VQB_REQUIRE(m_batch && n_batch && restart_rows && E && m_t && N_t, "vqb_vq_ema_update: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const float g = (float)gamma, omg = (float)(1.0 - gamma);  // python-float arithmetic, then cast (VectorQuantizer.py:128)
  vq_ema_kernel<<<cdiv(K, 32), dim3(32, 8), 0, st>>>(D, K, g, omg, threshold, m_batch, n_batch, restart_rows, E, m_t, N_t);
  VQB_LAUNCH_CHECK();
  if (metrics) {
    vq_metrics_kernel<<<1, 256, 0, st>>>(K, threshold, n_batch, N_t, metrics);
    VQB_LAUNCH_CHECK();
  }
  return VQB_OK;
}

int vqb_gather_rows(const float* x, int64_t N, int32_t D, const int64_t* ids, int32_t n_ids, int64_t N_total,
                    int64_t row_offset, float* rows, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(x && ids && rows && N > 0 && D > 0 && n_ids >= 0 && N_total > 0, "vqb_gather_rows: bad argument");
  const long n = (long)n_ids * D;
  if (n == 0) return VQB_OK;
  gather_rows_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, N, D, ids, n_ids, N_total, row_offset, rows);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

int vqb_restart_ids(int64_t N, int32_t K, uint64_t seed, const int64_t* step_counter, int64_t* ids, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(N > 0 && K > 0 && ids, "vqb_restart_ids: bad argument");
  restart_ids_kernel<<<cdiv(K, 128), 128, 0, (cudaStream_t)stream>>>(N, K, seed, step_counter, ids);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

int vqb_gather_codes(const float* E, int32_t D, int32_t K, const int64_t* idx, int64_t n, float* out, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(E && idx && out && D > 0 && K > 0 && n >= 0, "vqb_gather_codes: bad argument");
  if (n == 0) return VQB_OK;
  gather_codes_kernel<<<cdiv(n * D, 256), 256, 0, (cudaStream_t)stream>>>(E, D, K, idx, n, out);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

int vqb_mse(const float* x, const float* r, int64_t n, float loss_scale, const float* dr_add, float* loss,
            float* dr, void* workspace, size_t workspace_bytes, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(x && r && loss && n > 0, "vqb_mse: bad argument");
  const size_t need = vqb_reduce_workspace_bytes(n);
  if (!workspace || workspace_bytes < need)
    return set_err(VQB_ERR_WORKSPACE, "vqb_mse workspace: need %zu bytes, got %zu", need, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = cdiv(n, 2048);
  mse_kernel<<<nb, 256, 0, st>>>(x, r, n, loss_scale * 2.f / (float)n, dr_add, dr, (float*)workspace);
  VQB_LAUNCH_CHECK();
  final_sum_kernel<<<1, 256, 0, st>>>((const float*)workspace, nb, 1.f / (float)n, loss);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

int vqb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2,
                  float eps, float grad_scale, const int64_t* step_counter, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(p && g && m && v && step_counter && n >= 0, "vqb_adam_step: bad argument");
  if (n == 0) return VQB_OK;
  int nb = cdiv(n, 256);
  if (nb > 148 * 16) nb = 148 * 16;
  adam_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, nullptr, b1, b2, eps, grad_scale, step_counter);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

int vqb_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev, float b1, float b2,
                      float eps, float grad_scale, const int64_t* step_counter, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(p && g && m && v && lr_dev && step_counter && n >= 0, "vqb_adam_step_dev: bad argument");
  if (n == 0) return VQB_OK;
  int nb = cdiv(n, 256);
  if (nb > 148 * 16) nb = 148 * 16;
  adam_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, 0.f, lr_dev, b1, b2, eps, grad_scale, step_counter);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

int vqb_lincomb(int32_t n_out, const int32_t* term_start, const float* const* term_ptr, const float* term_coef, float* out,
                void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(n_out >= 0 && n_out <= 32 && (n_out == 0 || (term_start && term_ptr && term_coef && out)), "vqb_lincomb: bad argument");
  if (n_out == 0) return VQB_OK;
  VQB_REQUIRE(term_start[0] == 0 && term_start[n_out] <= 96, "vqb_lincomb: at most 96 terms (got %d)", term_start[n_out]);
  LincombParams p{};
  p.n_out = n_out;
  for (int i = 0; i <= n_out; ++i) {
    VQB_REQUIRE(i == 0 || term_start[i] >= term_start[i - 1], "vqb_lincomb: term_start must be non-decreasing");
    p.start[i] = term_start[i];
  }
  for (int j = 0; j < term_start[n_out]; ++j) {
    VQB_REQUIRE(term_ptr[j] != nullptr, "vqb_lincomb: NULL term %d", j);
    p.ptr[j] = term_ptr[j]; p.coef[j] = term_coef[j];
  }
  lincomb_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p, out);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

int vqb_increment(int64_t* counter, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(counter, "vqb_increment: NULL pointer");
  increment_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

}  // extern "C"
