// wgrad_tc.cu — tensor-core weight gradient of the k=3, stride-1, 32->32 convolutions (the 208 residual-block
// convolutions of SMALL_VQ_VAE):   dW[j][ci][co] = sum_{b,t} act(x)[b, t + (j-1)*dil, ci] * dy[b, t, co],
//                                  dbias[co]     = sum_{b,t} dy[b, t, co].
// The reduction over time is the MMA K dimension: per tile of TK time rows both operands are staged once in the
// "plane layout" of tc.cuh and consumed as MN-major operands (rows = time = K):
//   A (M side)  = dy tile,            M = 128 rows of which the first 32 are the output channels (the MMA costs the
//                                     same for M = 64 and 128; rows 32..127 read whatever follows in shared memory and
//                                     their accumulator rows are never read),
//   B (N side)  = act(x) tile shifted by (j-1)*dil rows for tap j, N = 32 input channels; the centre tap uses N = 48
//                 with an extra constant plane of ones, so that accumulator column 32 is the bias gradient.
// Accumulators (3 taps x fp32 [128 x 32/48]) stay in TMEM for ALL tiles a CTA processes (persistent CTAs, two-stage
// shared-memory pipeline: the MMAs of tile i run while tile i+1 is being staged); one partial result per CTA goes to
// the workspace and is reduced in a fixed order (deterministic).
#include "common.cuh"
#include "tc.cuh"

namespace vqb {

using namespace tc;

struct WgTcParams {
  const float* ga;  // gather side  [B, L, 32]  (x)
  const float* ot;  // other side   [B, L, 32]  (dy)
  float* partial;   // [gridDim.x][3*32*32 + 32]
  int B, L, dil, relu_ga, tiles_per_b, total_tiles;
};

// S = number of bf16 pieces each operand is split into (1: bf16, 2: bf16x2, 3: bf16x3 = fp32-grade products)
template <int S_>
struct WgCfg {
  static constexpr int S = S_;
  static constexpr int T = 8;                     // channels per 16-byte chunk (bf16)
  static constexpr int NP = 4;                    // data planes per operand
  static constexpr int KMMA = 16;                 // K per MMA
  static constexpr int TK = S == 1 ? 256 : 128;   // time rows per tile
  static constexpr int DMAX = 32;
  static constexpr int NPB = 6;                   // B planes incl. the ones plane and zero padding up to N = 48
  static constexpr int PLANE_A = TK * 16 + 32;
  static constexpr int PLANE_B = (TK + 2 * DMAX) * 16 + 32;
  static constexpr int TILE_A = NP * PLANE_A;
  static constexpr int TILE_B = NPB * PLANE_B;
  static constexpr int BUF = S * (TILE_A + TILE_B);
  static constexpr int SPAN = 16 * PLANE_A;       // bytes an M = 128 A descriptor may touch from its start
  static constexpr int NEED = BUF + (S - 1) * TILE_A + SPAN;
  static constexpr int SMEM = (2 * BUF > NEED ? 2 * BUF : NEED) + 128;
  static constexpr int PART = 3 * 32 * 32 + 32;
  static constexpr int MINB = SMEM > 113 * 1024 ? 1 : 2;
};

template <int S>
__device__ __forceinline__ void stage_split(uint8_t* tile, int tile_bytes, int plane_bytes, int r, int q, float4 v) {
  float a[3], b[3], c[3], d[3];
  split_bf16<S>(v.x, a); split_bf16<S>(v.y, b); split_bf16<S>(v.z, c); split_bf16<S>(v.w, d);
#pragma unroll
  for (int s = 0; s < S; ++s)
    *reinterpret_cast<uint2*>(tile + s * tile_bytes + (q >> 1) * plane_bytes + r * 16 + (q & 1) * 8) =
        make_uint2(pack_bf16(a[s], b[s]), pack_bf16(c[s], d[s]));
}

template <int S>
__global__ void __launch_bounds__(256, WgCfg<S>::MINB) wgrad_tc_kernel(const WgTcParams p) {
  using Cfg = WgCfg<S>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[3];
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) tmem_alloc(&tslot, 128);
  if (tid == 32) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); fence_mbar_init(); }
  // constant planes of every B tile: piece 0 has channel 32 = 1 (bias column); everything else in channels 32..47 is 0
  for (int t = 0; t < 2 * S; ++t) {
    const int buf = t / S, sp = t - buf * S;
    uint8_t* Bt = smem + buf * Cfg::BUF + S * Cfg::TILE_A + sp * Cfg::TILE_B;
    for (int e = tid; e < (Cfg::NPB - Cfg::NP) * (Cfg::PLANE_B / 16); e += 256) {
      const int pl = e / (Cfg::PLANE_B / 16), r = e - pl * (Cfg::PLANE_B / 16);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (pl == 0 && sp == 0) v.x = 0x00003F80u;  // element 0 of the chunk = bf16 1.0
      *reinterpret_cast<uint4*>(Bt + (Cfg::NP + pl) * Cfg::PLANE_B + r * 16) = v;
    }
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tslot;

  const long first = (long)blockIdx.x * p.total_tiles / gridDim.x;
  const long last = (long)(blockIdx.x + 1) * p.total_tiles / gridDim.x;
  const int rowsB = Cfg::TK + 2 * p.dil;
  const uint32_t idesc32 = instr_desc(FMT_BF16, 128, 32, true, true);
  const uint32_t idesc48 = instr_desc(FMT_BF16, 128, 48, true, true);

  int it = 0;
  for (long tile = first; tile < last; ++tile, ++it) {
    const int buf = it & 1;
    const int b = (int)(tile / p.tiles_per_b);
    const int t0 = (int)(tile - (long)b * p.tiles_per_b) * Cfg::TK;
    uint8_t* At = smem + buf * Cfg::BUF;
    uint8_t* Bt = At + S * Cfg::TILE_A;
    if (it >= 2) mbar_wait(&bars[buf], ((it >> 1) - 1) & 1);  // the MMAs that read this buffer have completed
    const float* otb = p.ot + (size_t)b * p.L * 32;
    const float* gab = p.ga + (size_t)b * p.L * 32;
    for (int e = tid; e < Cfg::TK * 8; e += 256) {
      const int r = e >> 3, q = e & 7;
      const int g = t0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < p.L) v = *reinterpret_cast<const float4*>(otb + (size_t)g * 32 + q * 4);
      stage_split<S>(At, Cfg::TILE_A, Cfg::PLANE_A, r, q, v);
    }
    for (int e = tid; e < rowsB * 8; e += 256) {
      const int r = e >> 3, q = e & 7;
      const int g = t0 - p.dil + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g >= 0 && g < p.L) v = *reinterpret_cast<const float4*>(gab + (size_t)g * 32 + q * 4);
      if (p.relu_ga) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      stage_split<S>(Bt, Cfg::TILE_B, Cfg::PLANE_B, r, q, v);
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t a0 = smem_u32(At), b0 = smem_u32(Bt);
      uint32_t acc = it != 0;
      // The S pieces of dy are consecutive plane groups of the A tile, i.e. consecutive 32-row blocks of ONE M = 128
      // operand: a single MMA multiplies all of them with one piece of act(x); accumulator row block r then holds
      // dy_r^T act(x) summed over the x pieces, and the epilogue adds the S row blocks.
#pragma unroll 1
      for (int ks = 0; ks < Cfg::TK / Cfg::KMMA; ++ks) {
        const uint64_t ad = smem_desc(a0 + ks * Cfg::KMMA * 16, 128, Cfg::PLANE_A);
#pragma unroll
        for (int sb = S - 1; sb >= 0; --sb) {
          const uint32_t bb = b0 + sb * Cfg::TILE_B + (ks * Cfg::KMMA) * 16;
          // tap j reads act(x) rows t + (j-1)*dil = B-tile rows ks*KMMA + j*dil
          mma<false>(tmem + 0, ad, smem_desc(bb, 128, Cfg::PLANE_B), idesc32, acc);
          mma<false>(tmem + 32, ad, smem_desc(bb + p.dil * 16, 128, Cfg::PLANE_B), idesc48, acc);
          mma<false>(tmem + 80, ad, smem_desc(bb + 2 * p.dil * 16, 128, Cfg::PLANE_B), idesc32, acc);
          acc = 1;
        }
      }
      commit(&bars[buf]);
    }
    __syncwarp();
  }
  if (tid == 0) commit(&bars[2]);  // everything issued so far
  __syncwarp();
  float* out = p.partial + (size_t)blockIdx.x * Cfg::PART;
  if (it == 0) {  // no tiles: zero partial
    for (int e = tid; e < Cfg::PART; e += 256) out[e] = 0.f;
  } else {
    mbar_wait(&bars[2], 0);
    fence_after_sync();
    float* red = reinterpret_cast<float*>(smem);  // [S][PART]: the operand tiles are dead now
    if (warp < S) {  // accumulator rows 32*warp .. +31 = output channel co = lane, for dy piece `warp`
      float v[32];
      const int co = tid & 31;
      float* r = red + warp * Cfg::PART;
      const uint32_t ta = tmem + (((uint32_t)warp * 32u) << 16);
      tmem_ld32(ta + 0, v);
#pragma unroll
      for (int n = 0; n < 32; ++n) r[(0 * 32 + n) * 32 + co] = v[n];
      tmem_ld32(ta + 32, v);
#pragma unroll
      for (int n = 0; n < 32; ++n) r[(1 * 32 + n) * 32 + co] = v[n];
      tmem_ld32(ta + 80, v);
#pragma unroll
      for (int n = 0; n < 32; ++n) r[(2 * 32 + n) * 32 + co] = v[n];
      tmem_ld32(ta + 64, v);  // column 64 = centre tap's column 32 = bias gradient
      r[3 * 32 * 32 + co] = v[0];
    }
    fence_before_sync();
    __syncthreads();
    for (int e = tid; e < Cfg::PART; e += 256) {
      float t = red[e];
#pragma unroll
      for (int w = 1; w < S; ++w) t += red[w * Cfg::PART + e];
      out[e] = t;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

static int wg_split(int precision) { return precision == VQB_PREC_BF16X3 ? 3 : precision == VQB_PREC_BF16X2 ? 2 : 1; }

bool wgrad_tc_supported(const vqb_conv_desc* d) {
  // kind::tf32 with MN-major (time-as-K) operands returned zeros on B200, so VQB_PREC_TF32 has no weight-gradient kernel
  return d->k == 3 && d->stride == 1 && d->C_in == 32 && d->C_out == 32 && d->dilation >= 1 && d->dilation <= 32 &&
         (d->precision == VQB_PREC_BF16 || d->precision == VQB_PREC_BF16X2 || d->precision == VQB_PREC_BF16X3);
}

static int wgrad_tc_grid(const vqb_conv_desc* d, int* tiles_per_b) {
  const int TK = wg_split(d->precision) == 1 ? WgCfg<1>::TK : WgCfg<2>::TK;
  *tiles_per_b = cdiv(d->L, TK);
  const long total = (long)d->B * *tiles_per_b;
  return (int)(total < 296 ? (total > 0 ? total : 1) : 296);
}

size_t wgrad_tc_workspace_bytes(const vqb_conv_desc* d) {
  int tpb;
  return (size_t)wgrad_tc_grid(d, &tpb) * WgCfg<1>::PART * sizeof(float) + 64;
}

template <int S>
static int launch_wg(const WgTcParams& p, int grid, cudaStream_t st) {
  using Cfg = WgCfg<S>;
  static bool attr_set = false;
  if (!attr_set) {
    VQB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  wgrad_tc_kernel<S><<<grid, 256, Cfg::SMEM, st>>>(p);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

int conv1d_wgrad_tc(const vqb_conv_desc* d, const float* x, const float* dy, float* dw, float* dbias, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
  const size_t need = wgrad_tc_workspace_bytes(d);
  if (!ws || ws_bytes < need) return set_err(VQB_ERR_WORKSPACE, "tensor-core wgrad workspace: need %zu bytes, got %zu", need, ws_bytes);
  WgTcParams p{};
  p.ga = x; p.ot = dy; p.partial = (float*)ws;
  p.B = d->B; p.L = d->L; p.dil = d->dilation; p.relu_ga = d->relu_in;
  const int grid = wgrad_tc_grid(d, &p.tiles_per_b);
  p.total_tiles = d->B * p.tiles_per_b;
  const int S = wg_split(d->precision);
  int rc = S == 3 ? launch_wg<3>(p, grid, st) : S == 2 ? launch_wg<2>(p, grid, st) : launch_wg<1>(p, grid, st);
  if (rc) return rc;
  constexpr int PART = WgCfg<1>::PART;
  // dw and dbias are separate buffers: two fixed-order reductions over the per-CTA partials
  reduce_chunks_strided(p.partial, grid, PART, 0, 3 * 32 * 32, dw, st);
  VQB_LAUNCH_CHECK();
  if (dbias) {
    reduce_chunks_strided(p.partial, grid, PART, 3 * 32 * 32, 32, dbias, st);
    VQB_LAUNCH_CHECK();
  }
  return VQB_OK;
}

}  // namespace vqb
