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

template <bool TF32>
struct WgCfg {
  static constexpr int ES = TF32 ? 4 : 2;
  static constexpr int T = 16 / ES;               // channels per 16-byte chunk
  static constexpr int NP = 32 / T;               // data planes per operand
  static constexpr int KMMA = 32 / ES;            // K per MMA (16 bf16 / 8 tf32)
  static constexpr int TK = TF32 ? 128 : 256;     // time rows per tile
  static constexpr int DMAX = 32;
  static constexpr int NPB = 48 / T;              // B planes incl. the ones plane and zero padding up to N = 48
  static constexpr int PLANE_A = TK * 16 + 32;
  static constexpr int PLANE_B = (TK + 2 * DMAX) * 16 + 32;
  static constexpr int BUF = NP * PLANE_A + NPB * PLANE_B;
  static constexpr int MCHUNKS = 128 / T;         // 16-byte chunks an M = 128 operand spans
  static constexpr int SPAN = MCHUNKS * PLANE_A;  // bytes the A descriptor may touch from its start
  static constexpr int SMEM = (2 * BUF > BUF + SPAN ? 2 * BUF : BUF + SPAN) + 128;
  static constexpr int PART = 3 * 32 * 32 + 32;
};

template <bool TF32>
__global__ void __launch_bounds__(256, 2) wgrad_tc_kernel(const WgTcParams p) {
  using Cfg = WgCfg<TF32>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[3];
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) tmem_alloc(&tslot, 128);
  if (tid == 32) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); fence_mbar_init(); }
  // constant planes of both buffers: channel 32 = 1 (bias column), channels 33..47 = 0
  for (int buf = 0; buf < 2; ++buf) {
    uint8_t* Bt = smem + buf * Cfg::BUF + Cfg::NP * Cfg::PLANE_A;
    for (int e = tid; e < (Cfg::NPB - Cfg::NP) * (Cfg::PLANE_B / 16); e += 256) {
      const int pl = e / (Cfg::PLANE_B / 16), r = e - pl * (Cfg::PLANE_B / 16);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (pl == 0) v.x = TF32 ? __float_as_uint(1.0f) : 0x00003F80u;  // element 0 of the chunk = 1.0
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
  const uint32_t idesc32 = instr_desc(TF32 ? FMT_TF32 : FMT_BF16, 128, 32, true, true);
  const uint32_t idesc48 = instr_desc(TF32 ? FMT_TF32 : FMT_BF16, 128, 48, true, true);

  int it = 0;
  for (long tile = first; tile < last; ++tile, ++it) {
    const int buf = it & 1;
    const int b = (int)(tile / p.tiles_per_b);
    const int t0 = (int)(tile - (long)b * p.tiles_per_b) * Cfg::TK;
    uint8_t* At = smem + buf * Cfg::BUF;
    uint8_t* Bt = At + Cfg::NP * Cfg::PLANE_A;
    if (it >= 2) mbar_wait(&bars[buf], ((it >> 1) - 1) & 1);  // the MMAs that read this buffer have completed
    const float* otb = p.ot + (size_t)b * p.L * 32;
    const float* gab = p.ga + (size_t)b * p.L * 32;
    for (int e = tid; e < Cfg::TK * 8; e += 256) {
      const int r = e >> 3, q = e & 7;
      const int g = t0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g < p.L) v = *reinterpret_cast<const float4*>(otb + (size_t)g * 32 + q * 4);
      if (TF32) *reinterpret_cast<float4*>(At + q * Cfg::PLANE_A + r * 16) = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
      else *reinterpret_cast<uint2*>(At + (q >> 1) * Cfg::PLANE_A + r * 16 + (q & 1) * 8) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
    for (int e = tid; e < rowsB * 8; e += 256) {
      const int r = e >> 3, q = e & 7;
      const int g = t0 - p.dil + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g >= 0 && g < p.L) v = *reinterpret_cast<const float4*>(gab + (size_t)g * 32 + q * 4);
      if (p.relu_ga) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      if (TF32) *reinterpret_cast<float4*>(Bt + q * Cfg::PLANE_B + r * 16) = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
      else *reinterpret_cast<uint2*>(Bt + (q >> 1) * Cfg::PLANE_B + r * 16 + (q & 1) * 8) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t a0 = smem_u32(At), b0 = smem_u32(Bt);
#pragma unroll 1
      for (int ks = 0; ks < Cfg::TK / Cfg::KMMA; ++ks) {
        const uint64_t ad = smem_desc(a0 + ks * Cfg::KMMA * 16, 128, Cfg::PLANE_A);
        const uint32_t acc = (it | ks) != 0;
        // tap j reads act(x) rows t + (j-1)*dil = B-tile rows ks*KMMA + j*dil
        mma<TF32>(tmem + 0, ad, smem_desc(b0 + (ks * Cfg::KMMA) * 16, 128, Cfg::PLANE_B), idesc32, acc);
        mma<TF32>(tmem + 32, ad, smem_desc(b0 + (ks * Cfg::KMMA + p.dil) * 16, 128, Cfg::PLANE_B), idesc48, acc);
        mma<TF32>(tmem + 80, ad, smem_desc(b0 + (ks * Cfg::KMMA + 2 * p.dil) * 16, 128, Cfg::PLANE_B), idesc32, acc);
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
    if (warp == 0) {  // accumulator rows 0..31 = output channel co = lane
      float v[32];
      const int co = tid;
      tmem_ld32(tmem + 0, v);
#pragma unroll
      for (int n = 0; n < 32; ++n) out[(0 * 32 + n) * 32 + co] = v[n];
      tmem_ld32(tmem + 32, v);
#pragma unroll
      for (int n = 0; n < 32; ++n) out[(1 * 32 + n) * 32 + co] = v[n];
      tmem_ld32(tmem + 80, v);
#pragma unroll
      for (int n = 0; n < 32; ++n) out[(2 * 32 + n) * 32 + co] = v[n];
      tmem_ld32(tmem + 64, v);  // column 64 = centre tap's column 32 = bias gradient
      out[3 * 32 * 32 + co] = v[0];
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

bool wgrad_tc_supported(const vqb_conv_desc* d) {
  return d->k == 3 && d->stride == 1 && d->C_in == 32 && d->C_out == 32 && d->dilation >= 1 && d->dilation <= 32 &&
         d->precision == VQB_PREC_BF16;  // kind::tf32 with MN-major (time-as-K) operands returned zeros on B200: bf16 only
}

static int wgrad_tc_grid(const vqb_conv_desc* d, int* tiles_per_b) {
  const int TK = d->precision == VQB_PREC_TF32 ? WgCfg<true>::TK : WgCfg<false>::TK;
  *tiles_per_b = cdiv(d->L, TK);
  const long total = (long)d->B * *tiles_per_b;
  return (int)(total < 296 ? (total > 0 ? total : 1) : 296);
}

size_t wgrad_tc_workspace_bytes(const vqb_conv_desc* d) {
  int tpb;
  return (size_t)wgrad_tc_grid(d, &tpb) * WgCfg<false>::PART * sizeof(float) + 64;
}

template <bool TF32>
static int launch_wg(const WgTcParams& p, int grid, cudaStream_t st) {
  using Cfg = WgCfg<TF32>;
  static bool attr_set = false;
  if (!attr_set) {
    VQB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  wgrad_tc_kernel<TF32><<<grid, 256, Cfg::SMEM, st>>>(p);
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
  int rc = d->precision == VQB_PREC_TF32 ? launch_wg<true>(p, grid, st) : launch_wg<false>(p, grid, st);
  if (rc) return rc;
  constexpr int PART = WgCfg<false>::PART;
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
