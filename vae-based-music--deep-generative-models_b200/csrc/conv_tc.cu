// conv_tc.cu — the stride-2 convolutions of the encoder / decoder stages as tcgen05 implicit GEMMs (32 -> 32 channels):
//   DOWN kernel   Conv1D(32, k=4, strides=2) forward          (encdec.py:33)        y[t]  = sum_j x[2t + j - 1] W_j + b
//                 Conv1DTranspose(k=4, s=2) data gradient      (encdec.py:67-68)     dx[m] = sum_k dy[2m + k - 1] Wt_k^T
//   UP kernel     Conv1DTranspose(32, k=4, strides=2) forward  (encdec.py:67-68)     y[2m] = x[m] Wt_1 + x[m-1] Wt_3 + b,
//                                                                                    y[2m+1] = x[m+1] Wt_0 + x[m] Wt_2 + b
//                 Conv1D(k=4, s=2) data gradient                                     same two-phase form over dy with W_j^T
// Each pair shares its tap table and differs only in how the weight tensor is indexed.
//
// The strided access is removed at staging time: the DOWN kernel splits the input rows of a tile by parity into two
// operand tiles (even rows / odd rows) in the plane layout of tc.cuh, after which every tap is a unit-stride view of one of
// them (descriptor start shifted by 0, 1 or 2 rows); the UP kernel keeps one input tile and produces the two output
// phases as two accumulators.  M = 128 rows per MMA, N = 32 output channels x S weight pieces, K = 32 input channels per
// tap; bf16 operands split into S pieces (bf16 / bf16x2 / bf16x3 as in resblock_tc.cu), fp32 accumulation in TMEM.
// One persistent CTA per SM, 512 threads: the global loads of tile i+2 are in flight and the MMAs of tile i+1 run (into the
// other TMEM buffer) while the epilogue of tile i drains its accumulators through the per-warp staged row stores.
#include "common.cuh"
#include "tc.cuh"
#include "tc_rows.cuh"

namespace vqb {

using namespace tc;

struct ConvTcParams {
  const float* in;    // [B, Lin, 32]
  float* out;         // [B, Lout, 32]
  const float* w;     // element (tap j, in-channel k, out-channel n) at w[j*sj + k*si + n*so]
  const float* bias;  // [32] or null
  int sj, si, so;
  int B, Lin, Lout;
  int tiles_x, total_tiles;
};

template <int S_, bool UP_>
struct CtCfg {
  static constexpr int S = S_;
  static constexpr bool UP = UP_;
  static constexpr int NT = 512;
  static constexpr int NP = 4;                       // planes of 8 bf16 channels
  static constexpr int KSTEPS = 2;
  static constexpr int ROWS = UP ? 128 : 256;        // MMA rows per tile: input rows (UP) / output rows (DOWN)
  static constexpr int NSUB = UP ? 1 : 2;            // parity sub-tiles of the input
  static constexpr int AROWS = ROWS + 2;             // one guard row on each side
  static constexpr int PLANE = AROWS * 16;           // 2080 / 4128 bytes: both = 32 (mod 128), the planes' banks interleave
  static constexpr int TILE = NP * PLANE;            // one piece of one sub-tile
  static constexpr int NW = 32 * S;
  static constexpr int WPLANE = NW * 16;
  static constexpr int WTAP = NP * WPLANE;
  static constexpr int WALL = 4 * WTAP;
  static constexpr int ACC = 2 * NW;                 // accumulator columns per TMEM buffer: 2 M blocks (DOWN) / 2 phases (UP)
  static constexpr int TCOLS = 2 * ACC <= 128 ? 128 : 2 * ACC <= 256 ? 256 : 512;
  static constexpr int STG = (NT / 32) * 2048;
  static constexpr int SMEM = NSUB * S * TILE + WALL + STG + 64 + 128;
  static constexpr int NU = (NSUB * AROWS * 4 + NT - 1) / NT;  // 8-channel units of the input tile per thread
  static constexpr int OUT_ROWS = 256;               // output rows per tile (both kernels)
};

// dst[(g0 + row*gstep) * 32 + half*16 ..] <- v of the thread owning tile row `row` (= lane), rows with g < L only
__device__ __forceinline__ void warp_store_rows_step(float* __restrict__ dst, long batch_off, int g0, int gstep, int L, int half,
                                                     uint8_t* stg, int lane, const float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<float4*>(stg + stg_off(lane, q)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = k * 32 + lane, row = idx >> 2, q = idx & 3;
    const int g = g0 + row * gstep;
    if (g >= 0 && g < L)
      *reinterpret_cast<float4*>(dst + (batch_off + g) * 32 + half * 16 + q * 4) =
          *reinterpret_cast<const float4*>(stg + stg_off(row, q));
  }
  __syncwarp();
}

template <int S, bool UP>
__device__ __forceinline__ void ct_issue(uint32_t tmem_buf, uint32_t a_base, uint32_t w_base) {
  using Cfg = CtCfg<S, UP>;
  // (weight tap, sub-tile, row offset) per MMA group; UP: phase 0 = {W1 @ +0, W3 @ -1}, phase 1 = {W0 @ +1, W2 @ +0}
  //                                                   DOWN: W0 on odd rows @ -1, W1 even @ 0, W2 odd @ 0, W3 even @ +1
  constexpr int TAPW[4] = {UP ? 1 : 0, UP ? 3 : 1, UP ? 0 : 2, UP ? 2 : 3};
  constexpr int TSUB[4] = {UP ? 0 : 1, 0, UP ? 0 : 1, 0};
  constexpr int TOFF[4] = {UP ? 1 : 0, UP ? 0 : 1, UP ? 2 : 1, UP ? 1 : 2};  // 1 + shift (guard row in front)
  const uint64_t ad0 = smem_desc(a_base, Cfg::PLANE, 128);
  const uint64_t bd0 = smem_desc(w_base, Cfg::WPLANE, 128);
  constexpr int NG = UP ? 2 : 1;        // accumulation groups (output phases)
  constexpr int TPG = 4 / NG;           // taps per group
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    uint32_t acc = 0;
#pragma unroll
    for (int tt = 0; tt < TPG; ++tt) {
      const int n = g * TPG + tt;
#pragma unroll
      for (int kk = 0; kk < Cfg::KSTEPS; ++kk)
#pragma unroll
        for (int sa = 0; sa < S; ++sa) {  // widest first (initialises every column block); piece sa meets S - sa weight pieces
          const uint32_t idesc = instr_desc(FMT_BF16, 128, 32 * (S - sa), false, false);
          const uint64_t bd = bd0 + (uint64_t)((TAPW[n] * Cfg::WTAP + kk * 2 * Cfg::WPLANE) >> 4);
          const uint64_t ad = ad0 + (uint64_t)(((TSUB[n] * S + sa) * Cfg::TILE + kk * 2 * Cfg::PLANE) >> 4) + (uint64_t)TOFF[n];
          if (UP) {
            mma<false>(tmem_buf + g * Cfg::NW, ad, bd, idesc, acc);
          } else {
            mma<false>(tmem_buf, ad, bd, idesc, acc);                        // M block 0
            mma<false>(tmem_buf + Cfg::NW, ad + (uint64_t)128, bd, idesc, acc);  // M block 1: 128 rows further
          }
          acc = 1;
        }
    }
  }
}

template <int S, bool UP>
__global__ void __launch_bounds__(CtCfg<S, UP>::NT, 1) conv_tc_kernel(const ConvTcParams p) {
  using Cfg = CtCfg<S, UP>;
  constexpr int NT = Cfg::NT, NU = Cfg::NU;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* A = smem;                                  // [NSUB][S pieces][4 planes]
  uint8_t* W = A + Cfg::NSUB * S * Cfg::TILE;         // [4 taps][4 planes][NW rows]
  uint8_t* stg_base = W + Cfg::WALL;
  uint64_t* bar = reinterpret_cast<uint64_t*>(stg_base + Cfg::STG);  // bar[b]: MMAs into TMEM buffer b done
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);
  float* bias_s = reinterpret_cast<float*>(bar + 4);  // [32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tmem_alloc(tslot, Cfg::TCOLS);
  if (tid == 32) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  {  // all weight loads of this thread first (independent, in flight together), then the packing from registers
    constexpr int PER = 4 * 32 * 32 / NT;
    float wv[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int e = tid + q * NT, n = e & 31, k = (e >> 5) & 31, j = e >> 10;
      wv[q] = p.w[(size_t)j * p.sj + (size_t)k * p.si + (size_t)n * p.so];
    }
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int e = tid + q * NT, n = e & 31, k = (e >> 5) & 31, j = e >> 10;
      uint8_t* a = W + j * Cfg::WTAP + (k >> 3) * Cfg::WPLANE + n * 16 + (k & 7) * 2;
      float pc[3];
      split_bf16<S>(wv[q], pc);
#pragma unroll
      for (int s = 0; s < S; ++s) *reinterpret_cast<__nv_bfloat16*>(a + s * 32 * 16) = __float2bfloat16_rn(pc[s]);  // row s*32 + n
    }
  }
  if (tid < 32) bias_s[tid] = p.bias ? p.bias[tid] : 0.f;
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tslot;
  pdl_launch_dependents();  // the prologue read only this layer's weights (common.cuh)
  pdl_wait();
  const int ntiles = (int)blockIdx.x < p.total_tiles ? (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  // epilogue role: TMEM lane quadrant, accumulator group (DOWN: M block, UP: output phase), 16-channel half
  const int qd = warp & 3, grp = (warp >> 2) & 1, half = warp >> 3;
  uint8_t* stg = stg_base + warp * 2048;
  const int oct = tid & 3;

  float4 ra[NU], rb[NU];
  // tile `it` of this CTA covers output rows [o0, o0 + 256) of batch item b; staged input rows start at g_in0
  auto coords = [&](int it, int& b, int& o0) {
    const int tile = blockIdx.x + it * gridDim.x;
    b = tile / p.tiles_x;
    o0 = (tile - b * p.tiles_x) * Cfg::OUT_ROWS;
  };
  auto load = [&](int it) {
    int b, o0;
    coords(it, b, o0);
    // DOWN: output row t needs input rows 2t-1 .. 2t+2: staged rows 2(t0-1) .. ; UP: output rows 2u, 2u+1 need u-1 .. u+1
    const int g0 = UP ? o0 / 2 - 1 : 2 * (o0 - 1);
    const float* inb = p.in + (size_t)b * p.Lin * 32 + oct * 8;
#pragma unroll
    for (int k = 0; k < NU; ++k) {
      const int rr = (tid + k * NT) >> 2;
      const int g = g0 + rr;
      const bool ok = rr < Cfg::NSUB * Cfg::AROWS && g >= 0 && g < p.Lin;
      ra[k] = ok ? *reinterpret_cast<const float4*>(inb + (long)g * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[k] = ok ? *reinterpret_cast<const float4*>(inb + (long)g * 32 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto convert = [&]() {
#pragma unroll
    for (int k = 0; k < NU; ++k) { reg_fence(ra[k]); reg_fence(rb[k]); }
#pragma unroll
    for (int k = 0; k < NU; ++k) {
      const int rr = (tid + k * NT) >> 2;
      if (rr < Cfg::NSUB * Cfg::AROWS) {
        const int sub = UP ? 0 : (rr & 1), ru = UP ? rr : (rr >> 1);
        uint4 pc[S];
        split8<S>(ra[k], rb[k], pc);
#pragma unroll
        for (int s = 0; s < S; ++s) *reinterpret_cast<uint4*>(A + (sub * S + s) * Cfg::TILE + oct * Cfg::PLANE + ru * 16) = pc[s];
      }
    }
    fence_proxy_async();
  };

  if (ntiles > 0) {
    load(0);
    convert();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 1 && elect_one()) {
      ct_issue<S, UP>(tmem, smem_u32(A), smem_u32(W));
      commit(&bar[0]);
    }
    __syncwarp();
    if (ntiles > 1) load(1);
  }
#pragma unroll 1
  for (int it = 0; it < ntiles; ++it) {
    const int buf = it & 1;
    mbar_wait(&bar[buf], (it >> 1) & 1);
    fence_after_sync();
    if (it + 1 < ntiles) {  // the input tile is free (its MMAs completed): stage the next tile and start its MMAs
      convert();
      fence_before_sync();
      __syncthreads();      // also: every warp finished reading the other TMEM buffer (epilogue of tile it-1)
      fence_after_sync();
      if (warp == 1 && elect_one()) {
        ct_issue<S, UP>(tmem + (buf ^ 1) * Cfg::ACC, smem_u32(A), smem_u32(W));
        commit(&bar[buf ^ 1]);
      }
      __syncwarp();
      if (it + 2 < ntiles) load(it + 2);
    }
    // ---- epilogue: TMEM -> (+bias) -> out
    int b, o0;
    coords(it, b, o0);
    const uint32_t taddr = tmem + (((uint32_t)qd * 32u) << 16) + (uint32_t)(buf * Cfg::ACC + grp * Cfg::NW + half * 16);
    float v[16], m[16];
    tmem_ld16(taddr, v);
#pragma unroll
    for (int sp = 1; sp < S; ++sp) {
      tmem_ld16(taddr + sp * 32, m);
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] += m[c];
    }
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(bias_s + half * 16 + c);
      v[c] += bv.x; v[c + 1] += bv.y; v[c + 2] += bv.z; v[c + 3] += bv.w;
    }
    // the warp's 32 tile rows map to output rows: DOWN o0 + grp*128 + qd*32 + row; UP o0 + 2*(qd*32 + row) + grp
    const int g0 = UP ? o0 + 2 * (qd * 32) + grp : o0 + grp * 128 + qd * 32;
    warp_store_rows_step(p.out, (long)b * p.Lout, g0, UP ? 2 : 1, p.Lout, half, stg, lane, v);
    fence_before_sync();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, Cfg::TCOLS);
}

template <int S, bool UP>
static int launch_ct(const ConvTcParams& p, cudaStream_t st) {
  using Cfg = CtCfg<S, UP>;
  static bool attr_set = false;
  if (!attr_set) {
    VQB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<S, UP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    VQB_CUDA(cudaGetDevice(&dev));
    VQB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  ConvTcParams q = p;
  q.tiles_x = cdiv(p.Lout, Cfg::OUT_ROWS);
  q.total_tiles = q.tiles_x * p.B;
  if (q.total_tiles == 0) return VQB_OK;
  const int grid = q.total_tiles < num_sms ? q.total_tiles : num_sms;
  VQB_CUDA(launch_pdl(conv_tc_kernel<S, UP>, dim3(grid), dim3(Cfg::NT), (size_t)Cfg::SMEM, st, q));
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

template <bool UP>
static int dispatch_ct(int precision, const ConvTcParams& p, cudaStream_t st) {
  switch (precision) {
    case VQB_PREC_BF16: return launch_ct<1, UP>(p, st);
    case VQB_PREC_BF16X2: return launch_ct<2, UP>(p, st);
    case VQB_PREC_BF16X3:
    case VQB_PREC_FP16X2: return launch_ct<3, UP>(p, st);  // fp16x2 is a residual-block mode: bf16x3 here
  }
  return set_err(VQB_ERR_INVALID, "tensor-core strided convolution: precision %d has no kernel", precision);
}

// k = 4, stride 2, 32 -> 32 channels, no fused ReLU, bf16-family precision
bool conv_tc_supported(const vqb_conv_desc* d) {
  return d->k == 4 && d->stride == 2 && d->dilation == 1 && d->C_in == 32 && d->C_out == 32 && !d->relu_in &&
         (d->precision == VQB_PREC_BF16 || d->precision == VQB_PREC_BF16X2 || d->precision == VQB_PREC_BF16X3 ||
          d->precision == VQB_PREC_FP16X2);
}

// Conv1D forward: y [B, ceil(L/2), 32]
int conv1d_fwd_tc(const vqb_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
  ConvTcParams p{};
  p.in = x; p.out = y; p.w = w; p.bias = bias;
  p.sj = 32 * 32; p.si = 32; p.so = 1;          // W_j[ci][co]
  p.B = d->B; p.Lin = d->L; p.Lout = (d->L + 1) / 2;
  return dispatch_ct<false>(d->precision, p, st);
}
// Conv1D data gradient: dx [B, L, 32] from dy [B, ceil(L/2), 32]
int conv1d_dgrad_tc(const vqb_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st) {
  ConvTcParams p{};
  p.in = dy; p.out = dx; p.w = w; p.bias = nullptr;
  p.sj = 32 * 32; p.si = 1; p.so = 32;          // in-channel = co, out-channel = ci: w[j][ci][co]
  p.B = d->B; p.Lin = (d->L + 1) / 2; p.Lout = d->L;
  return dispatch_ct<true>(d->precision, p, st);
}
// Conv1DTranspose forward: y [B, 2L, 32]; kernel [k, Cout, Cin]
int conv1d_transpose_fwd_tc(const vqb_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
  ConvTcParams p{};
  p.in = x; p.out = y; p.w = w; p.bias = bias;
  p.sj = 32 * 32; p.si = 1; p.so = 32;          // in-channel = ci, out-channel = co: w[k][co][ci]
  p.B = d->B; p.Lin = d->L; p.Lout = 2 * d->L;
  return dispatch_ct<true>(d->precision, p, st);
}
// Conv1DTranspose data gradient: dx [B, L, 32] from dy [B, 2L, 32]
int conv1d_transpose_dgrad_tc(const vqb_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st) {
  ConvTcParams p{};
  p.in = dy; p.out = dx; p.w = w; p.bias = nullptr;
  p.sj = 32 * 32; p.si = 32; p.so = 1;          // in-channel = co, out-channel = ci: w[k][co][ci]
  p.B = d->B; p.Lin = 2 * d->L; p.Lout = d->L;
  return dispatch_ct<false>(d->precision, p, st);
}

}  // namespace vqb
