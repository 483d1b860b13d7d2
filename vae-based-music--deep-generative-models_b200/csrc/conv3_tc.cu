// conv3_tc.cu — the latent-rate k = 3, stride-1 convolutions between the 32-channel stacks and the 64-wide latent
// (encdec.py:38: EncoderConvBlock's projection Conv1D(output_emb_width, 3, 1), 32 -> 64; encdec.py:60: DecoderConvBlock's first
// Conv1D(width, 3, 1), 64 -> 32) and their data gradients, as tcgen05 implicit GEMMs:
//     out[b, t, n] = bias[n] + sum_j sum_k in[b, t + (j - 1) * dil, k] * Wt(j, k, n)          (SAME padding)
//   M = 128 time rows per tile, K = CIN per tap, N = COUT; the three taps are row-shifted views of ONE operand tile (plane
//   layout of tc.cuh).  bf16 pieces as in conv_tc.cu: the activation pieces a_0..a_{S-1} meet the weight pieces stacked along N,
//   a_s x [w_0 | .. | w_{S-1-s}], so one accumulator of S * COUT columns collects every product with s_a + s_w < S and the
//   epilogue adds its S column blocks.  The data gradient is the same kernel on the transposed, tap-flipped weights.
// These layers see 1/8 .. 1/64 of the audio rate (a few hundred tiles per call), so the kernel is kept simple: no persistence
// across phases, two CTAs per SM overlap one tile's loads and epilogue with the other's MMAs.  It replaces the exact-fp32
// CUDA-core kernel (tgc_kernel, 20-30 us per call at FP32 peak) in the tensor-core precision modes.
#include "common.cuh"
#include "tc.cuh"

namespace vqb {

using namespace tc;

struct Conv3Params {
  const float* in;    // [B, L, CIN]
  float* out;         // [B, L, COUT]
  const float* w;     // element (tap j, in-channel k, out-channel n) at w[(flip ? 2 - j : j) * sj + k * si + n * so]
  const float* bias;  // [COUT] or NULL
  long sj, si, so;
  int flip;
  int B, L, dil, tiles_per_b, total_tiles;
};

template <int S_, int CIN_, int COUT_>
struct C3Cfg {
  static constexpr int S = S_, CIN = CIN_, COUT = COUT_;
  static constexpr int NT = 256;
  static constexpr int TM = 128;                        // output rows per tile
  static constexpr int DMAX = 8;                        // largest dilation
  static constexpr int NPA = CIN / 8;                   // 16-byte planes (8 bf16 channels) per activation piece
  static constexpr int ROWS = TM + 2 * DMAX;
  static constexpr int PA = ROWS * 16 + 32;             // plane pitch
  static constexpr int TILE_A = NPA * PA;               // one piece
  static constexpr int NW = S * COUT;                   // weight rows per plane: pieces stacked along N
  static constexpr int WPLANE = NW * 16;
  static constexpr int WTAP = NPA * WPLANE;
  static constexpr int OFF_W = ((S * TILE_A + 127) / 128) * 128;
  static constexpr int SP = COUT + 4;                   // row pitch (floats) of the epilogue's staging tile: conflict-free float4 rows
  static constexpr int OFF_STG = OFF_W + 3 * WTAP;
  static constexpr int SMEM = OFF_STG + TM * SP * 4 + 128;
  static constexpr int TCOLS = NW <= 32 ? 32 : NW <= 64 ? 64 : NW <= 128 ? 128 : 256;
  static constexpr int UNITS = (ROWS * NPA + NT - 1) / NT;  // 8-channel units of the input tile per thread
  static_assert(2 * SMEM <= 232448 - 2048, "two CTAs per SM");
  static_assert(NW <= 256 && COUT % 16 == 0 && CIN % 16 == 0, "shape");
};

template <int S, int CIN, int COUT>
__global__ void __launch_bounds__(256, 2) conv3_tc_kernel(const Conv3Params p) {
  using Cfg = C3Cfg<S, CIN, COUT>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* A = smem;
  uint8_t* W = smem + Cfg::OFF_W;
  float* stg = reinterpret_cast<float*>(smem + Cfg::OFF_STG);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) tmem_alloc(&tslot, Cfg::TCOLS);
  if (tid == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
  // weight operand image: [tap][plane = k / 8][row = piece * COUT + n][8 in-channels]
  for (int e = tid; e < 3 * CIN * COUT; e += Cfg::NT) {
    int j, k, n;
    if (p.so == 1) { n = e % COUT; k = (e / COUT) % CIN; j = e / (COUT * CIN); }   // walk the contiguous index fastest
    else { k = e % CIN; n = (e / CIN) % COUT; j = e / (COUT * CIN); }
    const float wv = p.w[(long)(p.flip ? 2 - j : j) * p.sj + (long)k * p.si + (long)n * p.so];
    float pc[3];
    split_bf16<S>(wv, pc);
    uint8_t* a = W + j * Cfg::WTAP + (k >> 3) * Cfg::WPLANE + n * 16 + (k & 7) * 2;
#pragma unroll
    for (int s = 0; s < S; ++s) *reinterpret_cast<__nv_bfloat16*>(a + s * COUT * 16) = __float2bfloat16_rn(pc[s]);
  }
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tslot;
  pdl_launch_dependents();
  pdl_wait();  // the input is the previous kernel's output

  uint32_t phase = 0;
  const int dil = p.dil, rows_in = Cfg::TM + 2 * dil;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    const int b = tile / p.tiles_per_b;
    const int t0 = (tile - b * p.tiles_per_b) * Cfg::TM;
    const float* inb = p.in + (size_t)b * p.L * CIN;
    // ---- input rows t0 - dil .. t0 + 127 + dil -> bf16 pieces in the plane layout (all loads of a thread in flight together)
    float4 ra[Cfg::UNITS][2];
#pragma unroll
    for (int u = 0; u < Cfg::UNITS; ++u) {
      const int idx = tid + u * Cfg::NT, row = idx / Cfg::NPA, pl = idx - row * Cfg::NPA, g = t0 - dil + row;
      const bool ok = row < rows_in && g >= 0 && g < p.L;
      const float* src = inb + (long)g * CIN + pl * 8;
      ra[u][0] = ok ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
      ra[u][1] = ok ? *reinterpret_cast<const float4*>(src + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < Cfg::UNITS; ++u) {
      const int idx = tid + u * Cfg::NT, row = idx / Cfg::NPA, pl = idx - row * Cfg::NPA;
      if (row < rows_in) {
        uint4 pc[S];
        split8<S>(ra[u][0], ra[u][1], pc);
#pragma unroll
        for (int s = 0; s < S; ++s) *reinterpret_cast<uint4*>(A + s * Cfg::TILE_A + pl * Cfg::PA + row * 16) = pc[s];
      }
    }
    fence_proxy_async();
    fence_before_sync();  // the previous tile's epilogue reads of tensor memory precede this tile's MMAs
    __syncthreads();
    // ---- MMAs: tap j of output row m reads tile row m + j * dil
    if (warp == 0) {
      if (elect_one()) {
        fence_after_sync();
        const uint32_t a_base = smem_u32(A), w_base = smem_u32(W);
        uint32_t acc = 0u;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
          for (int kk = 0; kk < CIN / 16; ++kk)
#pragma unroll
            for (int sa = 0; sa < S; ++sa) {
              const uint32_t idesc = instr_desc(FMT_BF16, 128, COUT * (S - sa), false, false);
              const uint64_t ad = smem_desc(a_base + sa * Cfg::TILE_A + kk * 2 * Cfg::PA, Cfg::PA, 128) + (uint64_t)(j * dil);
              const uint64_t bd = smem_desc(w_base + j * Cfg::WTAP + kk * 2 * Cfg::WPLANE, Cfg::WPLANE, 128);
              mma<false>(tmem, ad, bd, idesc, acc);
              acc = 1u;
            }
        commit(&bar);
      }
      __syncwarp();
    }
    mbar_wait(&bar, phase);
    phase ^= 1u;
    fence_after_sync();
    // ---- epilogue: thread = tile row (warps 0-3): sum of the S column blocks + bias -> staging tile -> coalesced rows
    if (warp < 4) {
      const uint32_t ta = tmem + (((uint32_t)warp * 32u) << 16);
      float* srow = stg + tid * Cfg::SP;
#pragma unroll
      for (int c0 = 0; c0 < COUT; c0 += 32) {
        float v[32], m[32];
        tmem_ld32(ta + c0, v);
#pragma unroll
        for (int s = 1; s < S; ++s) {
          tmem_ld32(ta + s * COUT + c0, m);
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] += m[c];
        }
        if (p.bias) {
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] += __ldg(p.bias + c0 + c);
        }
#pragma unroll
        for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(srow + c0 + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
      }
      fence_before_sync();
    }
    __syncthreads();
    float* outb = p.out + ((size_t)b * p.L + t0) * COUT;
    const int nrows = min(Cfg::TM, p.L - t0);
    for (int e = tid; e < nrows * (COUT / 4); e += Cfg::NT) {
      const int row = e / (COUT / 4), q = e - row * (COUT / 4);
      *reinterpret_cast<float4*>(outb + (size_t)row * COUT + q * 4) = *reinterpret_cast<const float4*>(stg + row * Cfg::SP + q * 4);
    }
    // the next tile's loads / conversions may start at once: A was consumed by the MMAs, the staging tile is rewritten only after
    // that tile's own barrier
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, Cfg::TCOLS);
}

static bool c3_prec(int precision) {
  return precision == VQB_PREC_BF16 || precision == VQB_PREC_BF16X2 || precision == VQB_PREC_BF16X3 || precision == VQB_PREC_FP16X2;
}

// k = 3, stride 1, 32 -> 64 or 64 -> 32 channels, dilation <= 8, no fused ReLU, bf16-family precision
bool conv3_tc_supported(const vqb_conv_desc* d) {
  return d->k == 3 && d->stride == 1 && d->dilation >= 1 && d->dilation <= C3Cfg<1, 32, 64>::DMAX && !d->relu_in &&
         ((d->C_in == 32 && d->C_out == 64) || (d->C_in == 64 && d->C_out == 32)) && c3_prec(d->precision);
}

template <int S, int CIN, int COUT>
static int launch_c3(const Conv3Params& p, cudaStream_t st) {
  using Cfg = C3Cfg<S, CIN, COUT>;
  static bool attr_set = false;
  if (!attr_set) {
    VQB_CUDA(cudaFuncSetAttribute(conv3_tc_kernel<S, CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    VQB_CUDA(cudaGetDevice(&dev));
    VQB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = p.total_tiles < 2 * num_sms ? p.total_tiles : 2 * num_sms;
  VQB_CUDA(launch_pdl(conv3_tc_kernel<S, CIN, COUT>, dim3(grid), dim3(Cfg::NT), (size_t)Cfg::SMEM, st, p));
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

template <int CIN, int COUT>
static int dispatch_c3(int precision, const Conv3Params& p, cudaStream_t st) {
  switch (precision) {
    case VQB_PREC_BF16: return launch_c3<1, CIN, COUT>(p, st);
    case VQB_PREC_BF16X2: return launch_c3<2, CIN, COUT>(p, st);
    case VQB_PREC_BF16X3:
    case VQB_PREC_FP16X2: return launch_c3<3, CIN, COUT>(p, st);  // fp16x2 is a residual-stack mode: bf16x3 here
  }
  return set_err(VQB_ERR_INVALID, "tensor-core k=3 convolution: precision %d has no kernel", precision);
}

static int run_c3(const vqb_conv_desc* d, Conv3Params& p, int cin, cudaStream_t st) {
  p.B = d->B; p.L = d->L; p.dil = d->dilation;
  if (d->B == 0 || d->L == 0) return VQB_OK;
  p.tiles_per_b = cdiv(d->L, 128);
  p.total_tiles = p.tiles_per_b * d->B;
  return cin == 32 ? dispatch_c3<32, 64>(d->precision, p, st) : dispatch_c3<64, 32>(d->precision, p, st);
}

// Conv1D forward: y [B, L, C_out]; kernel [3, C_in, C_out]
int conv3_fwd_tc(const vqb_conv_desc* d, const float* x, const float* w, const float* bias, float* y, cudaStream_t st) {
  Conv3Params p{};
  p.in = x; p.out = y; p.w = w; p.bias = bias;
  p.sj = (long)d->C_in * d->C_out; p.si = d->C_out; p.so = 1; p.flip = 0;
  return run_c3(d, p, d->C_in, st);
}
// Conv1D data gradient: dx [B, L, C_in] from dy [B, L, C_out]: the convolution of dy with the transposed, tap-flipped kernel
int conv3_dgrad_tc(const vqb_conv_desc* d, const float* dy, const float* w, float* dx, cudaStream_t st) {
  Conv3Params p{};
  p.in = dy; p.out = dx; p.w = w; p.bias = nullptr;
  p.sj = (long)d->C_in * d->C_out; p.si = 1; p.so = d->C_out; p.flip = 1;  // in-channel = co, out-channel = ci: w[j][ci][co]
  return run_c3(d, p, d->C_out, st);
}

}  // namespace vqb
