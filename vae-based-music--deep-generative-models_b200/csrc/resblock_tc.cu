// resblock_tc.cu — the pre-activation residual block (resnet.py:11-18,29) and its data gradient as ONE fused tcgen05
// kernel each: a chain of two k=3 32->32 convolutions whose intermediate never leaves the SM.
//
//   stage 1:  out1 = conv_{d1}(act1(in1)) + bias1          (* (mask1 > 0))           -> stored (h / dh)
//   stage 2:  out2 = conv_{d2}(act2(out1)) + bias2         (* (mask2 > 0)) (+ add2)  -> stored (y / dx)
//   forward : in1 = x, act1 = ReLU, d1 = dilation, act2 = ReLU, d2 = 1, add2 = x
//   backward: in1 = dy, d1 = 1, W = conv2^T (taps flipped), mask1 = h, d2 = dilation, W = conv1^T, mask2 = x, add2 = dy
//
// Each convolution is an implicit GEMM on the 5th-gen tensor cores: M = 128 time positions per MMA (two M blocks per
// CTA), N = 32 output channels, K = 32 input channels per tap; the three taps are three accumulating groups of MMAs
// whose A descriptors are the SAME shared-memory tile shifted by (tap-1)*dilation rows (tc.cuh, "plane layout").
// Operands are bf16 (kind::f16) or tf32 (kind::tf32), accumulation is fp32 in TMEM; bias, ReLU / masks, the fp32
// residual add and the conversion of the intermediate to the next A operand happen in the TMEM->register epilogue.
// Activations are converted on the fly while being staged (fp32 global -> act -> bf16/tf32 shared), which is why the
// A tile is written by threads rather than by TMA.
#include "common.cuh"
#include "tc.cuh"
#include "tc_rows.cuh"

namespace vqb {

using namespace tc;

struct RbTcParams {
  const float* in1;
  const float* mask1;
  float* out1;
  const float* mask2;
  const float* add2;
  float* out2;
  const float* w1;  // conv of stage 1: element (tap j, in-channel k, out-channel n) at w1[jj*sj1 + k*si1 + n*so1], jj = flip ? 2-j : j
  const float* w2;
  const float* bias1;
  const float* bias2;
  int sj1, si1, so1, flip1;
  int sj2, si2, so2, flip2;
  int B, L, d1, d2, relu1, relu2;
  int tiles_x, total_tiles;  // time tiles per batch item, B * tiles_x
};

// MODE 0: bf16   1: tf32   2: bf16x2 (operands split hi+lo, 3 MMAs per product, ~2^-16)   3: bf16x3 (hi+mid+lo, 6 MMAs,
// all 24 mantissa bits of both operands: fp32-grade products with fp32 accumulation)
template <int MODE>
struct RbCfg {
  static constexpr bool TF32 = MODE == 1;
  static constexpr int S = MODE == 2 ? 2 : MODE == 3 ? 3 : 1;  // bf16 pieces per operand
  static constexpr int R = 256;     // stage-1 rows per CTA (2 x M128)
  static constexpr int NT = 512;    // threads: two per tile row (one per 16-channel half)
  static constexpr int DMAX = 32;   // largest supported dilation (guard rows of the operand tiles)
  static constexpr int C = 32;
  static constexpr int ES = TF32 ? 4 : 2;
  static constexpr int T = 16 / ES;
  static constexpr int NP = C / T;          // planes per operand tile
  static constexpr int KSTEPS = NP / 2;     // one MMA consumes 32 bytes of K = 2 planes
  static constexpr int PLANE = (R + 2 * DMAX) * 16 + (TF32 ? 16 : 32);  // bytes; padding de-aliases the planes' banks
  static constexpr int NW = S * 32;         // MMA N: the S bf16 pieces of the weights are stacked along N
  static constexpr int WPLANE = NW * 16;
  static constexpr int WTAP = NP * WPLANE;
  static constexpr int WCONV = 3 * WTAP;
  static constexpr int TILE = NP * PLANE;   // one operand tile (one split piece)
  static constexpr int STG = (NT / 32) * 2048;  // per-warp row staging (32 rows x 64 B)
  static constexpr int SMEM = 2 * S * TILE + 2 * WCONV + STG + 64 + 256;
  static constexpr int TCOLS = 4 * NW <= 128 ? 128 : 4 * NW <= 256 ? 256 : 512;  // TMEM columns: 2 stages x 2 M blocks x NW
  static constexpr int NCV = NT - 64;   // threads that load / convert the stage-1 input (all but the two MMA-issuing warps)
  static constexpr int NU = ((R + 2 * DMAX) * 4 + NCV - 1) / NCV;  // 8-channel units of the stage-1 input tile per converter thread
};

template <int MODE>
__device__ __forceinline__ void pack_weights(uint8_t* dst, const float* __restrict__ w, int sj, int si, int so, int flip) {
  using Cfg = RbCfg<MODE>;
  for (int e = threadIdx.x; e < 3 * 32 * 32; e += blockDim.x) {
    const int n = e & 31, k = (e >> 5) & 31, j = e >> 10;
    const int jj = flip ? 2 - j : j;
    const float v = w[(size_t)jj * sj + (size_t)k * si + (size_t)n * so];
    uint8_t* a = dst + j * Cfg::WTAP + (k / Cfg::T) * Cfg::WPLANE + n * 16 + (k % Cfg::T) * Cfg::ES;
    if (Cfg::TF32) {
      *reinterpret_cast<float*>(a) = to_tf32(v);
    } else {
      float pc[3];
      split_bf16<Cfg::S>(v, pc);
#pragma unroll
      for (int s = 0; s < Cfg::S; ++s) *reinterpret_cast<__nv_bfloat16*>(a + s * 32 * 16) = __float2bfloat16_rn(pc[s]);  // row s*32 + n
    }
  }
}

// 3 taps x KSTEPS x S accumulating MMAs for ONE of the two M blocks of a stage (two threads of different warps issue the
// two blocks concurrently: the issue loop is on every warp's critical path between two block barriers).  In the split modes
// the S weight pieces are stacked along N and activation piece `sa` is multiplied with the first S - sa of them into the
// SAME accumulator (hi x {hi, mid, lo}, mid x {hi, mid}, lo x {hi}: every product down to 2^-16 of the leading one plus
// mid x mid; the three omitted ones are below 2^-23): column block c then holds the sum over the activation pieces of
// a . W_c, and the epilogue adds the S column blocks.
template <int MODE>
__device__ __forceinline__ void issue_stage(uint32_t tmem, uint32_t a_base, int row_shift0, int dil, uint32_t w_base, int mb) {
  using Cfg = RbCfg<MODE>;
  // descriptors differ only in their start-address field (units of 16 bytes = one tile row): add offsets to two bases
  const uint64_t ad0 = smem_desc(a_base + (uint32_t)row_shift0 * 16u, Cfg::PLANE, 128);
  const uint64_t bd0 = smem_desc(w_base, Cfg::WPLANE, 128);
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int kk = 0; kk < Cfg::KSTEPS; ++kk)
#pragma unroll
      for (int sa = 0; sa < Cfg::S; ++sa) {  // the widest MMA first: it initialises every column block
        const uint32_t idesc = instr_desc(Cfg::TF32 ? FMT_TF32 : FMT_BF16, 128, 32 * (Cfg::S - sa), false, false);
        const uint64_t bd = bd0 + (uint64_t)((j * Cfg::WTAP + kk * 2 * Cfg::WPLANE) >> 4);
        const uint64_t ad = ad0 + (uint64_t)((sa * Cfg::TILE + kk * 2 * Cfg::PLANE) >> 4) + (uint64_t)(mb * 128 + j * dil);
        mma<Cfg::TF32>(tmem + mb * Cfg::NW, ad, bd, idesc, acc);
        acc = 1;
      }
}

// 8 channels o*8..o*8+7 of row r (two float4) -> operand tile(s): one 16-byte chunk per bf16 piece, two for tf32
template <int MODE>
__device__ __forceinline__ void stage8(uint8_t* tile, int r, int o, const float4& a, const float4& b) {
  using Cfg = RbCfg<MODE>;
  if (Cfg::TF32) {
    *reinterpret_cast<float4*>(tile + (2 * o) * Cfg::PLANE + r * 16) = make_float4(to_tf32(a.x), to_tf32(a.y), to_tf32(a.z), to_tf32(a.w));
    *reinterpret_cast<float4*>(tile + (2 * o + 1) * Cfg::PLANE + r * 16) = make_float4(to_tf32(b.x), to_tf32(b.y), to_tf32(b.z), to_tf32(b.w));
  } else {
    uint4 pc[Cfg::S];
    split8<Cfg::S>(a, b, pc);
#pragma unroll
    for (int s = 0; s < Cfg::S; ++s) *reinterpret_cast<uint4*>(tile + s * Cfg::TILE + o * Cfg::PLANE + r * 16) = pc[s];
  }
}

// Software pipeline over the tiles of a persistent CTA (one CTA per SM, 512 threads, all threads take part in every phase):
//     load(i+1) -> registers | wait MMA1(i) | epilogue 1(i): TMEM -> out1, -> A2 | issue MMA2(i) |
//     convert(i+1): registers -> A1 | issue MMA1(i+1) | wait MMA2(i) | epilogue 2(i): TMEM -> out2
// so the global loads of the next tile fly during epilogue 1, the stage-2 MMAs run under the conversion of the next tile
// and the next tile's stage-1 MMAs under epilogue 2.  A1 / A2 and the two stages' TMEM accumulators are separate buffers.
// (A variant with a dedicated MMA-issuing warp and mbarrier-only hand-offs measured slower: 17 warps cap the register
// file at 96 per thread; so did epilogues that access their rows in global memory directly instead of through the
// per-warp staging transposes: 32 lines per access instruction saturate the L1 pipeline.)
template <int MODE>
__global__ void __launch_bounds__(RbCfg<MODE>::NT, 1) rb_tc_kernel(const RbTcParams p) {
  using Cfg = RbCfg<MODE>;
  constexpr int NT = Cfg::NT, NU = Cfg::NU;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* A1 = smem;
  uint8_t* A2 = A1 + Cfg::S * Cfg::TILE;
  uint8_t* W1 = A2 + Cfg::S * Cfg::TILE;
  uint8_t* W2 = W1 + Cfg::WCONV;
  uint8_t* stg_base = W2 + Cfg::WCONV;
  uint64_t* bar = reinterpret_cast<uint64_t*>(stg_base + Cfg::STG);  // bar[0]: stage-1 MMAs done, bar[1]: stage-2 MMAs done
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);
  float* bias_s = reinterpret_cast<float*>(bar + 4);  // [64]: bias1, bias2 (zeros when absent); 16-byte aligned

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = p.L;
  const int Rout = Cfg::R - 2 * p.d2;
  const int rows1 = Cfg::R + 2 * p.d1;

  if (warp == 0) tmem_alloc(tslot, Cfg::TCOLS);
  if (tid == 32) { mbar_init(&bar[0], 2); mbar_init(&bar[1], 2); fence_mbar_init(); }  // one arrival per issuing thread
  pack_weights<MODE>(W1, p.w1, p.sj1, p.si1, p.so1, p.flip1);
  pack_weights<MODE>(W2, p.w2, p.sj2, p.si2, p.so2, p.flip2);
  if (tid < 64) bias_s[tid] = tid < 32 ? (p.bias1 ? p.bias1[tid] : 0.f) : (p.bias2 ? p.bias2[tid - 32] : 0.f);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tslot;

  // epilogue role of this thread: TMEM lane quadrant (warp % 4), M block, 16-channel half
  const int qd = warp & 3, mb = (warp >> 2) & 1, half = warp >> 3;
  const int i0 = mb * 128 + qd * 32;   // first tile row of this warp
  const int i = i0 + lane;             // tile row (= TMEM lane) of this thread
  const uint32_t taddr = tmem + (((uint32_t)qd * 32u) << 16) + (uint32_t)(mb * Cfg::NW + half * 16);
  uint8_t* stg = stg_base + warp * 2048;
  // warps 1 and 5 issue the MMAs (one M block each) while the other 14 warps stage the next tile: they take no part in load /
  // convert, so that both groups reach the barrier behind the conversion at about the same time
  const bool issuer = warp == 1 || warp == 5;
  const int ctid = (warp - (warp > 1) - (warp > 5)) * 32 + lane;  // index among the converter threads
  const int oct = ctid & 3;  // 8-channel unit of this thread in the staging loops (NCV % 4 == 0)
  constexpr int NCV = Cfg::NCV;

  float4 ra[NU], rb[NU];
  // global rows of tile `tile`: A1 row r holds in1 row g1 + r
  auto load = [&](int tile) {
    if (issuer) return;
    const int b = tile / p.tiles_x;
    const int g1 = (tile - b * p.tiles_x) * Rout - p.d2 - p.d1;
    const float* inb = p.in1 + (size_t)b * L * 32 + oct * 8;
#pragma unroll
    for (int k = 0; k < NU; ++k) {
      const int r = (ctid + k * NCV) >> 2;
      const int g = g1 + r;
      const bool ok = r < rows1 && g >= 0 && g < L;
      ra[k] = ok ? *reinterpret_cast<const float4*>(inb + (long)g * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[k] = ok ? *reinterpret_cast<const float4*>(inb + (long)g * 32 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto convert = [&]() {
    if (issuer) return;
#pragma unroll
    for (int k = 0; k < NU; ++k) { reg_fence(ra[k]); reg_fence(rb[k]); }  // keep the conversion below the waits it follows
#pragma unroll
    for (int k = 0; k < NU; ++k) {
      const int r = (ctid + k * NCV) >> 2;
      if (r < rows1) {
        float4 a = ra[k], b = rb[k];
        if (p.relu1) {
          a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
          b.x = fmaxf(b.x, 0.f); b.y = fmaxf(b.y, 0.f); b.z = fmaxf(b.z, 0.f); b.w = fmaxf(b.w, 0.f);
        }
        stage8<MODE>(A1, r, oct, a, b);
      }
    }
    fence_proxy_async();
  };

  int tile = blockIdx.x;
  if (tile < p.total_tiles) {
    load(tile);
    convert();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (tid == 32 || tid == 160) {
      issue_stage<MODE>(tmem, smem_u32(A1), 0, p.d1, smem_u32(W1), tid >> 7);
      commit(&bar[0]);
    }
    __syncwarp();
  }
  uint32_t phase = 0;
#pragma unroll 1
  for (; tile < p.total_tiles; tile += gridDim.x, phase ^= 1) {
    const int b = tile / p.tiles_x;
    const int t0 = (tile - b * p.tiles_x) * Rout;  // first out2 row of this tile
    const int s0 = t0 - p.d2;          // out1 row held by A2 row DMAX (A2 has DMAX guard rows in front)
    const long boff = (long)b * L;
    const int g = s0 + i;              // global row of this thread
    const bool inrange = g >= 0 && g < L;
    const int next = tile + gridDim.x;
    const bool has_next = next < p.total_tiles;

    if (tid == 0 && has_next) {  // L2 hints: the epilogue operands of the next tile and the stage-1 input of the one after it
      const int nb = next / p.tiles_x;
      const int ns0 = (next - nb * p.tiles_x) * Rout - p.d2;
      prefetch_rows(p.mask1, nb, L, ns0, ns0 + Cfg::R);
      prefetch_rows(p.mask2, nb, L, ns0 + p.d2, ns0 + Cfg::R - p.d2);
      if (p.add2 != p.in1) prefetch_rows(p.add2, nb, L, ns0 + p.d2, ns0 + Cfg::R - p.d2);
      const int nn = next + gridDim.x;
      if (nn < p.total_tiles) {
        const int nnb = nn / p.tiles_x;
        const int ng1 = (nn - nnb * p.tiles_x) * Rout - p.d2 - p.d1;
        prefetch_rows(p.in1, nnb, L, ng1, ng1 + rows1);
      }
    }
    if (has_next) load(next);

    float v[16], m[16];
    float4 f[4];
    // ---- epilogue 1: TMEM -> (+bias, mask) -> out1 (global, owned rows) and act2(out1) -> A2 (shared)
    if (p.mask1) warp_fetch_rows(p.mask1, boff, s0 + i0, L, half, lane, f);
    mbar_wait(&bar[0], phase);
    fence_after_sync();
    tmem_ld16(taddr, v);
#pragma unroll
    for (int sp = 1; sp < Cfg::S; ++sp) {  // split modes: add the column blocks of the other weight pieces
      tmem_ld16(taddr + sp * 32, m);
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] += m[c];
    }
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(bias_s + half * 16 + c);
      v[c] += bv.x; v[c + 1] += bv.y; v[c + 2] += bv.z; v[c + 3] += bv.w;
    }
    if (p.mask1) {
      warp_unpack_rows(f, stg, lane, m);
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = m[c] > 0.f ? v[c] : 0.f;
    }
    if (p.out1) warp_store_rows(p.out1, boff, s0 + i0, L, half, i0, p.d2, Cfg::R - p.d2, stg, lane, v);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float a = inrange ? v[c] : 0.f;  // rows outside [0, L) are conv2's zero padding
      v[c] = p.relu2 ? fmaxf(a, 0.f) : a;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q)
      stage8<MODE>(A2, Cfg::DMAX + i, half * 2 + q, make_float4(v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3]),
                   make_float4(v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]));
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();  // A2 complete; every warp has drained the stage-1 accumulators
    fence_after_sync();
    if (tid == 32 || tid == 160) {
      // out2 tile row i uses A2 rows DMAX + i + (j-1)*d2
      issue_stage<MODE>(tmem + 2 * Cfg::NW, smem_u32(A2), Cfg::DMAX - p.d2, p.d2, smem_u32(W2), tid >> 7);
      commit(&bar[1]);
    }
    __syncwarp();
    if (has_next) {  // A1 is free (its MMAs completed before epilogue 1): stage the next tile under the stage-2 MMAs
      convert();
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
      if (tid == 32 || tid == 160) {
        issue_stage<MODE>(tmem, smem_u32(A1), 0, p.d1, smem_u32(W1), tid >> 7);
        commit(&bar[0]);
      }
      __syncwarp();
    }

    // ---- epilogue 2: TMEM -> (+bias, mask, + add) -> out2 (owned rows)
    float4 f2[4];
    if (p.mask2) warp_fetch_rows(p.mask2, boff, s0 + i0, L, half, lane, f);
    if (p.add2) warp_fetch_rows(p.add2, boff, s0 + i0, L, half, lane, f2);
    mbar_wait(&bar[1], phase);
    fence_after_sync();
    tmem_ld16(taddr + 2 * Cfg::NW, v);
#pragma unroll
    for (int sp = 1; sp < Cfg::S; ++sp) {
      tmem_ld16(taddr + 2 * Cfg::NW + sp * 32, m);
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] += m[c];
    }
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(bias_s + 32 + half * 16 + c);
      v[c] += bv.x; v[c + 1] += bv.y; v[c + 2] += bv.z; v[c + 3] += bv.w;
    }
    if (p.mask2) {
      warp_unpack_rows(f, stg, lane, m);
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = m[c] > 0.f ? v[c] : 0.f;
    }
    if (p.add2) {
      warp_unpack_rows(f2, stg, lane, m);
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] += m[c];
    }
    warp_store_rows(p.out2, boff, s0 + i0, L, half, i0, p.d2, Cfg::R - p.d2, stg, lane, v);
    fence_before_sync();  // orders this tile's TMEM reads before the barriers of the next iteration
  }  // tile loop
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, Cfg::TCOLS);
}

template <int MODE>
static int launch_rb(const RbTcParams& p, cudaStream_t st) {
  using Cfg = RbCfg<MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    VQB_CUDA(cudaFuncSetAttribute(rb_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    VQB_CUDA(cudaGetDevice(&dev));
    VQB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  RbTcParams q = p;
  const int Rout = Cfg::R - 2 * p.d2;
  q.tiles_x = cdiv(p.L, Rout);
  q.total_tiles = q.tiles_x * p.B;
  const int grid = q.total_tiles < num_sms ? q.total_tiles : num_sms;
  rb_tc_kernel<MODE><<<grid, Cfg::NT, Cfg::SMEM, st>>>(q);
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

static int dispatch_rb(int precision, const RbTcParams& p, cudaStream_t st) {
  switch (precision) {
    case VQB_PREC_BF16: return launch_rb<0>(p, st);
    case VQB_PREC_TF32: return launch_rb<1>(p, st);
    case VQB_PREC_BF16X2: return launch_rb<2>(p, st);
    case VQB_PREC_BF16X3: return launch_rb<3>(p, st);
  }
  return set_err(VQB_ERR_INVALID, "unknown precision %d", precision);
}

bool resblock_tc_supported(const vqb_resblock_desc* d) {
  return d->C == 32 && d->F == 32 && d->dilation >= 1 && d->dilation <= 32 &&
         d->precision >= VQB_PREC_TF32 && d->precision <= VQB_PREC_BF16X3;
}

int resblock_fwd_tc(const vqb_resblock_desc* d, const float* x, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* h, float* y, cudaStream_t st) {
  if (!resblock_tc_supported(d))
    return set_err(VQB_ERR_UNIMPLEMENTED, "tensor-core residual block: C=F=32, dilation<=32, precision bf16|tf32 only (got C=%d F=%d dil=%d prec=%d)",
                   d->C, d->F, d->dilation, d->precision);
  if (d->B == 0 || d->L == 0) return VQB_OK;
  RbTcParams p{};
  p.in1 = x; p.out1 = h; p.add2 = x; p.out2 = y;
  p.w1 = w1; p.bias1 = b1; p.sj1 = 32 * 32; p.si1 = 32; p.so1 = 1; p.flip1 = 0;   // B[n=co][k=ci] = W1[j][ci][co]
  p.w2 = w2; p.bias2 = b2; p.sj2 = 32 * 32; p.si2 = 32; p.so2 = 1; p.flip2 = 0;
  p.B = d->B; p.L = d->L; p.d1 = d->dilation; p.d2 = 1; p.relu1 = 1; p.relu2 = 1;
  return dispatch_rb(d->precision, p, st);
}

int resblock_bwd_tc(const vqb_resblock_desc* d, const float* x, const float* h, const float* dy, const float* w1,
                    const float* w2, float* dh, float* dx, cudaStream_t st) {
  if (!resblock_tc_supported(d))
    return set_err(VQB_ERR_UNIMPLEMENTED, "tensor-core residual block backward: unsupported shape / precision");
  if (d->B == 0 || d->L == 0) return VQB_OK;
  RbTcParams p{};
  p.in1 = dy; p.mask1 = h; p.out1 = dh; p.mask2 = x; p.add2 = dy; p.out2 = dx;
  // stage 1 = conv2^T: dA[t][f] = sum_j sum_c W2[j][f][c] dy[t + (1-j)*1][c]  -> tap n = 2-j, B[n=f][k=c]
  p.w1 = w2; p.sj1 = 32 * 32; p.si1 = 1; p.so1 = 32; p.flip1 = 1;
  // stage 2 = conv1^T with the block's dilation
  p.w2 = w1; p.sj2 = 32 * 32; p.si2 = 1; p.so2 = 32; p.flip2 = 1;
  p.B = d->B; p.L = d->L; p.d1 = 1; p.d2 = d->dilation; p.relu1 = 0; p.relu2 = 0;
  return dispatch_rb(d->precision, p, st);
}

}  // namespace vqb
