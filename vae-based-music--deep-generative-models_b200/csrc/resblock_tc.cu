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
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc.cuh"
#include "tc_rows.cuh"
#include "tma.cuh"

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
  // Sign masks, one uint32 per time position (bit c = channel c is > 0).  The forward kernel can emit them for its input
  // (xbits_out, from the residual operand of epilogue 2) and for its intermediate (hbits_out, epilogue 1), one uint16 per
  // row and 16-channel half; the data-gradient kernel then takes them (m1bits = h > 0, m2bits = x > 0) instead of re-reading
  // the two fp32 tensors: 392 instead of 640 bytes per position and no staged mask fetches.
  uint16_t* xbits_out;
  uint16_t* hbits_out;
  const uint16_t* m1bits;
  const uint16_t* m2bits;
};

// MODE 0: bf16   1: tf32   2: bf16x2 (operands split hi+lo, 3 MMAs per product, ~2^-16)   3: bf16x3 (hi+mid+lo, 6 MMAs,
// all 24 mantissa bits of both operands: fp32-grade products with fp32 accumulation)   4: fp16x2 (operands scaled by a
// power of two — per tile for the activations, per convolution for the weights — and split into two fp16 pieces, 11 + 11
// mantissa bits: fp32-grade products for the MMA count of bf16x2; see tc.cuh)
template <int MODE, int MB = 2, bool TMA_ = false>  // MB = 128-row M blocks per CTA: 2 -> 256-row tiles, 512 threads, one CTA per SM;
struct RbCfg {                   //      1 -> 128-row tiles, 256 threads, two CTAs per SM (modes with <= 2 pieces)
  static constexpr bool TF32 = MODE == 1;
  static constexpr bool F16 = MODE == 4;
  static constexpr int S = (MODE == 2 || MODE == 4) ? 2 : MODE == 3 ? 3 : 1;  // 16-bit pieces per operand
  static constexpr int R = 128 * MB;  // stage-1 rows per CTA
  static constexpr int NT = 256 * MB;  // threads: two per tile row (one per 16-channel half)
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
  // TMA: out1 / out2 leave through dense 128-byte-swizzled row images in shared memory and one bulk-tensor store per
  // tile (tma.cuh) instead of per-warp staging transposes + st.global; needs 2 x R rows of shared memory, which fits for
  // the two-piece modes with 256-row tiles
  static constexpr bool TMA = TMA_ && S == 2 && MB == 2;
  static constexpr int OUTB = TMA ? R * 128 : 0;   // bytes of one output image
  static constexpr int SMEM = 2 * S * TILE + 2 * WCONV + STG + 64 + 256 + (TMA ? 2 * OUTB + 1024 : 0);
  static constexpr int TCOLS = 2 * MB * NW <= 64 ? 64 : 2 * MB * NW <= 128 ? 128 : 2 * MB * NW <= 256 ? 256 : 512;  // TMEM columns: 2 stages x MB M blocks x NW
  static constexpr int NCV = NT - 32 * MB;   // threads that load / convert the stage-1 input (all but the MMA-issuing warps)
  static constexpr int NU = ((R + 2 * DMAX) * 4 + NCV - 1) / NCV;  // 8-channel units of the stage-1 input tile per converter thread
};

// element e = (tap j, in-channel k, out-channel n) of a convolution: its global offset and its place in the operand image
__device__ __forceinline__ size_t weight_offset(int e, int sj, int si, int so, int flip) {
  const int n = e & 31, k = (e >> 5) & 31, j = e >> 10;
  return (size_t)(flip ? 2 - j : j) * sj + (size_t)k * si + (size_t)n * so;
}
template <int MODE, int MB, bool TMA_>
__device__ __forceinline__ void pack_weight(uint8_t* dst, int e, float v, float scale) {
  using Cfg = RbCfg<MODE, MB, TMA_>;
  const int n = e & 31, k = (e >> 5) & 31, j = e >> 10;
  uint8_t* a = dst + j * Cfg::WTAP + (k / Cfg::T) * Cfg::WPLANE + n * 16 + (k % Cfg::T) * Cfg::ES;
  if (Cfg::TF32) {
    *reinterpret_cast<float*>(a) = to_tf32(v);
  } else if (Cfg::F16) {
    const __half hi = __float2half_rn(v * scale);
    const __half lo = __float2half_rn(v * scale - __half2float(hi));
    *reinterpret_cast<__half*>(a) = hi;                // row n
    *reinterpret_cast<__half*>(a + 32 * 16) = lo;      // row 32 + n
  } else {
    float pc[3];
    split_bf16<Cfg::S>(v, pc);
#pragma unroll
    for (int s = 0; s < Cfg::S; ++s) *reinterpret_cast<__nv_bfloat16*>(a + s * 32 * 16) = __float2bfloat16_rn(pc[s]);  // row s*32 + n
  }
}

// 3 taps x KSTEPS x S accumulating MMAs for ONE of the two M blocks of a stage (two threads of different warps issue the
// two blocks concurrently: the issue loop is on every warp's critical path between two block barriers).  In the split modes
// the S weight pieces are stacked along N and activation piece `sa` is multiplied with the first S - sa of them into the
// SAME accumulator (hi x {hi, mid, lo}, mid x {hi, mid}, lo x {hi}: every product down to 2^-16 of the leading one plus
// mid x mid; the three omitted ones are below 2^-23): column block c then holds the sum over the activation pieces of
// a . W_c, and the epilogue adds the S column blocks.
template <int MODE, int MB, bool TMA_>
__device__ __forceinline__ void issue_stage(uint32_t tmem, uint32_t a_base, int row_shift0, int dil, uint32_t w_base, int mb) {
  using Cfg = RbCfg<MODE, MB, TMA_>;
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
        const uint32_t idesc = instr_desc(Cfg::TF32 ? FMT_TF32 : Cfg::F16 ? FMT_F16 : FMT_BF16, 128, 32 * (Cfg::S - sa), false, false);
        const uint64_t bd = bd0 + (uint64_t)((j * Cfg::WTAP + kk * 2 * Cfg::WPLANE) >> 4);
        const uint64_t ad = ad0 + (uint64_t)((sa * Cfg::TILE + kk * 2 * Cfg::PLANE) >> 4) + (uint64_t)(mb * 128 + j * dil);
        mma<Cfg::TF32>(tmem + mb * Cfg::NW, ad, bd, idesc, acc);
        acc = 1;
      }
}

// 8 channels o*8..o*8+7 of row r (two float4) -> operand tile(s): one 16-byte chunk per bf16 piece, two for tf32
template <int MODE, int MB, bool TMA_>
__device__ __forceinline__ void stage8(uint8_t* tile, int r, int o, const float4& a, const float4& b, float scale) {
  using Cfg = RbCfg<MODE, MB, TMA_>;
  if (Cfg::F16) {
    uint4 pc[2];
    split8_f16(a, b, scale, pc);
    *reinterpret_cast<uint4*>(tile + o * Cfg::PLANE + r * 16) = pc[0];
    *reinterpret_cast<uint4*>(tile + Cfg::TILE + o * Cfg::PLANE + r * 16) = pc[1];
  } else if (Cfg::TF32) {
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
// D1MAX: largest stage-1 dilation this instantiation stages (sizes the register prefetch of the input tile: 3 instead of 4
// 8-channel units per thread at 128-row tiles when d1 <= 9, i.e. everywhere but the forward of the dilation-27 blocks)
// KIND: which of the runtime flags are known at compile time (the kernel is instruction-issue bound: every uniform branch and
// dead path costs issue slots).  0 = generic; 1 = training forward (ReLU before both stages, residual add, h stored, both
// sign-mask words written, no input masks); 2 = data gradient from sign-mask words (no ReLU, masks from m1bits / m2bits);
// 3 = inference forward (h not stored, no mask words).
template <int KIND>
__device__ __forceinline__ RbTcParams rb_specialize(RbTcParams q) {
  if (KIND == 1) {
    q.relu1 = 1; q.relu2 = 1; q.mask1 = nullptr; q.mask2 = nullptr; q.m1bits = nullptr; q.m2bits = nullptr;
    __builtin_assume(q.out1 != nullptr); __builtin_assume(q.add2 != nullptr);
    __builtin_assume(q.xbits_out != nullptr); __builtin_assume(q.hbits_out != nullptr);
  } else if (KIND == 2) {
    q.relu1 = 0; q.relu2 = 0; q.mask1 = nullptr; q.mask2 = nullptr; q.xbits_out = nullptr; q.hbits_out = nullptr;
    __builtin_assume(q.out1 != nullptr); __builtin_assume(q.add2 != nullptr);
    __builtin_assume(q.m1bits != nullptr); __builtin_assume(q.m2bits != nullptr);
  } else if (KIND == 3) {  // inference forward: nothing kept for a backward pass
    q.relu1 = 1; q.relu2 = 1; q.mask1 = nullptr; q.mask2 = nullptr; q.m1bits = nullptr; q.m2bits = nullptr;
    q.out1 = nullptr; q.xbits_out = nullptr; q.hbits_out = nullptr;
    __builtin_assume(q.add2 != nullptr);
  }
  return q;
}

template <int MODE, int MB, bool TMA_, int D1MAX, int KIND>
__global__ void __launch_bounds__(RbCfg<MODE, MB, TMA_>::NT, 3 - MB)
    rb_tc_kernel(const RbTcParams pp, const __grid_constant__ CUtensorMap tm_out1, const __grid_constant__ CUtensorMap tm_out2) {
  const RbTcParams p = rb_specialize<KIND>(pp);
  static_assert(D1MAX <= RbCfg<MODE, MB, TMA_>::DMAX, "D1MAX");
  using Cfg = RbCfg<MODE, MB, TMA_>;
  constexpr int NT = Cfg::NT, NU = ((Cfg::R + 2 * D1MAX) * 4 + Cfg::NCV - 1) / Cfg::NCV;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // the swizzled output images must start on a 1024-byte boundary
  uint8_t* smem = Cfg::TMA ? smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u) : smem_raw;
  uint8_t* OUT1 = smem;
  uint8_t* OUT2 = smem + Cfg::OUTB;
  uint8_t* A1 = smem + 2 * Cfg::OUTB;
  uint8_t* A2 = A1 + Cfg::S * Cfg::TILE;
  uint8_t* W1 = A2 + Cfg::S * Cfg::TILE;
  uint8_t* W2 = W1 + Cfg::WCONV;
  uint8_t* stg_base = W2 + Cfg::WCONV;
  uint64_t* bar = reinterpret_cast<uint64_t*>(stg_base + Cfg::STG);  // bar[0]: stage-1 MMAs done, bar[1]: stage-2 MMAs done
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);
  float* bias_s = reinterpret_cast<float*>(bar + 4);  // [64]: bias1, bias2 (zeros when absent); 16-byte aligned
  // fp16x2: [0], [1] largest |stage-1 input| of the tile in flight (alternating slots), [2] max|w1|, [3] max|w2|, [4] max|bias1|
  uint32_t* tmx = reinterpret_cast<uint32_t*>(bias_s + 64);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = p.L;
  const int Rout = Cfg::R - 2 * p.d2;
  const int rows1 = Cfg::R + 2 * p.d1;

  if (warp == 0) tmem_alloc(tslot, Cfg::TCOLS);
  if (tid == 32) { mbar_init(&bar[0], MB); mbar_init(&bar[1], MB); fence_mbar_init(); }  // one arrival per issuing thread
  // Weights: every thread first loads ALL its elements of both convolutions (2 x 3072 / NT independent loads in flight: the
  // prologue costs one L2 round trip instead of one per element), then packs them from registers.
  constexpr int PER = 3 * 32 * 32 / Cfg::NT;
  float wv1[PER], wv2[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    wv1[q] = p.w1[weight_offset(tid + q * Cfg::NT, p.sj1, p.si1, p.so1, p.flip1)];
    wv2[q] = p.w2[weight_offset(tid + q * Cfg::NT, p.sj2, p.si2, p.so2, p.flip2)];
  }
  float sw1 = 1.f, sw2 = 1.f;  // fp16x2: power-of-two scales of the two weight tensors
  if (Cfg::F16) {
    if (tid < 8) tmx[tid] = 0u;
    __syncthreads();
    uint32_t m1 = 0u, m2 = 0u, mb1 = 0u;
#pragma unroll
    for (int q = 0; q < PER; ++q) { m1 = max(m1, absbits(wv1[q])); m2 = max(m2, absbits(wv2[q])); }
    if (tid < 32 && p.bias1) mb1 = absbits(p.bias1[tid]);
    m1 = __reduce_max_sync(0xffffffffu, m1);
    m2 = __reduce_max_sync(0xffffffffu, m2);
    mb1 = __reduce_max_sync(0xffffffffu, mb1);
    if (lane == 0) { atomicMax(&tmx[2], m1); atomicMax(&tmx[3], m2); if (warp == 0) tmx[4] = mb1; }
    __syncthreads();
    sw1 = pow2_scale(__uint_as_float(tmx[2]));
    sw2 = pow2_scale(__uint_as_float(tmx[3]));
  }
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    pack_weight<MODE, MB, TMA_>(W1, tid + q * Cfg::NT, wv1[q], sw1);
    pack_weight<MODE, MB, TMA_>(W2, tid + q * Cfg::NT, wv2[q], sw2);
  }
  if (tid < 64) bias_s[tid] = tid < 32 ? (p.bias1 ? p.bias1[tid] : 0.f) : (p.bias2 ? p.bias2[tid - 32] : 0.f);
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tslot;
  // fp16x2 bound on |out1| used to scale the stage-2 operand without a block-wide reduction:
  // |out1| <= 96 max|w1| max|in1 of the tile| + max|bias1|   (masks and ReLU only shrink it)
  const float w1bound = Cfg::F16 ? 96.f * __uint_as_float(tmx[2]) : 0.f;
  const float b1bound = Cfg::F16 ? __uint_as_float(tmx[4]) : 0.f;
  const float isw1 = pow2_inv(sw1), isw2 = pow2_inv(sw2);

  // epilogue role of this thread: TMEM lane quadrant (warp % 4), M block, 16-channel half
  const int qd = warp & 3, mb = (warp >> 2) & (MB - 1), half = warp / (4 * MB);
  const int i0 = mb * 128 + qd * 32;   // first tile row of this warp
  const int i = i0 + lane;             // tile row (= TMEM lane) of this thread
  const uint32_t taddr = tmem + (((uint32_t)qd * 32u) << 16) + (uint32_t)(mb * Cfg::NW + half * 16);
  uint8_t* stg = stg_base + warp * 2048;
  // warps 1 and 5 issue the MMAs (one M block each) while the other 14 warps stage the next tile: they take no part in load /
  // convert, so that both groups reach the barrier behind the conversion at about the same time
  const bool issuer = warp == 1 || (MB == 2 && warp == 5);
  const int ctid = (warp - (warp > 1) - (MB == 2 && warp > 5)) * 32 + lane;  // index among the converter threads
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
  // fp16x2: largest magnitude of the loaded tile -> tmx[slot] (order-independent, hence deterministic)
  auto publish_max = [&](int slot) {
    if (!Cfg::F16 || issuer) return;
    uint32_t m = 0u;
#pragma unroll
    for (int k = 0; k < NU; ++k) {
      m = max(max(max(m, absbits(ra[k].x)), max(absbits(ra[k].y), absbits(ra[k].z))), absbits(ra[k].w));
      m = max(max(max(m, absbits(rb[k].x)), max(absbits(rb[k].y), absbits(rb[k].z))), absbits(rb[k].w));
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0) atomicMax(&tmx[slot], m);
  };
  auto convert = [&](float scale) {
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
        stage8<MODE, MB, TMA_>(A1, r, oct, a, b, scale);
      }
    }
    fence_proxy_async();
  };

  // everything above read only this layer's weights: the previous kernel of the stream may still be running (common.cuh)
  pdl_launch_dependents();
  pdl_wait();
  int tile = blockIdx.x;
  float amax = 0.f, sa1 = 1.f;  // fp16x2: largest |stage-1 input| of the current tile and its operand scale
  int slot = 0;                 // tmx slot of the NEXT tile
  if (tile < p.total_tiles) {
    load(tile);
    if (Cfg::F16) {
      publish_max(0);
      __syncthreads();
      amax = __uint_as_float(tmx[0]);
      sa1 = pow2_scale(amax);
      slot = 1;
    }
    convert(sa1);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (issuer && elect_one()) {
      issue_stage<MODE, MB, TMA_>(tmem, smem_u32(A1), 0, p.d1, smem_u32(W1), warp >> 2);
      commit(&bar[0]);
    }
    if (Cfg::F16 && tid == 0) tmx[0] = 0u;  // read by every thread before the barrier above
    __syncwarp();
  }
  uint32_t phase = 0;
  int pend_t0 = 0, pend_b = -1;  // TMA: tile whose out2 image waits in OUT2 for its bulk store
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
    uint32_t mb1 = 0u, mb2 = 0u;  // this thread's row / half of the sign masks
    if (p.m1bits && inrange) mb1 = p.m1bits[((size_t)boff + g) * 2 + half];
    if (p.m2bits && inrange) mb2 = p.m2bits[((size_t)boff + g) * 2 + half];
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
    const float inv1 = Cfg::F16 ? pow2_inv(sa1) * isw1 : 1.f;  // undoes the operand scales (exact: a power of two)
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(bias_s + half * 16 + c);
      if (Cfg::F16) {
        v[c] = fmaf(v[c], inv1, bv.x); v[c + 1] = fmaf(v[c + 1], inv1, bv.y);
        v[c + 2] = fmaf(v[c + 2], inv1, bv.z); v[c + 3] = fmaf(v[c + 3], inv1, bv.w);
      } else {
        v[c] += bv.x; v[c + 1] += bv.y; v[c + 2] += bv.z; v[c + 3] += bv.w;
      }
    }
    if (p.m1bits) {
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = (mb1 >> c) & 1u ? v[c] : 0.f;
    } else if (p.mask1) {
      warp_unpack_rows(f, stg, lane, m);
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = m[c] > 0.f ? v[c] : 0.f;
    }
    if (p.hbits_out) {
      const int j = i - p.d2;
      if (j >= 0 && j < Rout && g < L) {
        uint32_t hm = 0u;
#pragma unroll
        for (int c = 0; c < 16; ++c) hm |= (uint32_t)(v[c] > 0.f) << c;
        p.hbits_out[((size_t)boff + g) * 2 + half] = (uint16_t)hm;
      }
    }
    if (Cfg::TMA) {
      const int j = i - p.d2;  // row of the stored box: global row t0 + j
      if (p.out1 && j >= 0 && j < Rout) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(OUT1 + tma::swz(j, half * 4 + q)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    } else if (p.out1) {
      warp_store_rows(p.out1, boff, s0 + i0, L, half, i0, p.d2, Cfg::R - p.d2, stg, lane, v);
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float a = inrange ? v[c] : 0.f;  // rows outside [0, L) are conv2's zero padding
      v[c] = p.relu2 ? fmaxf(a, 0.f) : a;
    }
    const float sa2 = Cfg::F16 ? pow2_scale(fmaf(w1bound, amax, b1bound)) : 1.f;
#pragma unroll
    for (int q = 0; q < 2; ++q)
      stage8<MODE, MB, TMA_>(A2, Cfg::DMAX + i, half * 2 + q, make_float4(v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3]),
                   make_float4(v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]), sa2);
    if (has_next) publish_max(slot);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();  // A2 complete; every warp has drained the stage-1 accumulators; the next tile's maximum is published
    fence_after_sync();
    if (Cfg::F16 && has_next) {
      amax = __uint_as_float(tmx[slot]);
      sa1 = pow2_scale(amax);
    }
    if (issuer && elect_one()) {
      // out2 tile row i uses A2 rows DMAX + i + (j-1)*d2
      issue_stage<MODE, MB, TMA_>(tmem + MB * Cfg::NW, smem_u32(A2), Cfg::DMAX - p.d2, p.d2, smem_u32(W2), warp >> 2);
      commit(&bar[1]);
      if (Cfg::TMA && warp == 1) {  // this tile's out1 image and the previous tile's out2 image are complete
        if (pend_b >= 0) tma::store_rows(&tm_out2, OUT2, pend_t0, pend_b);
        if (p.out1) tma::store_rows(&tm_out1, OUT1, t0, b);
        tma::commit_group();
      }
    }
    __syncwarp();
    pend_t0 = t0; pend_b = b;
    if (has_next) {  // A1 is free (its MMAs completed before epilogue 1): stage the next tile under the stage-2 MMAs
      convert(sa1);
      if (Cfg::TMA && warp == 1 && elect_one()) tma::wait_read();  // both images may be overwritten after the barrier
      fence_before_sync();
      __syncthreads();
      fence_after_sync();
      if (Cfg::F16 && tid == 0) tmx[slot] = 0u;  // every thread has read it; written again two tiles from now
      slot ^= 1;
      if (issuer && elect_one()) {
        issue_stage<MODE, MB, TMA_>(tmem, smem_u32(A1), 0, p.d1, smem_u32(W1), warp >> 2);
        commit(&bar[0]);
      }
      __syncwarp();
    } else if (Cfg::TMA) {
      if (warp == 1 && elect_one()) tma::wait_read();
      __syncthreads();
    }

    // ---- epilogue 2: TMEM -> (+bias, mask, + add) -> out2 (owned rows)
    float4 f2[4];
    if (p.mask2) warp_fetch_rows(p.mask2, boff, s0 + i0, L, half, lane, f);
    if (p.add2) warp_fetch_rows(p.add2, boff, s0 + i0, L, half, lane, f2);
    mbar_wait(&bar[1], phase);
    fence_after_sync();
    tmem_ld16(taddr + MB * Cfg::NW, v);
#pragma unroll
    for (int sp = 1; sp < Cfg::S; ++sp) {
      tmem_ld16(taddr + MB * Cfg::NW + sp * 32, m);
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] += m[c];
    }
    const float inv2 = Cfg::F16 ? pow2_inv(sa2) * isw2 : 1.f;
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(bias_s + 32 + half * 16 + c);
      if (Cfg::F16) {
        v[c] = fmaf(v[c], inv2, bv.x); v[c + 1] = fmaf(v[c + 1], inv2, bv.y);
        v[c + 2] = fmaf(v[c + 2], inv2, bv.z); v[c + 3] = fmaf(v[c + 3], inv2, bv.w);
      } else {
        v[c] += bv.x; v[c + 1] += bv.y; v[c + 2] += bv.z; v[c + 3] += bv.w;
      }
    }
    if (p.m2bits) {
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = (mb2 >> c) & 1u ? v[c] : 0.f;
    } else if (p.mask2) {
      warp_unpack_rows(f, stg, lane, m);
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = m[c] > 0.f ? v[c] : 0.f;
    }
    if (p.add2) {
      warp_unpack_rows(f2, stg, lane, m);
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] += m[c];
      if (p.xbits_out) {  // forward: add2 is the block input x, this thread holds its row / half
        const int j = i - p.d2;
        if (j >= 0 && j < Rout && g < L) {
          uint32_t xm = 0u;
#pragma unroll
          for (int c = 0; c < 16; ++c) xm |= (uint32_t)(m[c] > 0.f) << c;
          p.xbits_out[((size_t)boff + g) * 2 + half] = (uint16_t)xm;
        }
      }
    }
    if (Cfg::TMA) {
      const int j = i - p.d2;
      if (j >= 0 && j < Rout) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(OUT2 + tma::swz(j, half * 4 + q)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    } else {
      warp_store_rows(p.out2, boff, s0 + i0, L, half, i0, p.d2, Cfg::R - p.d2, stg, lane, v);
    }
    fence_before_sync();  // orders this tile's TMEM reads before the barriers of the next iteration
  }  // tile loop
  if (Cfg::TMA) {  // the last tile's out2 image
    fence_proxy_async();
    __syncthreads();
    if (warp == 1 && elect_one()) {
      if (pend_b >= 0) tma::store_rows(&tm_out2, OUT2, pend_t0, pend_b);
      tma::commit_group();
      tma::wait_all();
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, Cfg::TCOLS);
}

template <int MODE, int MB, bool TMA_, int D1MAX = 32, int KIND = 0>
static int launch_rb(const RbTcParams& p, cudaStream_t st) {
  using Cfg = RbCfg<MODE, MB, TMA_>;
  static bool attr_set = false;
  if (!attr_set) {
    VQB_CUDA(cudaFuncSetAttribute(rb_tc_kernel<MODE, MB, TMA_, D1MAX, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
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
  const int slots = num_sms * (3 - MB);
  const int grid = q.total_tiles < slots ? q.total_tiles : slots;
  CUtensorMap tm1, tm2;
  memset(&tm1, 0, sizeof(tm1));
  memset(&tm2, 0, sizeof(tm2));
  if (Cfg::TMA) {
    if ((p.out1 && !tma::make_rows_map(&tm1, p.out1, p.B, p.L, Rout)) || !tma::make_rows_map(&tm2, p.out2, p.B, p.L, Rout))
      return set_err(VQB_ERR_CUDA, "cuTensorMapEncodeTiled failed for a [%d, %d, 32] fp32 tensor (box rows %d)", p.B, p.L, Rout);
  }
  VQB_CUDA(launch_pdl(rb_tc_kernel<MODE, MB, TMA_, D1MAX, KIND>, dim3(grid), dim3(Cfg::NT), (size_t)Cfg::SMEM, st, q, tm1, tm2));
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

static int dispatch_rb(int precision, const RbTcParams& p, cudaStream_t st) {
  // Tile height (modes with <= 2 pieces fit two CTAs of 128-row tiles per SM): the halo a tile recomputes is 2 * d2 rows,
  // so 128-row tiles pay off while d2 is small (every forward block: d2 = 1; backward blocks of dilation <= 3).
  // VQB_RB_MB=1|2 forces one of them (tuning).
  // VQB_RB_TMA=1 (fp16x2, 256-row tiles) sends out1 / out2 through bulk-tensor stores (tma.cuh).  Measured on B200 it is
  // correct but 2-9 % SLOWER than the per-warp staged st.global path (fwd 71.4 vs 69.8 us, bwd 100 vs 91 us at
  // [32, 14080, 32]): the kernel is bound by how little of its DRAM traffic overlaps its compute phases, not by LSU work,
  // and the bulk stores bunch the writes of a tile right behind one barrier.  Kept as the groundwork for TMA-fed INPUTS.
  const char* e_mb = getenv("VQB_RB_MB");
  const char* e_tma = getenv("VQB_RB_TMA");
  const int mb_env = e_mb ? atoi(e_mb) : 0;
  const bool use_tma = e_tma && atoi(e_tma) == 1;
  const bool half_tiles = !use_tma && (mb_env ? mb_env == 1 : p.d2 <= 3);
  switch (precision) {
    case VQB_PREC_BF16: return launch_rb<0, 2, false>(p, st);
    case VQB_PREC_TF32: return launch_rb<1, 2, false>(p, st);
    case VQB_PREC_BF16X2: return half_tiles ? launch_rb<2, 1, false>(p, st) : launch_rb<2, 2, false>(p, st);
    case VQB_PREC_BF16X3: return launch_rb<3, 2, false>(p, st);
    case VQB_PREC_FP16X2:
      if (use_tma) return launch_rb<4, 2, true>(p, st);
      {  // the two shapes of a training step get their own instantiation (rb_specialize)
        const bool fwd_train = p.relu1 && p.relu2 && !p.mask1 && !p.mask2 && !p.m1bits && !p.m2bits && p.out1 && p.add2 &&
                               p.xbits_out && p.hbits_out;
        const bool bwd_bits = !p.relu1 && !p.relu2 && !p.mask1 && !p.mask2 && p.m1bits && p.m2bits && p.out1 && p.add2 &&
                              !p.xbits_out && !p.hbits_out;
        if (fwd_train)
          return half_tiles ? (p.d1 <= 9 ? launch_rb<4, 1, false, 9, 1>(p, st) : launch_rb<4, 1, false, 32, 1>(p, st))
                            : launch_rb<4, 2, false, 32, 1>(p, st);
        if (bwd_bits)
          return half_tiles ? launch_rb<4, 1, false, 9, 2>(p, st) : launch_rb<4, 2, false, 32, 2>(p, st);
        const bool fwd_infer = p.relu1 && p.relu2 && !p.mask1 && !p.mask2 && !p.m1bits && !p.m2bits && !p.out1 && p.add2 &&
                               !p.xbits_out && !p.hbits_out;
        if (fwd_infer)
          return half_tiles ? (p.d1 <= 9 ? launch_rb<4, 1, false, 9, 3>(p, st) : launch_rb<4, 1, false, 32, 3>(p, st))
                            : launch_rb<4, 2, false, 32, 3>(p, st);
      }
      return half_tiles ? (p.d1 <= 9 ? launch_rb<4, 1, false, 9>(p, st) : launch_rb<4, 1, false>(p, st))
                        : launch_rb<4, 2, false>(p, st);
  }
  return set_err(VQB_ERR_INVALID, "unknown precision %d", precision);
}

bool resblock_tc_supported(const vqb_resblock_desc* d) {
  return d->C == 32 && d->F == 32 && d->dilation >= 1 && d->dilation <= 32 &&
         d->precision >= VQB_PREC_TF32 && d->precision <= VQB_PREC_FP16X2;
}

int resblock_fwd_tc(const vqb_resblock_desc* d, const float* x, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* h, float* y, uint32_t* xbits, uint32_t* hbits, cudaStream_t st) {
  if (!resblock_tc_supported(d))
    return set_err(VQB_ERR_UNIMPLEMENTED, "tensor-core residual block: C=F=32, dilation<=32, precision bf16|tf32 only (got C=%d F=%d dil=%d prec=%d)",
                   d->C, d->F, d->dilation, d->precision);
  if (d->B == 0 || d->L == 0) return VQB_OK;
  RbTcParams p{};
  p.in1 = x; p.out1 = h; p.add2 = x; p.out2 = y;
  p.xbits_out = reinterpret_cast<uint16_t*>(xbits); p.hbits_out = reinterpret_cast<uint16_t*>(hbits);
  p.w1 = w1; p.bias1 = b1; p.sj1 = 32 * 32; p.si1 = 32; p.so1 = 1; p.flip1 = 0;   // B[n=co][k=ci] = W1[j][ci][co]
  p.w2 = w2; p.bias2 = b2; p.sj2 = 32 * 32; p.si2 = 32; p.so2 = 1; p.flip2 = 0;
  p.B = d->B; p.L = d->L; p.d1 = d->dilation; p.d2 = 1; p.relu1 = 1; p.relu2 = 1;
  return dispatch_rb(d->precision, p, st);
}

int resblock_bwd_tc(const vqb_resblock_desc* d, const float* x, const float* h, const float* dy, const float* w1,
                    const float* w2, float* dh, float* dx, const uint32_t* xbits, const uint32_t* hbits, cudaStream_t st) {
  if (!resblock_tc_supported(d))
    return set_err(VQB_ERR_UNIMPLEMENTED, "tensor-core residual block backward: unsupported shape / precision");
  if (d->B == 0 || d->L == 0) return VQB_OK;
  RbTcParams p{};
  p.in1 = dy; p.mask1 = h; p.out1 = dh; p.mask2 = x; p.add2 = dy; p.out2 = dx;
  if (xbits && hbits) {  // sign masks from the forward kernel replace the two fp32 tensors
    p.mask1 = nullptr; p.mask2 = nullptr;
    p.m1bits = reinterpret_cast<const uint16_t*>(hbits); p.m2bits = reinterpret_cast<const uint16_t*>(xbits);
  }
  // stage 1 = conv2^T: dA[t][f] = sum_j sum_c W2[j][f][c] dy[t + (1-j)*1][c]  -> tap n = 2-j, B[n=f][k=c]
  p.w1 = w2; p.sj1 = 32 * 32; p.si1 = 1; p.so1 = 32; p.flip1 = 1;
  // stage 2 = conv1^T with the block's dilation
  p.w2 = w1; p.sj2 = 32 * 32; p.si2 = 1; p.so2 = 32; p.flip2 = 1;
  p.B = d->B; p.L = d->L; p.d1 = 1; p.d2 = d->dilation; p.relu1 = 0; p.relu2 = 0;
  return dispatch_rb(d->precision, p, st);
}

}  // namespace vqb
