// resstack_tc.cu — a whole DilatedResnet1D (resnet.py:40-59: up to 4 pre-activation residual blocks = 8 k=3 32->32
// convolutions) and its data-gradient chain as ONE persistent, warp-specialised tcgen05 kernel.
//
//   per tile of R = 128 * MB time rows (MB = 4, or 3 for short sequences) of one batch item, halo H = sum(dilation_i + 1) rows
//   per side (R - 2 H of the rows are output rows):
//     producer warp : TMA-loads the raw fp32 [R, 32] input rows (3-D tensor map, 128-byte swizzle; rows outside [0, L) arrive as
//                     zeros = the SAME padding; the next tile is prefetched into L2) and, per convolution, the pre-packed
//                     fp16x2 operand image of its weights (bulk copy)
//     issuer warp   : per convolution and M block 12 tcgen05.mma (3 taps = the SAME operand tile addressed with a row shift of
//                     (tap-1)*dilation, 2 K steps, 2 piece instructions N = 64 + N = 32), accumulators in tensor memory
//     4 * MB epilogue warps (thread = tile row = TMEM lane; its 32 channels as two halves): TMEM -> registers, scale / bias /
//                     sign mask / residual (kept in registers across the whole stack), -> the next convolution's operand row
//                     (ReLU folded into a truncating hi / lo fp16 split, power-of-two scale), and — only where the caller
//                     wants a tensor back — a swizzled row image that leaves as the warp's own TMA store
//   ONE operand buffer, rewritten in place: the epilogue of (convolution k, M block mb) writes the operand rows of convolution
//   k+1 over those of k.  Safe because every MMA of k that reads rows of block mb — its own three taps, tap +1 of block mb-1
//   (head rows) and tap -1 of block mb+1 (tail rows) — is issued before the commit that releases the block (issue order per
//   convolution: t-1(0) t0(0) | t-1(1) t+1(0) commit0 | t0(1) | t-1(2) t+1(1) commit1 | ...).  Hand-offs are mbarriers per M
//   block, so the tensor pipe works on one block while the epilogue warps of the others convert; the raw input of the NEXT tile
//   lands in the same buffer once the last convolution's MMAs are through, under the last epilogue.
//
// Arithmetic = the fp16x2 mode of resblock_tc.cu (operands scaled by a power of two and split into two fp16 pieces, three
// piece products, fp32 accumulation).  Operand scales: exact maximum of the tile input (one reduction per tile), then for every
// convolution the bound  L1(W) * max|input| (+ bound of the residual stream) + max|bias|  with max|input| the EXACT maximum of
// the previous convolution's operand (published through shared memory by its epilogues, complete by the time it is needed: the
// epilogue of convolution k waits for all epilogues of k-1, which finish under the MMAs of k anyway).
//
// What bounds it (profiles/README.md): tensor memory is read at 32 B/clk/SM, and the [hi | lo] accumulator blocks are 64 fp32
// columns per row and convolution.
//
// Three instantiations: KIND 0 inference forward (stores y only: 256 B per position for the whole stack), KIND 1 training
// forward (stores h_i, y_i and both sign masks per block), KIND 2 data gradient (masks from the sign words, stores dh_i, dx_i).
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "tc.cuh"
#include "tma.cuh"

namespace vqb {

using namespace tc;

constexpr int RS_MAXC = 2 * VQB_RESSTACK_MAX_BLOCKS;  // convolutions per launch

struct RsW {  // packed weight record of one convolution (rs_pack_kernel): operand image + what the epilogues need
  static constexpr int NP = 4;                          // 16-byte planes (8 fp16 channels each) per piece
  static constexpr int NW = 64, WPLANE = NW * 16, WTAP = NP * WPLANE, WCONV = 3 * WTAP;  // image: [tap][plane][hi 32 | lo 32 rows][16 B]
  static constexpr int META = 36;                       // floats: weight scale, L1 bound, max|bias|, -, bias[32]
  static constexpr int WREC = WCONV + 256;              // record in global memory: image + META floats (padded)
};
using RsCfg = RsW;

struct RsParams {
  const uint8_t* wpack;                // nconv records of RsCfg::WREC bytes (rs_pack_kernel)
  int B, L, nconv, H, Rout, tiles_x, total_tiles;
  int dil[RS_MAXC];
  int store[RS_MAXC];                  // 1: the output of convolution k is stored (tensor maps out[k][*])
  uint32_t* bits_in0;                  // KIND 1: sign mask of the chain input
  uint32_t* bits_out[RS_MAXC];         // KIND 1: sign mask of the output of convolution k (or NULL)
  const uint32_t* bits_mask[RS_MAXC];  // KIND 2: sign mask applied to the output of convolution k
  long long* trace;                    // TRACE builds: clock64 stamps of CTA 0's second tile, [role 0..4][k 0..8][event 0..7]
};
struct RsMaps {
  CUtensorMap in;                      // box {32, 128, 1}
  CUtensorMap out[RS_MAXC][2];         // per-warp slices {32 channels, rows}: [k][1]: 32 rows, [k][0]: the partial quadrant at the
                                       // two ends of a tile's output rows (32 - H % 32 rows)
};

struct RsPackParams {  // up to 2 * RS_MAXC records: a training forward also packs the images of its data-gradient chain
  const float* w[2 * RS_MAXC];
  const float* bias[2 * RS_MAXC];
  int sj[2 * RS_MAXC], si[2 * RS_MAXC], so[2 * RS_MAXC], flip[2 * RS_MAXC];
  uint8_t* out;
  int early;  // launched with the programmatic-serialization attribute: `out` is private to the stack (nobody else touches it), so
              // the whole kernel may run before the previous kernel of the stream has completed (vqb_resstack_fwd_private_ws)
};

// One block per convolution: element (tap j, in-channel k, out-channel n) at w[jj*sj + k*si + n*so] (jj = flip ? 2-j : j) ->
// fp16 hi / lo pieces of w * 2^s in the operand image, plus the numbers the epilogues need: the scale, the exact worst-case
// gain  max_n sum_{j,k} |w|  (L1 bound) and the bias.
__global__ void __launch_bounds__(256) rs_pack_kernel(const RsPackParams p) {
  const int c = blockIdx.x, tid = threadIdx.x;
  if (p.early) pdl_launch_dependents();
  __shared__ float l1p[8][32];
  __shared__ uint32_t mx;
  if (tid == 0) mx = 0u;
  __syncthreads();
  float wv[12];
  uint32_t m = 0u;
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    const int e = tid + q * 256, n = e & 31, k = (e >> 5) & 31, j = e >> 10;
    wv[q] = p.w[c][(size_t)(p.flip[c] ? 2 - j : j) * p.sj[c] + (size_t)k * p.si[c] + (size_t)n * p.so[c]];
    m = max(m, absbits(wv[q]));
  }
  // L1 norms per output channel n = lane (all 12 elements of a thread share n): per-warp partials, summed in fixed order
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 12; ++q) s += fabsf(wv[q]);
  l1p[tid >> 5][tid & 31] = s;
  m = __reduce_max_sync(0xffffffffu, m);
  if ((tid & 31) == 0) atomicMax(&mx, m);
  __syncthreads();
  const float sw = pow2_scale(__uint_as_float(mx));
  uint8_t* img = p.out + (size_t)c * RsCfg::WREC;
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    const int e = tid + q * 256, n = e & 31, k = (e >> 5) & 31, j = e >> 10;
    uint8_t* a = img + j * RsCfg::WTAP + (k >> 3) * RsCfg::WPLANE + n * 16 + (k & 7) * 2;
    const __half hi = __float2half_rn(wv[q] * sw);
    const __half lo = __float2half_rn(wv[q] * sw - __half2float(hi));
    *reinterpret_cast<__half*>(a) = hi;
    *reinterpret_cast<__half*>(a + 32 * 16) = lo;
  }
  float* meta = reinterpret_cast<float*>(img + RsCfg::WCONV);
  if (tid < 32) {
    float l = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) l += l1p[w][tid];
    l = fmaxf(l, __shfl_xor_sync(0xffffffffu, l, 16)); l = fmaxf(l, __shfl_xor_sync(0xffffffffu, l, 8));
    l = fmaxf(l, __shfl_xor_sync(0xffffffffu, l, 4)); l = fmaxf(l, __shfl_xor_sync(0xffffffffu, l, 2));
    l = fmaxf(l, __shfl_xor_sync(0xffffffffu, l, 1));
    const float b = p.bias[c] ? p.bias[c][tid] : 0.f;
    uint32_t bm = __reduce_max_sync(0xffffffffu, absbits(b));
    meta[4 + tid] = b;
    if (tid == 0) { meta[0] = sw; meta[1] = l * 1.0001f; meta[2] = __uint_as_float(bm); meta[3] = 0.f; }
  }
  if (p.early) pdl_wait();  // completion stays transitive: whoever waits for this kernel has waited for its predecessors too
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void prefetch_rows_l2(const CUtensorMap* tm, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tm), "r"(0), "r"(c1), "r"(c2) : "memory");
}
// 16 fp32 columns of this thread's TMEM lane, without the wait (tmem_ld_wait() below covers any number of them)
__device__ __forceinline__ void tmem_ld16_nw(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// fp32 -> two fp16 pieces by TRUNCATION: hi = the value with its mantissa cut to fp16's 10 bits (one LOP3; exactly
// representable, so the conversion is exact), lo = rn_f16(x - hi) (exact difference, 13 bits rounded to 11): 2^-22 relative,
// one instruction per element less than the round-to-nearest split of tc.cuh (no fp16 -> fp32 unpack).  RELU folds the
// activation into the two conversions (cvt.rn.relu): for x < 0 both hi and x - hi are <= 0 and come out as +0, for x >= 0
// both are >= 0 and pass unchanged — no separate max.
template <bool RELU>
__device__ __forceinline__ uint32_t rs_cvt2(float lo_elem, float hi_elem) {
  uint32_t r;
  if (RELU) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
// 8 channels (already scaled) -> hi chunk, lo chunk; hm accumulates max |hi| (packed halves)
template <bool RELU>
__device__ __forceinline__ void rs_split8(const float* t, uint4& hi, uint4& lo, __half2& hm) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = t[2 * i], b = t[2 * i + 1];
    const float ah = __uint_as_float(__float_as_uint(a) & 0xffffe000u), bh = __uint_as_float(__float_as_uint(b) & 0xffffe000u);
    h[i] = rs_cvt2<RELU>(ah, bh);
    l[i] = rs_cvt2<RELU>(a - ah, b - bh);
    hm = __hmax2(hm, __habs2(*reinterpret_cast<const __half2*>(&h[i])));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// max of the two halves -> fp32 bits of an upper bound of max |operand| in true scale (inv = 1 / operand scale)
__device__ __forceinline__ uint32_t rs_hmax_bits(__half2 hm, float inv) {
  const float2 f = __half22float2(hm);
  return __float_as_uint(fmaxf(f.x, f.y) * inv * 1.002f);
}

// TRACE: clock64 stamps of the hand-offs of CTA 0's second tile (tools/trace_stack.py): role 0 = issuer, 1 + mb = first epilogue
// warp of M block mb
#define RS_TR(role, k, ev)                                                                                   \
  do {                                                                                                       \
    if (TRACE && p.trace && blockIdx.x == 0 && ti == 1 && lane == 0) p.trace[((role) * 9 + (k)) * 8 + (ev)] = clock64(); \
  } while (0)

template <int MBv>
struct RsTileCfg {
  static constexpr int MB = MBv, R = 128 * MB, G = 32;
  static constexpr int NP = 4;
  static constexpr int PLANE = (R + 2 * G) * 16 + 32;
  static constexpr int TILE = NP * PLANE;
  static constexpr int OPB = 2 * TILE;
  static constexpr int NW = RsCfg::NW, WPLANE = RsCfg::WPLANE, WTAP = RsCfg::WTAP, WCONV = RsCfg::WCONV, META = RsCfg::META, WREC = RsCfg::WREC;
  static constexpr int IMG = 128 * 128;
  static constexpr int NEPI = 4 * MB;
  static constexpr int OFF_OP = 0;
  static constexpr int OFF_OUT = ((OPB + 1023) / 1024) * 1024;
  static constexpr int OFF_W = OFF_OUT + NEPI * 4096;
  static constexpr int OFF_META = OFF_W + 2 * WCONV;
  static constexpr int OFF_AMAX = OFF_META + RS_MAXC * META * 4;
  static constexpr int OFF_BAR = OFF_AMAX + 2 * 16 * 4;
  static constexpr int NBAR = 2 * MB + 8;
  static constexpr int OFF_TSLOT = OFF_BAR + NBAR * 8;
  static constexpr int SMEM = OFF_TSLOT + 16 + 1024;
  static constexpr int NT = (NEPI + 2) * 32;
  static constexpr int TCOLS = MB * NW <= 256 ? 256 : 512;
  static_assert(OFF_W % 16 == 0 && OFF_BAR % 8 == 0 && R * 128 <= OPB, "layout");
  static_assert(SMEM <= 232448, "shared memory");
};
using RsCfg4 = RsTileCfg<4>;

// one tap of one M block: 2 K steps x (hi-activation x [W_hi | W_lo], lo-activation x W_hi)
template <class CfgT>
__device__ __forceinline__ void rs_tap(uint32_t tacc, uint32_t a_base, int row0, int dil, uint32_t w_base, int j, bool first) {
  const uint64_t ad0 = smem_desc(a_base + (uint32_t)row0 * 16u, CfgT::PLANE, 128);  // row0 = first row of tap 0
  const uint64_t bd0 = smem_desc(w_base, CfgT::WPLANE, 128);
#pragma unroll
  for (int kk = 0; kk < 2; ++kk)
#pragma unroll
    for (int sa = 0; sa < 2; ++sa) {
      const uint32_t idesc = instr_desc(FMT_F16, 128, 32 * (2 - sa), false, false);
      const uint64_t bd = bd0 + (uint64_t)((j * CfgT::WTAP + kk * 2 * CfgT::WPLANE) >> 4);
      const uint64_t ad = ad0 + (uint64_t)((sa * CfgT::TILE + kk * 2 * CfgT::PLANE) >> 4) + (uint64_t)(j * dil);
      mma<false>(tacc, ad, bd, idesc, (first && kk == 0 && sa == 0) ? 0u : 1u);
    }
}

template <int KIND, int MBv, bool TRACE = false>
__global__ void __launch_bounds__(RsTileCfg<MBv>::NT, 1) rs_kernel(const RsParams p, const __grid_constant__ RsMaps maps) {
  using Cfg = RsTileCfg<MBv>;
  constexpr bool FWD = KIND != 2;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* OP = smem + Cfg::OFF_OP;    // the operand buffer (all convolutions, in place); also receives the raw fp32 input tile
  uint8_t* OUT = smem + Cfg::OFF_OUT;  // one 4 KB swizzled row image per epilogue warp
  uint8_t* W = smem + Cfg::OFF_W;
  float* meta = reinterpret_cast<float*>(smem + Cfg::OFF_META);
  uint32_t* amax = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_AMAX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* mma_done = bars;
  uint64_t* ready = bars + Cfg::MB;
  uint64_t* w_full = bars + 2 * Cfg::MB;
  uint64_t* w_empty = w_full + 2;
  uint64_t* allepi = w_empty + 2;
  uint64_t* in_full = allepi + 1;
  uint64_t* op_empty = in_full + 1;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_TSLOT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nconv = p.nconv, L = p.L;

  if (warp == 0) tmem_alloc(tslot, Cfg::TCOLS);
  if (tid == 32) {
    for (int i = 0; i < Cfg::MB; ++i) { mbar_init(&mma_done[i], 1); mbar_init(&ready[i], 4); }
    mbar_init(&w_full[0], 1); mbar_init(&w_full[1], 1); mbar_init(&w_empty[0], 1); mbar_init(&w_empty[1], 1);
    mbar_init(allepi, Cfg::NEPI); mbar_init(in_full, 1); mbar_init(op_empty, 1);
    fence_mbar_init();
  }
  if (tid < 32) amax[tid] = 0u;
  pdl_launch_dependents();
  pdl_wait();
  for (int e = tid; e < nconv * Cfg::META; e += Cfg::NT)
    meta[e] = reinterpret_cast<const float*>(p.wpack + (size_t)(e / Cfg::META) * Cfg::WREC + Cfg::WCONV)[e % Cfg::META];
  fence_proxy_async();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tslot;

  if (warp == Cfg::NEPI + 1) {
    // ------------------------------------------------------------------------------------------- producer
    if (elect_one()) {
      uint32_t wuse[2] = {0u, 0u};
      int ti = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
        const int b = tile / p.tiles_x;
        const int g0 = (tile - b * p.tiles_x) * p.Rout - p.H;
        if (ti > 0) mbar_wait(op_empty, (uint32_t)(ti - 1) & 1u);  // the last convolution of the previous tile has read the buffer
        tma::expect_tx(in_full, Cfg::R * 128);
#pragma unroll
        for (int j = 0; j < Cfg::MB; ++j) tma::load_rows(&maps.in, OP + j * Cfg::IMG, in_full, g0 + j * 128, b);
        const int next = tile + gridDim.x;
        if (next < p.total_tiles) {
          const int nb = next / p.tiles_x;
          const int ng0 = (next - nb * p.tiles_x) * p.Rout - p.H;
#pragma unroll
          for (int j = 0; j < Cfg::MB; ++j) prefetch_rows_l2(&maps.in, ng0 + j * 128, nb);
        }
        for (int k = 0; k < nconv; ++k) {
          const int s = k & 1;
          if (wuse[s] > 0) mbar_wait(&w_empty[s], (wuse[s] - 1u) & 1u);
          tma::expect_tx(&w_full[s], Cfg::WCONV);
          bulk_g2s(W + s * Cfg::WCONV, p.wpack + (size_t)k * Cfg::WREC, Cfg::WCONV, &w_full[s]);
          ++wuse[s];
        }
      }
    }
    __syncwarp();
  } else if (warp == Cfg::NEPI) {
    // --------------------------------------------------------------------------------------------- issuer
    if (elect_one()) {
      uint32_t rdy[Cfg::MB], wf[2] = {0u, 0u};
#pragma unroll
      for (int i = 0; i < Cfg::MB; ++i) rdy[i] = 0u;
      const uint32_t a_base = smem_u32(OP);
      int ti = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
        for (int k = 0; k < nconv; ++k) {
          const int s = k & 1;
          mbar_wait(&w_full[s], wf[s] & 1u); ++wf[s];
          RS_TR(0, k, 0);
          const int dil = p.dil[k];
          const uint32_t wb = smem_u32(W + s * Cfg::WCONV);
          auto tap = [&](int mb, int j, bool first) { rs_tap<Cfg>(tmem + mb * Cfg::NW, a_base, Cfg::G + mb * 128 - dil, dil, wb, j, first); };
          auto wait_ready = [&](int mb) { mbar_wait(&ready[mb], rdy[mb] & 1u); ++rdy[mb]; fence_after_sync(); };
          // taps: 0 reads rows of blocks mb-1, mb; 1 of mb; 2 of mb, mb+1.  Every MMA that reads rows of block mb goes out before
          // commit(mb): the block's epilogue overwrites those rows.
          wait_ready(0);
          tap(0, 0, true); tap(0, 1, false);
#pragma unroll
          for (int mb = 0; mb < Cfg::MB; ++mb) {
            if (mb + 1 < Cfg::MB) {
              wait_ready(mb + 1);
              tap(mb + 1, 0, true);   // tail rows of block mb
            }
            tap(mb, 2, false);        // head rows of block mb+1 (operand of this convolution: not yet overwritten)
            commit(&mma_done[mb]);
            RS_TR(0, k, 1 + mb);
            if (mb + 1 < Cfg::MB) tap(mb + 1, 1, false);
          }
          commit(&w_empty[s]);
          if (k == nconv - 1) commit(op_empty);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------------ epilogues
    const int mb = warp >> 2, qd = warp & 3;
    const int r = warp * 32 + lane;  // tile row = TMEM lane (mb * 128 + qd * 32 + lane)
    const uint32_t taddr = tmem + (((uint32_t)qd * 32u) << 16) + (uint32_t)(mb * Cfg::NW);
    const int q0 = warp * 32;
    const int srow = max(q0, p.H), snum = max(min(q0 + 32, Cfg::R - p.H) - srow, 0);
    const bool own = r >= srow && r < srow + snum;
    const int jrow = (r - srow) & 31;
    const uint32_t img = smem_u32(OUT) + (uint32_t)warp * 4096u;
    const uint32_t imgrow = img + (uint32_t)jrow * 128u;
    const uint32_t jx = (uint32_t)(jrow & 7);
    const int mapsel = snum == 32 ? 1 : 0;
    const uint32_t oprow = smem_u32(OP) + (uint32_t)((Cfg::G + r) * 16);
    uint32_t md = 0u, ae = 0u;
    float res[32];
    float sa = 1.f, rbound = 0.f;
    bool inrange = false;
    size_t grow = 0;
    int g0 = 0, b = 0, ti = 0;
    uint32_t* am = amax;

    // 16 channels (half hf) -> hi / lo operand chunks of planes 2 hf, 2 hf + 1
    auto write_operand = [&](int hf, const float* v, float scale, __half2& hm) {
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        float t[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) t[c] = v[o * 8 + c] * scale;
        uint4 ph, pl;
        rs_split8<FWD>(t, ph, pl, hm);
        sts128(oprow + (hf * 2 + o) * Cfg::PLANE, ph);
        sts128(oprow + (hf * 2 + o) * Cfg::PLANE + Cfg::TILE, pl);
      }
    };

    auto epilogue = [&](int k, auto stage_tag) {
      constexpr int STAGE = decltype(stage_tag)::value;
      const float* mt = meta + k * Cfg::META;
      const bool last = k == nconv - 1;
      const bool store = p.store[k] != 0 && snum > 0;
      uint32_t mword = 0xffffffffu;
      if (KIND == 2) mword = inrange ? p.bits_mask[k][grow] : 0u;
      mbar_wait(&mma_done[mb], md & 1u); ++md;
      fence_after_sync();
      if (qd == 0) RS_TR(1 + mb, k, 0);
      if (k > 0) { mbar_wait(allepi, ae & 1u); ++ae; }  // all epilogues of convolution k-1: am[k] is final
      if (qd == 0) RS_TR(1 + mb, k, 1);
      const float inv = pow2_inv(sa) * pow2_inv(mt[0]);
      float bound = fmaf(mt[1], __uint_as_float(am[k]), mt[2]);
      if (STAGE) { bound += rbound; rbound = bound; }
      const float sa_next = pow2_scale(bound);
      const float sa_row = inrange ? sa_next : 0.f;  // rows outside [0, L) are the next convolution's zero padding
      if (store) {
        if (lane == 0) tma::wait_read();   // the warp's image: its previous store has been read out of shared memory
        __syncwarp();
      }
      __half2 hm = __float2half2_rn(0.f);
      uint32_t bits = 0u;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t ra[16], rb[16];
        tmem_ld16_nw(taddr + hf * 16, ra); tmem_ld16_nw(taddr + 32 + hf * 16, rb);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = fmaf(__uint_as_float(ra[c]) + __uint_as_float(rb[c]), inv, mt[4 + hf * 16 + c]);
        if (KIND == 2) {
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] = (mword >> (hf * 16 + c)) & 1u ? v[c] : 0.f;
        }
        if (STAGE) {
#pragma unroll
          for (int c = 0; c < 16; ++c) { res[hf * 16 + c] += v[c]; v[c] = res[hf * 16 + c]; }
        }
        if (KIND == 1) {
#pragma unroll
          for (int c = 0; c < 16; ++c) bits |= (uint32_t)(v[c] > 0.f) << (hf * 16 + c);
        }
        if (store && own) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            sts128(imgrow + ((((uint32_t)(hf * 4 + c)) ^ jx) << 4),
                   make_uint4(__float_as_uint(v[4 * c]), __float_as_uint(v[4 * c + 1]), __float_as_uint(v[4 * c + 2]), __float_as_uint(v[4 * c + 3])));
        }
        if (!last) write_operand(hf, v, sa_row, hm);
      }
      if (KIND == 1 && p.bits_out[k] && own && inrange) p.bits_out[k][grow] = bits;
      if (!last) {
        uint32_t mm = rs_hmax_bits(hm, pow2_inv(sa_next));
        mm = __reduce_max_sync(0xffffffffu, mm);
        if (lane == 0) atomicMax(&am[k + 1], mm);
        sa = sa_next;
      }
      if (qd == 0) RS_TR(1 + mb, k, 2);
      fence_proxy_async();
      fence_before_sync();
      __syncwarp();
      if (qd == 0) RS_TR(1 + mb, k, 3);
      if (lane == 0) {
        if (store) {
          tma::store_rows(&maps.out[k][mapsel], reinterpret_cast<const void*>(OUT + warp * 4096), g0 + srow, b);
          tma::commit_group();
        }
        if (!last) { mbar_arrive(allepi); mbar_arrive(&ready[mb]); }
      }
    };

    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
      const int par = ti & 1;
      am = amax + par * 16;
      b = tile / p.tiles_x;
      g0 = (tile - b * p.tiles_x) * p.Rout - p.H;
      const int g = g0 + r;
      inrange = g >= 0 && g < L;
      grow = (size_t)b * L + (size_t)(inrange ? g : 0);
      // ---- input phase: raw fp32 row -> residual registers, tile maximum, first operand (written over the raw tile: in place)
      mbar_wait(in_full, (uint32_t)ti & 1u);
      {
        const uint8_t* raw = OP + (r >> 7) * Cfg::IMG;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 f = *reinterpret_cast<const float4*>(raw + tma::swz(r & 127, c));
          res[4 * c] = f.x; res[4 * c + 1] = f.y; res[4 * c + 2] = f.z; res[4 * c + 3] = f.w;
        }
      }
      uint32_t m = 0u;
#pragma unroll
      for (int c = 0; c < 32; ++c) m = max(m, absbits(res[c]));
      m = __reduce_max_sync(0xffffffffu, m);
      if (lane == 0) atomicMax(&am[0], m);
      if (KIND == 1 && p.bits_in0 && own && inrange) {
        uint32_t w = 0u;
#pragma unroll
        for (int c = 0; c < 32; ++c) w |= (uint32_t)(res[c] > 0.f) << c;
        p.bits_in0[grow] = w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(allepi);
      mbar_wait(allepi, ae & 1u); ++ae;   // every raw row is in registers: the buffer may be rewritten
      if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) amax[(par ^ 1) * 16 + i] = 0u;
      }
      for (int e = tid; e < 2 * Cfg::G * 8; e += Cfg::NEPI * 32) {  // guard rows (the raw tile passed over them): 2 G rows x 8 chunk columns
        const int q = e / (2 * Cfg::G), gr = e % (2 * Cfg::G);
        const int prow = gr < Cfg::G ? gr : Cfg::R + gr;
        *reinterpret_cast<uint4*>(OP + (q >> 2) * Cfg::TILE + (q & 3) * Cfg::PLANE + prow * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
      rbound = __uint_as_float(am[0]);
      sa = pow2_scale(rbound);
      {
        __half2 hm = __float2half2_rn(0.f);
        write_operand(0, res, sa, hm);
        write_operand(1, res + 16, sa, hm);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ready[mb]);
      for (int k = 0; k < nconv; k += 2) {
        epilogue(k, std::integral_constant<int, 0>());
        epilogue(k + 1, std::integral_constant<int, 1>());
      }
      // the last epilogue does not arrive on allepi: keep the phase count even with the arrivals (one wait per phase)
    }
    if (lane == 0) tma::wait_all();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, Cfg::TCOLS);
}

// ---------------------------------------------------------------------------------------------------------- host side
bool resstack_tc_supported(const vqb_resstack_desc* d) {
  if (!d || d->C != 32 || d->precision != VQB_PREC_FP16X2 || d->n_blocks < 1 || d->n_blocks > VQB_RESSTACK_MAX_BLOCKS) return false;
  int H = 0;
  for (int i = 0; i < d->n_blocks; ++i) {
    if (d->dilations[i] < 1 || d->dilations[i] > RsCfg4::G - 1) return false;
    H += d->dilations[i] + 1;
  }
  return H <= 64;
}

size_t resstack_tc_workspace_bytes(const vqb_resstack_desc* d) { return (size_t)2 * RS_MAXC * RsCfg::WREC + 256; }

template <int KIND, int MBv, bool TRACE = false>
static int launch_rs(const RsParams& p, const RsMaps& maps, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    VQB_CUDA(cudaFuncSetAttribute(rs_kernel<KIND, MBv, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, RsTileCfg<MBv>::SMEM));
    attr_set = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    VQB_CUDA(cudaGetDevice(&dev));
    VQB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  VQB_CUDA(launch_pdl(rs_kernel<KIND, MBv, TRACE>, dim3(grid), dim3(RsTileCfg<MBv>::NT), (size_t)RsTileCfg<MBv>::SMEM, st, p, maps));
  VQB_LAUNCH_CHECK();
  return VQB_OK;
}

// Tile height: 512 rows (4 M blocks) or 384.  The larger tile loses fewer rows to the halo (424 vs 296 output rows per tile at halo
// 44) and wins as soon as every SM gets more than one tile; for one wave or less the shorter tile finishes sooner (measured at
// B = 32, forward under a tape: L = 14080: 167 vs 190 us, 3520: 50 vs 58 us; L = 1760: 42 vs 39 us).  640-row tiles (5 M blocks,
// 80 registers per thread) measured no better than 512 (116 / 167 / 190 us against 123 / 168 / 187 us for inference / tape
// forward / data gradient).  VQB_RS_MB=3|4 forces a tile height (tuning / A-B timing).
static int pick_mb(const vqb_resstack_desc* d, int H) {
  const char* e = getenv("VQB_RS_MB");
  if (e) return atoi(e) == 4 ? 4 : 3;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) num_sms = 148;
  }
  const long tiles4 = (long)cdiv(d->L, RsCfg4::R - 2 * H) * d->B;
  return 2 * tiles4 > 3L * num_sms ? 4 : 3;
}

// kind 0 / 1: forward (h == NULL -> inference), kind 2: data gradient.  A training forward (kind 1) packs the operand images of
// BOTH directions (records [0, nconv): forward, [nconv, 2 nconv): data gradient), so that the data gradient of the same step can
// run from the same workspace without a packing launch of its own (`prepacked`).
int resstack_tc(int kind, const vqb_resstack_desc* d, const float* in, const float* const* w1, const float* const* b1,
                const float* const* w2, const float* const* b2, float* const* o1, float* const* o2, uint32_t* const* xbits,
                uint32_t* const* hbits, void* ws, size_t ws_bytes, cudaStream_t st, bool prepacked = false, bool private_ws = false) {
  if (!resstack_tc_supported(d))
    return set_err(VQB_ERR_UNIMPLEMENTED, "fused residual stack: C = 32, 1..%d blocks, dilations < %d, fp16x2 only", VQB_RESSTACK_MAX_BLOCKS, RsCfg4::G);
  if (d->B == 0 || d->L == 0) return VQB_OK;
  const size_t need = resstack_tc_workspace_bytes(d);
  uint8_t* wsp = reinterpret_cast<uint8_t*>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  if (!ws || ws_bytes < need) return set_err(VQB_ERR_WORKSPACE, "vqb_resstack workspace: need %zu bytes, got %zu", need, ws_bytes);
  const int n = d->n_blocks, nconv = 2 * n;
  RsPackParams pk{};
  RsParams p{};
  alignas(64) RsMaps maps;  // filled per call, copied into the launch parameters
  memset(&maps, 0, sizeof(maps));
  pk.out = wsp;
  p.wpack = (kind == 2 && prepacked) ? wsp + (size_t)nconv * RsCfg::WREC : wsp;
  p.B = d->B; p.L = d->L; p.nconv = nconv;
  int H = 0;
  for (int i = 0; i < n; ++i) H += d->dilations[i] + 1;
  const int mbsel = pick_mb(d, H);
  p.H = H; p.Rout = 128 * mbsel - 2 * H;
  p.tiles_x = cdiv(d->L, p.Rout);
  p.total_tiles = p.tiles_x * d->B;
  float* outs[RS_MAXC];
  for (int k = 0; k < nconv; ++k) {
    const int stage = k & 1;
    if (kind != 2) {  // forward: block i = k / 2: conv1 (dilation d_i, bias b1) then conv2 (dilation 1, bias b2)
      const int i = k >> 1;
      pk.w[k] = stage ? w2[i] : w1[i];
      pk.bias[k] = stage ? (b2 ? b2[i] : nullptr) : (b1 ? b1[i] : nullptr);
      pk.sj[k] = 32 * 32; pk.si[k] = 32; pk.so[k] = 1; pk.flip[k] = 0;   // B[n = co][k = ci] = W[j][ci][co]
      p.dil[k] = stage ? 1 : d->dilations[i];
      outs[k] = stage ? (o2 ? o2[i] : nullptr) : (o1 ? o1[i] : nullptr);
      p.bits_out[k] = kind == 1 ? (stage ? (i + 1 < n ? xbits[i + 1] : nullptr) : hbits[i]) : nullptr;
    } else {          // data gradient: blocks n-1 .. 0: conv2^T (dilation 1, mask h > 0) then conv1^T (dilation d_i, mask x > 0)
      const int i = n - 1 - (k >> 1);
      pk.w[k] = stage ? w1[i] : w2[i];
      pk.bias[k] = nullptr;
      pk.sj[k] = 32 * 32; pk.si[k] = 1; pk.so[k] = 32; pk.flip[k] = 1;   // transposed, taps flipped
      p.dil[k] = stage ? d->dilations[i] : 1;
      outs[k] = stage ? o2[i] : o1[i];                                    // dx[i] / dh[i]
      p.bits_mask[k] = stage ? xbits[i] : hbits[i];
    }
    p.store[k] = outs[k] != nullptr;
  }
  if (kind == 1) p.bits_in0 = xbits[0];
  if (!p.store[nconv - 1]) return set_err(VQB_ERR_INVALID, "vqb_resstack: the output of the last block must be given");
  if (!tma::make_rows_map(&maps.in, in, d->B, d->L, 128))
    return set_err(VQB_ERR_CUDA, "cuTensorMapEncodeTiled failed for the [%d, %d, 32] input", d->B, d->L);
  for (int k = 0; k < nconv; ++k)
    if (p.store[k]) {
      const int part = 32 - H % 32;  // rows of the partial quadrant at either end of a tile's output rows
      if (!tma::make_rows_map(&maps.out[k][0], outs[k], d->B, d->L, part) || !tma::make_rows_map(&maps.out[k][1], outs[k], d->B, d->L, 32))
        return set_err(VQB_ERR_CUDA, "cuTensorMapEncodeTiled failed for an output of the residual stack");
    }
  int npack = nconv;
  if (kind == 1) {  // + the data-gradient chain of the same blocks: n-1 .. 0, conv2^T then conv1^T, no bias
    for (int k = 0; k < nconv; ++k) {
      const int i = n - 1 - (k >> 1), stage = k & 1;
      pk.w[nconv + k] = stage ? w1[i] : w2[i];
      pk.bias[nconv + k] = nullptr;
      pk.sj[nconv + k] = 32 * 32; pk.si[nconv + k] = 1; pk.so[nconv + k] = 32; pk.flip[nconv + k] = 1;
    }
    npack = 2 * nconv;
  }
  if (!(kind == 2 && prepacked)) {
    // Measured (three A/B pairs on one B200): launching the packing kernel early costs time instead of saving it (7.31-7.38 ms
    // per training step against 7.23-7.25 with the plain launch) — its CTAs land in the middle of the persistent kernel before
    // it.  So the early launch stays behind VQB_RS_EARLY_PACK=1; the private workspace itself (no allocation per call) is kept.
    static const bool early_ok = getenv("VQB_RS_EARLY_PACK") && atoi(getenv("VQB_RS_EARLY_PACK")) == 1;
    if (private_ws && early_ok) {  // the images go to memory no earlier kernel can be using: packing overlaps the predecessor's run
      pk.early = 1;
      VQB_CUDA(launch_pdl(rs_pack_kernel, dim3(npack), dim3(256), (size_t)0, st, pk));
    } else {
      rs_pack_kernel<<<npack, 256, 0, st>>>(pk);
    }
    VQB_LAUNCH_CHECK();
  }
  if (kind == 0 && mbsel == 4 && getenv("VQB_RS_TRACE")) {  // profiling aid: address of a device buffer of 5 * 9 * 8 int64 (tools/trace_stack.py)
    p.trace = reinterpret_cast<long long*>(strtoull(getenv("VQB_RS_TRACE"), nullptr, 0));
    return launch_rs<0, 4, true>(p, maps, st);
  }
  if (mbsel == 4) {
    switch (kind) {
      case 0: return launch_rs<0, 4>(p, maps, st);
      case 1: return launch_rs<1, 4>(p, maps, st);
      default: return launch_rs<2, 4>(p, maps, st);
    }
  }
  switch (kind) {
    case 0: return launch_rs<0, 3>(p, maps, st);
    case 1: return launch_rs<1, 3>(p, maps, st);
    default: return launch_rs<2, 3>(p, maps, st);
  }
}

}  // namespace vqb

using namespace vqb;

extern "C" {

int vqb_resstack_supports(const vqb_resstack_desc* d) { return resstack_tc_supported(d) ? 1 : 0; }

size_t vqb_resstack_workspace_bytes(const vqb_resstack_desc* d) { return d ? resstack_tc_workspace_bytes(d) : 0; }

int vqb_resstack_fwd(const vqb_resstack_desc* d, const float* x, const float* const* w1, const float* const* b1,
                     const float* const* w2, const float* const* b2, float* const* h, float* const* y,
                     uint32_t* const* xbits, uint32_t* const* hbits, void* workspace, size_t workspace_bytes, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(d && x && w1 && w2 && y, "vqb_resstack_fwd: NULL pointer");
  VQB_REQUIRE(d->B >= 0 && d->L >= 0, "vqb_resstack_fwd: bad shape B=%d L=%d", d->B, d->L);
  const bool train = h != nullptr;
  if (train) {
    VQB_REQUIRE(xbits && hbits, "vqb_resstack_fwd: h given without xbits / hbits (training forward stores all of them)");
    for (int i = 0; i < d->n_blocks && i < VQB_RESSTACK_MAX_BLOCKS; ++i)
      VQB_REQUIRE(h[i] && y[i] && xbits[i] && hbits[i], "vqb_resstack_fwd: training forward needs h, y, xbits, hbits of block %d", i);
  }
  return resstack_tc(train ? 1 : 0, d, x, w1, b1, w2, b2, h, y, xbits, hbits, workspace, workspace_bytes, (cudaStream_t)stream);
}

int vqb_resstack_fwd_private_ws(const vqb_resstack_desc* d, const float* x, const float* const* w1, const float* const* b1,
                                const float* const* w2, const float* const* b2, float* const* h, float* const* y,
                                uint32_t* const* xbits, uint32_t* const* hbits, void* workspace, size_t workspace_bytes, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(d && x && w1 && w2 && y, "vqb_resstack_fwd_private_ws: NULL pointer");
  VQB_REQUIRE(d->B >= 0 && d->L >= 0, "vqb_resstack_fwd_private_ws: bad shape B=%d L=%d", d->B, d->L);
  const bool train = h != nullptr;
  if (train) {
    VQB_REQUIRE(xbits && hbits, "vqb_resstack_fwd_private_ws: h given without xbits / hbits (training forward stores all of them)");
    for (int i = 0; i < d->n_blocks && i < VQB_RESSTACK_MAX_BLOCKS; ++i)
      VQB_REQUIRE(h[i] && y[i] && xbits[i] && hbits[i], "vqb_resstack_fwd_private_ws: training forward needs h, y, xbits, hbits of block %d", i);
  }
  return resstack_tc(train ? 1 : 0, d, x, w1, b1, w2, b2, h, y, xbits, hbits, workspace, workspace_bytes, (cudaStream_t)stream, false, true);
}

int vqb_resstack_bwd_data(const vqb_resstack_desc* d, const float* dy, const float* const* w1, const float* const* w2,
                          const uint32_t* const* xbits, const uint32_t* const* hbits, float* const* dh, float* const* dx,
                          void* workspace, size_t workspace_bytes, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(d && dy && w1 && w2 && xbits && hbits && dh && dx, "vqb_resstack_bwd_data: NULL pointer");
  VQB_REQUIRE(d->B >= 0 && d->L >= 0, "vqb_resstack_bwd_data: bad shape B=%d L=%d", d->B, d->L);
  for (int i = 0; i < d->n_blocks && i < VQB_RESSTACK_MAX_BLOCKS; ++i)
    VQB_REQUIRE(xbits[i] && hbits[i] && dh[i] && dx[i], "vqb_resstack_bwd_data: xbits, hbits, dh, dx of block %d", i);
  return resstack_tc(2, d, dy, w1, nullptr, w2, nullptr, dh, dx, const_cast<uint32_t* const*>(xbits),
                     const_cast<uint32_t* const*>(hbits), workspace, workspace_bytes, (cudaStream_t)stream);
}

int vqb_resstack_bwd_data_packed(const vqb_resstack_desc* d, const float* dy, const uint32_t* const* xbits,
                                 const uint32_t* const* hbits, float* const* dh, float* const* dx, void* fwd_workspace,
                                 size_t workspace_bytes, void* stream) {
  VQB_ARCH();
  VQB_REQUIRE(d && dy && xbits && hbits && dh && dx && fwd_workspace, "vqb_resstack_bwd_data_packed: NULL pointer");
  VQB_REQUIRE(d->B >= 0 && d->L >= 0, "vqb_resstack_bwd_data_packed: bad shape B=%d L=%d", d->B, d->L);
  for (int i = 0; i < d->n_blocks && i < VQB_RESSTACK_MAX_BLOCKS; ++i)
    VQB_REQUIRE(xbits[i] && hbits[i] && dh[i] && dx[i], "vqb_resstack_bwd_data_packed: xbits, hbits, dh, dx of block %d", i);
  const float* none[VQB_RESSTACK_MAX_BLOCKS] = {nullptr, nullptr, nullptr, nullptr};  // weights are not read again
  return resstack_tc(2, d, dy, none, nullptr, none, nullptr, dh, dx, const_cast<uint32_t* const*>(xbits),
                     const_cast<uint32_t* const*>(hbits), fwd_workspace, workspace_bytes, (cudaStream_t)stream, true);
}

}  // extern "C"
